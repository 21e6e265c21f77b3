"""Synthetic nanopore-like raw signal: piecewise-constant levels with
Geometric(mean 10) dwell, level ~ N(500, 70), i.i.d. noise N(0, 9), rounded, clamped to [0, 2047].
Calibrated against the reference: svb-zd ~1.27 B/sample, zlib-L6 ratio on the record ~0.68.
Works on any torch device; deterministic for a given (seed, device type)."""
import numpy as np
import torch


def nanopore_signal(total_samples, seed=42, device="cpu", dwell_mean=10.0, level_mu=500.0, level_sigma=70.0,
                    noise_sigma=9.0, chunk=1 << 24):
    """int16 tensor of `total_samples` samples (one long trace; callers cut it into reads)."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    out = torch.empty(total_samples, dtype=torch.int16, device=device)
    done = 0
    p = 1.0 / dwell_mean
    while done < total_samples:
        m = min(chunk, total_samples - done)
        nseg = int(m / dwell_mean * 1.2) + 64
        u = torch.rand(nseg, generator=g, device=device).clamp_(1e-12, 1.0)
        dwell = torch.floor(torch.log(u) / np.log1p(-p)).to(torch.int64) + 1  # Geometric(p) on {1,2,...}
        levels = torch.randn(nseg, generator=g, device=device) * level_sigma + level_mu
        trace = torch.repeat_interleave(levels, dwell)
        while trace.numel() < m:  # (practically never) not enough segments drawn
            trace = torch.cat([trace, trace])
        trace = trace[:m] + torch.randn(m, generator=g, device=device) * noise_sigma
        out[done:done + m] = trace.round_().clamp_(0, 2047).to(torch.int16)
        done += m
    return out


def lognormal_lengths(n_reads, seed=42, mu=np.log(30000.0), sigma=1.0, lo=500, hi=200000):
    """Config-4 read lengths: round(LogNormal(mu, sigma)) clipped to [lo, hi]."""
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.lognormal(mu, sigma, n_reads)), lo, hi).astype(np.uint32)


def adversarial(kind, n, seed=1):
    """Edge-case signals of SURVEY 8d as int16 numpy arrays."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(-32768, 32768, n, dtype=np.int64).astype(np.int16)
    if kind == "alternating":
        a = np.empty(n, np.int16)
        a[0::2] = -32768
        a[1::2] = 32767
        return a
    if kind == "constant":
        return np.full(n, 777, np.int16)
    if kind == "boundary":  # deltas around the 1<->2 byte boundary
        d = rng.choice(np.array([-129, -128, -127, 126, 127, 128, 0, 1, -1]), n)
        return np.cumsum(d).astype(np.int16)
    raise ValueError(kind)


# ---- packed BLOW5 records around the signal (per-record metadata of the real fixture, no aux fields) --------
REC_ID_LEN = 36
REC_HEAD = 2 + REC_ID_LEN + 4 + 32 + 8   # bytes before the signal: u16 id_len, id, u32 read_group, 4 x f64, u64 len_raw_signal


def record_bytes(n_samples):
    """stored size of one uncompressed record (slow5_rec_to_mem binary layout, slow5.c:3928-3990) without its size prefix"""
    return REC_HEAD + 2 * int(n_samples)


def blow5_records(sig, n_reads, n_samples, seed=42):
    """uint8 tensor [n_reads, record_bytes(n_samples)] on sig's device: uncompressed (none/none) BLOW5 records, one per
    read of `n_samples` consecutive samples of `sig`: read_id = 36-char UUID-shaped string from the RNG, read_group 0,
    digitisation 8192, offset 9, range 1444.86, sampling_rate 4000 (the real fixture's values), no aux fields."""
    dev = sig.device
    R, N = int(n_reads), int(n_samples)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed) + 7919)
    rec = torch.empty((R, REC_HEAD + 2 * N), dtype=torch.uint8, device=dev)
    head = np.zeros(REC_HEAD, np.uint8)
    head[0:2] = np.frombuffer(np.uint16(REC_ID_LEN).tobytes(), np.uint8)
    at = 2 + REC_ID_LEN
    head[at:at + 4] = 0
    head[at + 4:at + 36] = np.frombuffer(np.array([8192.0, 9.0, 1444.86, 4000.0], "<f8").tobytes(), np.uint8)
    head[at + 36:at + 44] = np.frombuffer(np.uint64(N).tobytes(), np.uint8)
    rec[:, :REC_HEAD] = torch.from_numpy(head).to(dev)
    hexchars = torch.tensor(list(b"0123456789abcdef"), dtype=torch.uint8, device=dev)
    ids = hexchars[torch.randint(0, 16, (R, REC_ID_LEN), generator=g, device=dev)]
    for dash in (8, 13, 18, 23):
        ids[:, dash] = ord("-")
    rec[:, 2:2 + REC_ID_LEN] = ids
    rec[:, REC_HEAD:] = sig[:R * N].view(R, N).contiguous().view(torch.uint8).view(R, 2 * N)
    return rec
