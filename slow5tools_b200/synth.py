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
