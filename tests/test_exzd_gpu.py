"""Parity of the CUDA ex-zd kernels (through the C-ABI) against the oracle and the reference's golden streams:
bit-exact bytes on encode, bit-exact samples and the oracle's verdicts on decode."""
import os

import numpy as np
import pytest
import torch

import slow5tools_b200 as s5
from slow5tools_b200 import codec, synth
from slow5tools_b200._capi import METHOD, lib

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "exzd_ref_vectors.npz"))
NAMES = sorted(k[4:] for k in GOLD.files if k.startswith("in__"))


@pytest.fixture(scope="module")
def cdc():
    c = s5.Codec(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_encode(cdc, reads, misalign=0):
    n = np.array([len(r) for r in reads], np.uint32)
    soff = s5.sig_layout(n)
    sig = np.zeros(int(soff[-1]) + 8, np.int16)
    for r, o in zip(reads, soff):
        sig[int(o):int(o) + len(r)] = r
    slot = np.array([int(lib.s5b_exzd_slot(int(k))) + misalign for k in n], np.uint64)
    ooff = np.zeros(len(n) + 1, np.uint64)
    np.cumsum(slot, out=ooff[1:])
    ooff += np.uint64(misalign)
    out = torch.full((int(ooff[-1]) + 32,), 0xAB, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(len(n), dtype=torch.int32, device="cuda")
    status = torch.full((len(n),), 99, dtype=torch.int32, device="cuda")
    cdc.exzd_encode_dev(dev(sig), dev(soff.view(np.int64)), dev(n.view(np.int32)), out, dev(ooff.view(np.int64)),
                        out_len, status)
    torch.cuda.synchronize()
    o_h, l_h, st = out.cpu().numpy(), out_len.cpu().numpy(), status.cpu().numpy()
    outs = [o_h[int(ooff[i]):int(ooff[i]) + l_h[i]].tobytes() if st[i] == 0 else None for i in range(len(n))]
    mask = np.ones(o_h.size, bool)
    for i in range(len(n)):
        mask[int(ooff[i]):int(ooff[i]) + int(l_h[i])] = False
    assert (o_h[mask] == 0xAB).all(), "encoder wrote outside its streams"
    return outs, st


def gpu_decode(cdc, streams, misalign=0, caps=None):
    lens = np.array([len(s) for s in streams], np.uint32)
    ioff = np.zeros(len(streams) + 1, np.uint64)
    pos = misalign
    for i, s in enumerate(streams):
        ioff[i] = pos
        pos += len(s) + misalign
    ioff[-1] = pos
    cap = (pos + 15) // 16 * 16 + 16
    buf = np.full(cap, 0xCD, np.uint8)
    for s, o in zip(streams, ioff):
        buf[int(o):int(o) + len(s)] = np.frombuffer(s, np.uint8)
    ns = np.array([min(int.from_bytes(s[1:9], "little"), 1 << 22) if len(s) >= 9 else 0 for s in streams], np.uint32)
    if caps is not None:
        ns = np.array(caps, np.uint32)
    soff = s5.sig_layout(ns)
    sig = torch.full((int(soff[-1]) + 8,), 0x5A5A, dtype=torch.int16, device="cuda")
    n_out = torch.zeros(len(streams), dtype=torch.int32, device="cuda")
    status = torch.full((len(streams),), 99, dtype=torch.int32, device="cuda")
    cdc.exzd_decode_dev(dev(buf), dev(ioff.view(np.int64)), dev(lens.view(np.int32)), sig, dev(soff.view(np.int64)),
                        n_out, status)
    torch.cuda.synchronize()
    s_h, st, nn = sig.cpu().numpy(), status.cpu().numpy(), n_out.cpu().numpy().view(np.uint32)
    outs = [s_h[int(soff[i]):int(soff[i]) + int(nn[i])].copy() if st[i] == 0 else None for i in range(len(streams))]
    return outs, st, nn


def test_golden_streams_from_the_reference(cdc):
    reads = [GOLD["in__" + k] for k in NAMES]
    want = [GOLD["exzd__" + k].tobytes() for k in NAMES]
    got, st = gpu_encode(cdc, reads)
    assert (st == 0).all(), dict(zip(NAMES, st))
    for k, g, w in zip(NAMES, got, want):
        assert g == w, k
    back, st, nn = gpu_decode(cdc, want)
    assert (st == 0).all(), dict(zip(NAMES, st))
    for k, b, r in zip(NAMES, back, reads):
        assert np.array_equal(b, r), k


@pytest.mark.parametrize("misalign", [0, 1, 7, 13])
def test_ragged_lengths_vs_oracle(cdc, oracle, misalign):
    rng = np.random.default_rng(100 + misalign)
    lens = [1, 2, 3, 7, 8, 9, 15, 16, 17, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049] + \
        [int(v) for v in np.clip(rng.lognormal(np.log(3000), 1.0, 40), 1, 60000)]
    reads = []
    for i, n in enumerate(lens):
        x = synth.nanopore_signal(n, seed=1000 + i).numpy()
        if i % 3 == 1:
            x = ((x.astype(np.int32) >> (i % 5 + 1)) << (i % 5 + 1)).astype(np.int16)   # QTS active
        if i % 4 == 2:
            x = (x.astype(np.int32) + rng.integers(-90, 90, n)).astype(np.int16)         # many exceptions
        reads.append(x)
    got, st = gpu_encode(cdc, reads, misalign)
    assert (st == 0).all()
    want = [oracle.exzd_compress(x) for x in reads]
    for i, (g, w) in enumerate(zip(got, want)):
        assert g == w, (i, lens[i])
    back, st, nn = gpu_decode(cdc, want, misalign)
    assert (st == 0).all()
    for i, (b, r) in enumerate(zip(back, reads)):
        assert np.array_equal(b, r), (i, lens[i])


def test_exception_heavy_and_long_reads(cdc, oracle):
    rng = np.random.default_rng(7)
    reads = [
        (500 + rng.integers(-200, 200, 150000)).astype(np.int16),                 # ~36 % exceptions, key windows slide
        synth.nanopore_signal(200000, seed=5).numpy(),                            # config[3]'s longest read
        np.where(rng.random(70000) < 0.3, rng.integers(-3000, 3000, 70000), 0).astype(np.int16).cumsum().astype(np.int16),
        (rng.integers(0, 2, 5000) * 700).astype(np.int16),                        # runs of exceptions
    ]
    got, st = gpu_encode(cdc, reads)
    want = [oracle.exzd_compress(x) for x in reads]
    for i, (g, w, s) in enumerate(zip(got, want, st)):
        if len(w) > 2 * len(reads[i]) + 1024:
            assert s == s5.ERR.PRESS, i      # the reference aborts: stream outgrows its count + 1024 buffer
        else:
            assert s == 0 and g == w, i
    ok = [w for w, r in zip(want, reads)]
    back, st, nn = gpu_decode(cdc, ok)
    assert (st == 0).all()
    for b, r in zip(back, reads):
        assert np.array_equal(b, r)


def test_reference_abort_cases_are_errors(cdc):
    rng = np.random.default_rng(1)
    reads = [rng.integers(-32768, 32768, 4096).astype(np.int16), np.zeros(0, np.int16), synth.nanopore_signal(100, seed=1).numpy()]
    got, st = gpu_encode(cdc, reads)
    assert st[0] == s5.ERR.PRESS     # 4 KiB of noise: > count + 1024 bytes, SLOW5_ASSERT in the reference
    assert st[1] == s5.ERR.ARG       # empty read: undefined in the reference
    assert st[2] == 0


def test_malformed_streams_match_oracle_verdicts(cdc, oracle):
    rng = np.random.default_rng(3)
    base = [GOLD["exzd__" + k].tobytes() for k in ("one_spike", "single_exception", "noisy_many_exceptions", "nanopore_1025",
                                                   "qts_2", "nanopore_1")]
    cases = []
    for g in base:
        cases += [g, b"\x01" + g[1:], g[:10], g[:-1], g + b"\x00", g[:15], g[:17], g[:len(g) // 2]]
        for _ in range(40):
            b = bytearray(g)
            i = int(rng.integers(0, min(len(b), 600)))
            b[i] ^= 1 << int(rng.integers(0, 8))
            cases.append(bytes(b))
    want = [oracle.exzd_depress(c, cap=1 << 20) for c in cases]
    caps = [min(int.from_bytes(c[1:9], "little"), 1 << 20) if len(c) >= 9 else 0 for c in cases]
    got, st, nn = gpu_decode(cdc, cases, caps=caps)
    for i, (c, (rc, arr)) in enumerate(zip(cases, want)):
        if rc == 0:
            assert st[i] == 0 and np.array_equal(got[i], arr), i
        else:
            assert st[i] != 0, (i, rc, st[i])


def test_pointer_array_and_solo_forms(cdc, oracle):
    reads = [synth.nanopore_signal(n, seed=n).numpy() for n in (1, 100, 4096, 30001)]
    rc, outs = cdc.compress_batch(METHOD.EX_ZD, [r.tobytes() for r in reads])
    assert rc == 0
    for o, r in zip(outs, reads):
        assert o == oracle.exzd_compress(r)
    rc, back = cdc.depress_batch(METHOD.EX_ZD, outs)
    assert rc == 0
    for b, r in zip(back, reads):
        assert b == r.tobytes()
    s = codec.ptr_compress_solo(METHOD.EX_ZD, reads[2].tobytes())
    assert s == oracle.exzd_compress(reads[2])
    assert codec.ptr_depress_solo(METHOD.EX_ZD, s) == reads[2].tobytes()
    assert codec.ptr_depress_solo(METHOD.EX_ZD, b"\x01" + s[1:]) is None


def test_full_size_roundtrip_and_ratio(cdc, oracle):
    """100k x 4096 (BASELINE config[1] shape): round trip at full size, size checked on a sample against the oracle."""
    R, N = 100000, 4096
    sig = synth.nanopore_signal(R * N, seed=42, device="cuda")
    n = torch.full((R,), N, dtype=torch.int32, device="cuda")
    soff = torch.arange(R + 1, dtype=torch.int64, device="cuda") * N
    slot = int(lib.s5b_exzd_slot(N))
    ooff = torch.arange(R + 1, dtype=torch.int64, device="cuda") * slot
    out = torch.zeros(R * slot + 16, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(R, dtype=torch.int32, device="cuda")
    st = torch.ones(R, dtype=torch.int32, device="cuda")
    cdc.exzd_encode_dev(sig, soff, n, out, ooff, out_len, st)
    back = torch.zeros_like(sig)
    n2 = torch.zeros_like(n)
    st2 = torch.ones_like(st)
    cdc.exzd_decode_dev(out, ooff, out_len, back, soff, n2, st2)
    torch.cuda.synchronize()
    assert int(st.abs().sum()) == 0 and int(st2.abs().sum()) == 0
    assert torch.equal(back, sig) and torch.equal(n2, n)
    lens = out_len.cpu().numpy()
    sig_h = sig[:50 * N].cpu().numpy()
    for i in range(50):
        assert lens[i] == len(oracle.exzd_compress(sig_h[i * N:(i + 1) * N]))
    assert 1.0 < lens.mean() / N < 1.2   # ~1.05 B/sample on the calibrated signal model
