"""Parity of the CUDA svb-zd kernels (through the C-ABI) against the oracle: bit-exact bytes on encode,
bit-exact samples and identical error verdicts on decode."""
import json
import os

import numpy as np
import pytest
import torch

import slow5tools_b200 as s5
from slow5tools_b200 import codec, synth
from slow5tools_b200._capi import METHOD

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "svbzd_kat.json")))


@pytest.fixture(scope="module")
def cdc():
    c = s5.Codec(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_encode(cdc, reads, out_misalign=0):
    """reads: list of int16 arrays.  Returns list of bytes (None where status != 0), status array."""
    n = np.array([len(r) for r in reads], np.uint32)
    soff = s5.sig_layout(n)
    sig = np.zeros(int(soff[-1]) + 8, np.int16)
    for r, o in zip(reads, soff):
        sig[int(o):int(o) + len(r)] = r
    ooff = s5.svb_slot_layout(n) if not out_misalign else None
    if out_misalign:  # odd slot starts: slots of bound+misalign bytes
        bound = 4 + (n.astype(np.uint64) + 3) // 4 + 3 * n.astype(np.uint64) + np.uint64(out_misalign)
        ooff = np.zeros(len(n) + 1, np.uint64)
        np.cumsum(bound, out=ooff[1:])
        ooff += np.uint64(out_misalign)
    svb = torch.full((int(ooff[-1]) + 32,), 0xAB, dtype=torch.uint8, device="cuda")
    svb_len = torch.zeros(len(n), dtype=torch.int32, device="cuda")
    status = torch.full((len(n),), 99, dtype=torch.int32, device="cuda")
    cdc.svbzd_encode_dev(dev(sig), dev(soff.view(np.int64)), dev(n.view(np.int32)), svb, dev(ooff.view(np.int64)),
                         svb_len, status)
    torch.cuda.synchronize()
    svb_h, len_h, st = svb.cpu().numpy(), svb_len.cpu().numpy(), status.cpu().numpy()
    outs = []
    for i in range(len(n)):
        o = int(ooff[i])
        outs.append(svb_h[o:o + len_h[i]].tobytes() if st[i] == 0 else None)
    # bytes outside [off, off+len) of every slot must be untouched
    mask = np.ones(svb_h.size, bool)
    for i in range(len(n)):
        mask[int(ooff[i]):int(ooff[i]) + int(len_h[i])] = False
    assert (svb_h[mask] == 0xAB).all(), "encoder wrote outside its streams"
    return outs, st


def gpu_decode(cdc, streams, in_misalign=0, caps=None):
    """streams: list of bytes.  Returns list of int16 arrays (None on error), status, n."""
    lens = np.array([len(s) for s in streams], np.uint32)
    ioff = np.zeros(len(streams) + 1, np.uint64)
    pos = in_misalign
    for i, s in enumerate(streams):
        ioff[i] = pos
        pos += len(s) + in_misalign
    ioff[-1] = pos
    cap = (pos + 15) // 16 * 16 + 16
    svb = np.full(cap, 0xCD, np.uint8)
    for s, o in zip(streams, ioff):
        svb[int(o):int(o) + len(s)] = np.frombuffer(s, np.uint8)
    ns = np.array([int.from_bytes(s[:4], "little") if len(s) >= 4 else 0 for s in streams], np.uint32)
    if caps is not None:
        ns_cap = np.array(caps, np.uint32)
    else:
        ns_cap = np.minimum(ns, 1 << 22)
    soff = s5.sig_layout(ns_cap)
    sig = torch.full((int(soff[-1]) + 8,), -21846, dtype=torch.int16, device="cuda")
    n_out = torch.zeros(len(streams), dtype=torch.int32, device="cuda")
    status = torch.full((len(streams),), 99, dtype=torch.int32, device="cuda")
    cdc.svbzd_decode_dev(dev(svb), dev(ioff.view(np.int64)), dev(lens.view(np.int32)), sig, dev(soff.view(np.int64)),
                         n_out, status)
    torch.cuda.synchronize()
    sig_h, st, nn = sig.cpu().numpy(), status.cpu().numpy(), n_out.cpu().numpy().view(np.uint32)
    outs = []
    for i in range(len(streams)):
        o = int(soff[i])
        outs.append(sig_h[o:o + int(nn[i])].copy() if st[i] == 0 else None)
    return outs, st, nn


def test_kat_vectors(cdc, oracle):
    reads = [np.array(c["in"], np.int16) for c in KAT["encode"]]
    outs, st = gpu_encode(cdc, reads)
    assert (st == 0).all()
    for c, o in zip(KAT["encode"], outs):
        assert o.hex() == c["hex"], c["name"]
    streams = [bytes.fromhex(c["hex"]) for c in KAT["encode"] + KAT["decode_only"]]
    want = [c["in"] for c in KAT["encode"]] + [c["out"] for c in KAT["decode_only"]]
    dec, st, _ = gpu_decode(cdc, streams)
    assert (st == 0).all()
    for d, w in zip(dec, want):
        assert d.tolist() == w


def test_golden_vectors_from_reference(cdc):
    g = np.load(os.path.join(HERE, "golden", "svbzd_ref_vectors.npz"))
    names = [k[4:] for k in g.files if k.startswith("in__")]
    reads = [g["in__" + k] for k in names]
    want = [g["svb__" + k].tobytes() for k in names]
    outs, st = gpu_encode(cdc, reads)
    assert (st == 0).all()
    for k, o, w in zip(names, outs, want):
        assert o == w, k
    dec, st, _ = gpu_decode(cdc, want)
    assert (st == 0).all()
    for k, d, r in zip(names, dec, reads):
        assert np.array_equal(d, r), k


@pytest.mark.parametrize("misalign", [0, 1, 7, 13])
def test_ragged_lengths_vs_oracle(cdc, oracle, misalign):
    rng = np.random.default_rng(11 + misalign)
    lens = list(range(0, 40)) + [255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097,
                                 5000, 8191, 8192, 8193, 20000, 65537]
    lens += [int(v) for v in rng.integers(0, 9000, 150)]
    base = synth.nanopore_signal(sum(lens) + 16, seed=3).numpy()
    reads, pos = [], 0
    for i, n in enumerate(lens):
        x = base[pos:pos + n].copy()
        pos += n
        if i % 5 == 1 and n:   # sprinkle big jumps (3-byte codes) and negatives
            idx = rng.integers(0, n, max(1, n // 50))
            x[idx] = rng.integers(-32768, 32768, len(idx)).astype(np.int16)
        reads.append(x)
    want = [oracle.compress(x) for x in reads]
    outs, st = gpu_encode(cdc, reads, out_misalign=misalign)
    assert (st == 0).all()
    for i, (o, w) in enumerate(zip(outs, want)):
        assert o == w, (i, lens[i])
    dec, st, nn = gpu_decode(cdc, want, in_misalign=misalign)
    assert (st == 0).all()
    for i, (d, x) in enumerate(zip(dec, reads)):
        assert np.array_equal(d, x), (i, lens[i])


@pytest.mark.parametrize("kind", ["uniform", "alternating", "constant", "boundary"])
def test_adversarial_signals(cdc, oracle, kind):
    reads = [synth.adversarial(kind, n, seed=n) for n in (1, 9, 256, 1000, 4096, 30001)]
    want = [oracle.compress(x) for x in reads]
    outs, st = gpu_encode(cdc, reads)
    assert (st == 0).all() and outs == want
    dec, st, _ = gpu_decode(cdc, want, in_misalign=3)
    assert (st == 0).all()
    assert all(np.array_equal(d, x) for d, x in zip(dec, reads))


def test_foreign_four_byte_codes(cdc, oracle):
    """Streams with 4-byte codes (never produced from int16 input, but valid svb): decode must follow
    streamvbyte_decode.c:36-58 and wrap like streamvbyte_zigzag.c:36-39."""
    rng = np.random.default_rng(2)
    streams = []
    for n in (1, 5, 64, 257, 1500):
        vals = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
        vals[rng.random(n) < 0.5] &= 0xFF
        keys = np.zeros((n + 3) // 4, np.uint8)
        data = bytearray()
        for i, v in enumerate(vals):
            v = int(v)
            c = (v > 0xFF) + (v > 0xFFFF) + (v > 0xFFFFFF)
            keys[i >> 2] |= c << ((i & 3) * 2)
            data += v.to_bytes(4, "little")[:c + 1]
        streams.append(int(n).to_bytes(4, "little") + keys.tobytes() + bytes(data))
    dec, st, _ = gpu_decode(cdc, streams, in_misalign=5)
    assert (st == 0).all()
    for s, d in zip(streams, dec):
        rc, w = oracle.depress(s)
        assert rc == 0 and np.array_equal(d, w)


def test_malformed_streams_match_oracle_verdicts(cdc, oracle):
    good = oracle.compress(synth.nanopore_signal(3000, seed=9).numpy())
    small = oracle.compress(np.arange(10, dtype=np.int16))
    bad_keys = bytearray(good)
    bad_keys[4 + 10] = 0xFF                       # claims more data bytes than present
    streams = [good, good[:-1], good + b"\0", b"\x01\x00", b"\x05\x00\x00\x00", bytes(bad_keys), small[:-1],
               small + b"\0\0\0", good[:700], good]
    want = [oracle.depress(s)[0] for s in streams]
    caps = [max(int.from_bytes(s[:4], "little") if len(s) >= 4 else 0, 8) for s in streams]
    dec, st, _ = gpu_decode(cdc, streams, caps=caps)
    assert st.tolist() == want
    assert np.array_equal(dec[0], dec[-1])


def test_slot_too_small_is_reported(cdc):
    x = synth.nanopore_signal(1000, seed=1).numpy()
    n = np.array([1000, 1000], np.uint32)
    soff = s5.sig_layout(n)
    sig = np.concatenate([x, x, np.zeros(8, np.int16)])
    ooff = np.array([0, 2000, 2000 + int(s5.lib.s5b_svbzd_slot(1000))], np.uint64)
    svb = torch.zeros(int(ooff[-1]) + 16, dtype=torch.uint8, device="cuda")
    svb_len = torch.zeros(2, dtype=torch.int32, device="cuda")
    status = torch.zeros(2, dtype=torch.int32, device="cuda")
    cdc.svbzd_encode_dev(dev(sig), dev(soff.view(np.int64)), dev(n.view(np.int32)), svb, dev(ooff.view(np.int64)),
                         svb_len, status)
    torch.cuda.synchronize()
    assert status.tolist() == [s5.ERR.NOSPACE, 0]
    assert svb_len.tolist()[0] == 0


def test_peek_and_compact(cdc, oracle):
    rng = np.random.default_rng(4)
    lens = [int(v) for v in rng.integers(0, 3000, 3000)]
    base = synth.nanopore_signal(sum(lens) + 8, seed=5).numpy()
    reads, pos = [], 0
    for n in lens:
        reads.append(base[pos:pos + n])
        pos += n
    n = np.array(lens, np.uint32)
    soff, ooff = s5.sig_layout(n), s5.svb_slot_layout(n)
    sig = np.zeros(int(soff[-1]) + 8, np.int16)
    for r, o in zip(reads, soff):
        sig[int(o):int(o) + len(r)] = r
    svb = torch.zeros(int(ooff[-1]) + 16, dtype=torch.uint8, device="cuda")
    svb_len = torch.zeros(len(n), dtype=torch.int32, device="cuda")
    status = torch.zeros(len(n), dtype=torch.int32, device="cuda")
    d_ooff = dev(ooff.view(np.int64))
    cdc.svbzd_encode_dev(dev(sig), dev(soff.view(np.int64)), dev(n.view(np.int32)), svb, d_ooff, svb_len, status)
    peek = torch.zeros(len(n), dtype=torch.int32, device="cuda")
    cdc.svbzd_peek_dev(svb, d_ooff, svb_len, peek)
    for align in (16, 1):
        dense = torch.zeros_like(svb)
        dense_off = torch.zeros(len(n) + 1, dtype=torch.int64, device="cuda")
        cdc.compact_dev(svb, d_ooff, svb_len, dense, dense_off, align=align)
        torch.cuda.synchronize()
        assert peek.tolist() == lens
        lh = svb_len.cpu().numpy().astype(np.int64)
        step = (lh + align - 1) // align * align
        assert dense_off.cpu().numpy().tolist() == np.concatenate([[0], np.cumsum(step)]).tolist()
        dh, doff = dense.cpu().numpy(), dense_off.cpu().numpy()
        for i in (0, 1, 2, 17, 1500, 2999):
            assert dh[doff[i]:doff[i] + lh[i]].tobytes() == oracle.compress(reads[i])


def test_host_slab_roundtrip_pipeline(cdc, oracle):
    """Host-buffer entry points (pinned slabs, chunked two-slot pipeline): bytes identical to the oracle."""
    rng = np.random.default_rng(8)
    lens = np.concatenate([rng.integers(0, 6000, 4000), [0, 1, 4096, 100000]]).astype(np.uint32)
    soff = s5.sig_layout(lens)
    sig = torch.zeros(int(soff[-1]) + 8, dtype=torch.int16).pin_memory()
    base = synth.nanopore_signal(int(lens.sum()) + 8, seed=6)
    pos = 0
    sig_np = sig.numpy()
    for n, o in zip(lens, soff):
        sig_np[int(o):int(o) + int(n)] = base[pos:pos + int(n)].numpy()
        pos += int(n)
    cap = int(s5.svb_slot_layout(lens)[-1]) + 64
    svb = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    svb_off = np.zeros(len(lens) + 1, np.uint64)
    svb_len = np.zeros(len(lens), np.uint32)
    status = np.zeros(len(lens), np.int32)
    old = os.environ.get("S5B_CHUNK_MB")
    os.environ["S5B_CHUNK_MB"] = "2"          # force many sub-batches through both pipeline slots
    c2 = s5.Codec(0)
    try:
        c2.svbzd_encode_host(sig, soff, lens, svb, svb_off, svb_len, status)
        assert (status == 0).all()
        svb_np = svb.numpy()
        want_out, want_len = oracle.compress_batch(sig_np, soff, lens, svb_off)
        assert np.array_equal(svb_len, want_len)
        for i in range(len(lens)):
            o = int(svb_off[i])
            assert svb_np[o:o + svb_len[i]].tobytes() == want_out[o:o + svb_len[i]].tobytes(), i
        sig2 = torch.zeros_like(sig)
        sig2_off = np.zeros(len(lens) + 1, np.uint64)
        n2 = np.zeros(len(lens), np.uint32)
        c2.svbzd_decode_host(svb, svb_off, svb_len, sig2, sig2_off, n2, status)
        assert (status == 0).all() and np.array_equal(n2, lens) and np.array_equal(sig2_off, soff)
        for i in range(len(lens)):
            o = int(soff[i])
            assert np.array_equal(sig2.numpy()[o:o + lens[i]], sig_np[o:o + lens[i]]), i
    finally:
        c2.close()
        if old is None:
            os.environ.pop("S5B_CHUNK_MB")
        else:
            os.environ["S5B_CHUNK_MB"] = old


def test_pointer_array_and_solo_forms(cdc, oracle):
    reads = [synth.nanopore_signal(n, seed=n + 1).numpy() for n in (0, 1, 10, 4096, 7777)]
    rc, outs = cdc.compress_batch(METHOD.SVB_ZD, [r.tobytes() for r in reads])
    assert rc == 0
    assert outs == [oracle.compress(r) for r in reads]
    rc, back = cdc.depress_batch(METHOD.SVB_ZD, outs + [outs[3][:-2]])
    assert rc == s5.ERR.PRESS and back[-1] is None
    assert [b for b in back[:-1]] == [r.tobytes() for r in reads]
    # slow5_ptr_compress_solo / slow5_ptr_depress_solo twins (unit_test_press.c:111-201)
    for c in KAT["encode"][:3]:
        x = np.array(c["in"], np.int16)
        enc = codec.ptr_compress_solo(METHOD.SVB_ZD, x.tobytes())
        assert enc.hex() == c["hex"]
        assert codec.ptr_depress_solo(METHOD.SVB_ZD, enc) == x.tobytes()
    assert codec.ptr_depress_solo(METHOD.SVB_ZD, b"\x05\0\0\0") is None
    assert s5.lib.s5b_last_error() == s5.ERR.PRESS
    rc, same = cdc.compress_batch(METHOD.NONE, [b"abc", b""])
    assert rc == 0 and same == [b"abc", b""]


def test_full_size_roundtrip_properties(cdc, oracle):
    """BASELINE config 2 shape at full size (100k x 4096): encode -> decode is the identity, sizes match
    the oracle on a sampled subset, and a checksum over all streams matches a second, reordered run."""
    R, N = 100_000, 4096
    sig = synth.nanopore_signal(R * N, seed=42, device="cuda")
    n = torch.full((R,), N, dtype=torch.int32, device="cuda")
    soff = torch.arange(R + 1, dtype=torch.int64, device="cuda") * N
    slot = int(s5.lib.s5b_svbzd_slot(N))
    ooff = torch.arange(R + 1, dtype=torch.int64, device="cuda") * slot
    svb = torch.zeros(R * slot + 16, dtype=torch.uint8, device="cuda")
    svb_len = torch.zeros(R, dtype=torch.int32, device="cuda")
    status = torch.ones(R, dtype=torch.int32, device="cuda")
    cdc.svbzd_encode_dev(sig, soff, n, svb, ooff, svb_len, status)
    assert int(status.abs().sum()) == 0
    back = torch.zeros_like(sig)
    n2 = torch.zeros_like(n)
    cdc.svbzd_decode_dev(svb, ooff, svb_len, back, soff, n2, status)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0 and torch.equal(n2, n) and torch.equal(back, sig)
    bps = float(svb_len.sum()) / (R * N)
    assert 1.23 <= bps <= 1.31, bps                  # calibration target of SURVEY 8d
    idx = [0, 1, 4095, 50_000, 99_999]
    lens = svb_len.cpu().numpy()
    for i in idx:
        want = oracle.compress(sig[i * N:(i + 1) * N].cpu().numpy())
        got = svb[i * slot:i * slot + int(lens[i])].cpu().numpy().tobytes()
        assert got == want, i
