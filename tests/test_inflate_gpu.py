"""Parity of the CUDA zlib decoder (inflate_kernel, through the C-ABI) with system zlib -- the library the
reference calls at slow5_press.c:973-1010 -- on the reference's own golden streams, real records cut from
its BLOW5 fixtures, every block type / strategy zlib can emit, truncated and corrupted streams."""
import os
import zlib

import numpy as np
import pytest
import torch

import slow5tools_b200 as s5
from slow5tools_b200 import codec, synth
from slow5tools_b200._capi import METHOD

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def cdc():
    c = s5.Codec(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ref_inflate(z):
    """What ptr_depress_zlib_solo returns: bytes, or None for Z_DATA_ERROR / Z_NEED_DICT; truncated input
    is not an error (slow5_press.c:993-1003)."""
    d = zlib.decompressobj(15)
    try:
        return d.decompress(z)
    except zlib.error:
        return None


def gpu_inflate(cdc, streams, caps=None, misalign=0):
    n = len(streams)
    lens = np.array([len(s) for s in streams], np.uint32)
    ioff = np.zeros(n + 1, np.uint64)
    pos = misalign
    for i, s in enumerate(streams):
        ioff[i] = pos
        pos += len(s) + misalign
    ioff[-1] = pos
    zin = np.full((pos + 15) // 16 * 16 + 16, 0xEE, np.uint8)
    for s, o in zip(streams, ioff):
        zin[int(o):int(o) + len(s)] = np.frombuffer(s, np.uint8)
    if caps is None:
        caps = [len(ref_inflate(s) or b"") for s in streams]
    ooff = np.zeros(n + 1, np.uint64)
    pos = misalign
    for i, c in enumerate(caps):
        ooff[i] = pos
        pos += c + misalign
    ooff[-1] = pos
    out = torch.full((pos + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    status = torch.full((n,), 99, dtype=torch.int32, device="cuda")
    cdc.zlib_inflate_dev(dev(zin), dev(ioff.view(np.int64)), dev(lens.view(np.int32)), out, dev(ooff.view(np.int64)),
                         out_len, status)
    torch.cuda.synchronize()
    oh, lh, st = out.cpu().numpy(), out_len.cpu().numpy(), status.cpu().numpy()
    res = []
    mask = np.ones(oh.size, bool)
    for i in range(n):
        o = int(ooff[i])
        res.append(oh[o:o + lh[i]].tobytes() if st[i] == 0 else None)
        mask[o:o + min(int(lh[i]), int(ooff[i + 1] - ooff[i]))] = False
        if st[i] == s5.ERR.PRESS:   # a failed stream may have written a prefix of its slot, never beyond it
            mask[o:int(ooff[i + 1])] = False
    assert (oh[mask] == 0x5A).all(), "inflate wrote outside its slots"
    return res, st, lh


def test_reference_press_golden(cdc):
    """slow5lib/test/data/exp/unit_test_exp_press: the 4 zlib streams unit_test_press.c:25-109 prints."""
    g = np.load(os.path.join(HERE, "golden", "zlib_records.npz"))
    blob = g["unit_test_exp_press"].tobytes()
    want = [b"12345\0", b"1234567890123456789012345678901234567890\0"[:41], b"hello", b"\nlol\n"]
    streams, pos = [], 0
    while pos < len(blob):
        d = zlib.decompressobj()
        d.decompress(blob[pos:])
        used = len(blob) - pos - len(d.unused_data)
        streams.append(blob[pos:pos + used])
        pos += used
    assert len(streams) == 4
    res, st, _ = gpu_inflate(cdc, streams)
    assert (st == 0).all()
    assert res == [zlib.decompress(s) for s in streams]
    assert res[0] == want[0] and res[2] == want[2] and res[3] == want[3]


def test_real_records_from_reference_fixtures(cdc):
    g = np.load(os.path.join(HERE, "golden", "zlib_records.npz"))
    names = [k[3:] for k in g.files if k.startswith("z__")]
    assert len(names) >= 20
    streams = [g["z__" + k].tobytes() for k in names]
    want = [zlib.decompress(s) for s in streams]
    for misalign in (0, 5):
        res, st, _ = gpu_inflate(cdc, streams, misalign=misalign)
        assert (st == 0).all()
        for k, r, w in zip(names, res, want):
            assert r == w, k


def corpus():
    rng = np.random.default_rng(12)
    sig = synth.nanopore_signal(60000, seed=3).numpy()
    text = (b"the quick brown fox jumps over the lazy dog. " * 400) + bytes(rng.integers(97, 123, 3000).astype(np.uint8))
    items = {
        "empty": b"", "one": b"a", "zeros_small": bytes(100), "zeros_100k": bytes(100_000),
        "text": text, "random_70k": rng.integers(0, 256, 70_000).astype(np.uint8).tobytes(),
        "raw_signal_bytes": sig.tobytes(),
        "svb_like": bytes(1030) + np.clip(rng.normal(9, 6, 5000), 0, 255).astype(np.uint8).tobytes(),
        "period_3": b"abc" * 20_000, "far_matches": (rng.integers(0, 256, 33_000).astype(np.uint8).tobytes()) * 3,
    }
    return items


def test_every_block_type_and_strategy(cdc):
    streams, want, names = [], [], []
    for name, raw in corpus().items():
        for level in (0, 1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                if level in (0, 1) and strat not in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED):
                    continue
                c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strat)
                z = c.compress(raw) + c.flush()
                streams.append(z)
                want.append(raw)
                names.append((name, level, strat))
    # multi-block streams with sync flushes (empty stored blocks in between) and small windows
    raw = corpus()["text"]
    c = zlib.compressobj(6, zlib.DEFLATED, 9)
    z = b"".join(c.compress(raw[i:i + 777]) + c.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(raw), 777)) + c.flush()
    streams.append(z); want.append(raw); names.append(("sync_flush_wbits9", 6, 0))
    res, st, _ = gpu_inflate(cdc, streams, misalign=3)
    assert (st == 0).all(), [n for n, s in zip(names, st) if s != 0]
    for n, r, w in zip(names, res, want):
        assert r == w, n


def test_trailing_bytes_are_ignored(cdc):
    z = zlib.compress(b"hello world" * 50)
    res, st, _ = gpu_inflate(cdc, [z + b"\x01\x02\x03junk"])
    assert st[0] == 0 and res[0] == b"hello world" * 50


def test_truncated_streams_return_the_decoded_prefix(cdc):
    raw = corpus()["text"][:6000] + synth.nanopore_signal(3000, seed=1).numpy().tobytes()
    z = zlib.compress(raw, 6)
    cuts = [0, 1, 2, 3, 5, 10, 50, 100, 500, len(z) // 2, len(z) - 5, len(z) - 4, len(z) - 1]
    streams = [z[:c] for c in cuts]
    want = [ref_inflate(s) for s in streams]
    res, st, _ = gpu_inflate(cdc, streams, caps=[len(raw)] * len(streams))
    assert (st == 0).all()
    for c, r, w in zip(cuts, res, want):
        assert r == w, c


def test_corrupted_streams_match_zlib_verdicts(cdc):
    rng = np.random.default_rng(7)
    raw = corpus()["svb_like"]
    z = bytearray(zlib.compress(raw, 6))
    streams = []
    for i in range(300):
        b = bytearray(z)
        k = int(rng.integers(0, len(b)))
        b[k] ^= 1 << int(rng.integers(0, 8))
        streams.append(bytes(b))
    streams += [b"\x78\x9d" + bytes(z[2:]), b"\x79\x9c" + bytes(z[2:]), b"\x78\xbb" + bytes(z[2:]), b"\x00\x00", b"\x78\x9c\x07"]
    want = [ref_inflate(s) for s in streams]
    res, st, _ = gpu_inflate(cdc, streams, caps=[len(raw) + 600] * len(streams))
    for i, (r, w) in enumerate(zip(res, want)):
        if w is None:
            assert st[i] == s5.ERR.PRESS, i
        elif st[i] == s5.ERR.NOSPACE:
            assert len(w) > len(raw) + 600
        else:
            assert st[i] == 0 and r == w, i


def test_slot_overflow_reports_needed_size(cdc):
    raw = bytes(50_000) + b"tail"
    z = zlib.compress(raw)
    res, st, lh = gpu_inflate(cdc, [z, z], caps=[100, len(raw)])
    assert st.tolist() == [s5.ERR.NOSPACE, 0] and lh[0] == len(raw) and res[1] == raw


def test_pointer_array_form_with_retry(cdc):
    """s5b_depress_batch_host(ZLIB): slots are guessed (4x + 1 KiB); highly compressible streams overflow the
    guess and are transparently decoded again with the exact size."""
    items = [b"", b"abc", bytes(300_000), corpus()["text"], synth.nanopore_signal(4096, seed=2).numpy().tobytes()]
    zs = [zlib.compress(x) for x in items]
    rc, out = cdc.depress_batch(METHOD.ZLIB, zs + [zs[3][:40] + b"\xff\xff" + zs[3][42:]])
    assert out[:5] == items
    assert rc == s5.ERR.PRESS and out[5] is None
    assert codec.ptr_depress_solo(METHOD.ZLIB, zs[4]) == items[4]
