"""Parity of the CUDA zstd frame decoder (zstd_decode_kernel, through the C-ABI) with system libzstd -- the
library the reference calls at slow5_press.c:1205-1230 -- on frames written at the reference's level (1) and
every other level, with and without content checksum, corrupted and truncated frames, and on the reference's own
zstd BLOW5 fixtures (the decode direction is what its tests pin: test/test_view.sh:216-229)."""
import ctypes as C
import filecmp
import os
import subprocess

import numpy as np
import pytest
import torch

import slow5tools_b200 as s5
from slow5tools_b200 import codec, synth
from slow5tools_b200._capi import METHOD

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")


class LibZstd:
    def __init__(self):
        self.z = z = C.CDLL("libzstd.so.1")
        z.ZSTD_compressBound.restype = C.c_size_t
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compress.restype = C.c_size_t
        z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        z.ZSTD_decompress.restype = C.c_size_t
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        z.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        z.ZSTD_createCCtx.restype = C.c_void_p
        z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        z.ZSTD_compress2.restype = C.c_size_t
        z.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_freeCCtx.argtypes = [C.c_void_p]

    def compress(self, data, level=1, checksum=False):
        cap = self.z.ZSTD_compressBound(len(data))
        out = C.create_string_buffer(cap)
        if checksum:
            c = self.z.ZSTD_createCCtx()
            self.z.ZSTD_CCtx_setParameter(c, 100, level)
            self.z.ZSTD_CCtx_setParameter(c, 201, 1)
            n = self.z.ZSTD_compress2(c, out, cap, data, len(data))
            self.z.ZSTD_freeCCtx(c)
        else:
            n = self.z.ZSTD_compress(out, cap, data, len(data), level)
        assert not self.z.ZSTD_isError(n)
        return out.raw[:n]

    def depress(self, frame):
        """What ptr_depress_zstd returns: bytes or None."""
        size = self.z.ZSTD_getFrameContentSize(frame, len(frame))
        if size >= 2**64 - 2:
            return None
        out = C.create_string_buffer(max(size, 1))
        n = self.z.ZSTD_decompress(out, size, frame, len(frame))
        if self.z.ZSTD_isError(n):
            return None
        return out.raw[:n]


@pytest.fixture(scope="module")
def zs():
    try:
        return LibZstd()
    except OSError:
        pytest.skip("libzstd.so.1 not available")


@pytest.fixture(scope="module")
def cdc():
    c = s5.Codec(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_zstd(cdc, frames, caps, misalign=0):
    n = len(frames)
    lens = np.array([len(f) for f in frames], np.uint32)
    ioff = np.zeros(n + 1, np.uint64)
    pos = misalign
    for i, f in enumerate(frames):
        ioff[i] = pos
        pos += len(f) + misalign
    ioff[-1] = pos
    zin = np.full((pos + 15) // 16 * 16 + 16, 0xEE, np.uint8)
    for f, o in zip(frames, ioff):
        zin[int(o):int(o) + len(f)] = np.frombuffer(f, np.uint8)
    ooff = np.zeros(n + 1, np.uint64)
    pos = misalign
    for i, c in enumerate(caps):
        ooff[i] = pos
        pos += c + misalign
    ooff[-1] = pos
    out = torch.full((pos + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    status = torch.full((n,), 99, dtype=torch.int32, device="cuda")
    cdc.zstd_decode_dev(dev(zin), dev(ioff.view(np.int64)), dev(lens.view(np.int32)), out, dev(ooff.view(np.int64)), out_len, status)
    torch.cuda.synchronize()
    oh, lh, st = out.cpu().numpy(), out_len.cpu().numpy(), status.cpu().numpy()
    res = []
    mask = np.ones(oh.size, bool)
    for i in range(n):
        o = int(ooff[i])
        res.append(oh[o:o + lh[i]].tobytes() if st[i] == 0 else None)
        mask[o:int(ooff[i + 1])] = False
    assert (oh[mask] == 0x5A).all(), "decoder wrote outside its slots"
    return res, st, lh


def corpus():
    rng = np.random.default_rng(21)
    sig = synth.nanopore_signal(40000, seed=3).numpy()
    text = (b"the quick brown fox jumps over the lazy dog. " * 300) + bytes(rng.integers(97, 123, 3000).astype(np.uint8))
    svb = bytes(1030) + np.clip(rng.normal(9, 6, 4200), 0, 255).astype(np.uint8).tobytes()
    return {"empty": b"", "one": b"a", "zeros": bytes(100_000), "text": text, "svb_like": svb,
            "random": rng.integers(0, 256, 50_000).astype(np.uint8).tobytes(), "raw_signal": sig.tobytes(),
            "period": b"abcdefg" * 10_000, "big_mixed": (svb + text) * 12,
            "far": rng.integers(0, 256, 70_000).astype(np.uint8).tobytes() * 2}


def test_frames_at_every_level(cdc, zs):
    frames, raws, names = [], [], []
    for name, raw in corpus().items():
        for level in (1, 2, 3, 6, 12, 19):
            for cs in (False, True):
                frames.append(zs.compress(raw, level, cs))
                raws.append(raw)
                names.append((name, level, cs))
    for misalign in (0, 3):
        res, st, _ = gpu_zstd(cdc, frames, [len(r) for r in raws], misalign=misalign)
        assert (st == 0).all(), [n for n, s in zip(names, st) if s != 0]
        for n, r, w in zip(names, res, raws):
            assert r == w, n


def test_svb_records_at_the_reference_level(cdc, zs, oracle):
    sig = synth.nanopore_signal(3000 * 4096, seed=5).numpy().reshape(3000, 4096)
    raws = [oracle.compress(s) for s in sig]
    frames = [zs.compress(r, 1) for r in raws]        # SLOW5_ZSTD_COMPRESS_LEVEL = 1 (slow5_press.h:58)
    res, st, _ = gpu_zstd(cdc, frames, [len(r) for r in raws])
    assert (st == 0).all() and res == raws
    rc, out = cdc.depress_batch(METHOD.ZSTD, frames[:50] + [frames[0][:-3]])
    assert rc == s5.ERR.PRESS and out[:50] == raws[:50] and out[50] is None
    assert codec.ptr_depress_solo(METHOD.ZSTD, frames[7]) == raws[7]


def test_corrupted_and_truncated_frames_match_libzstd(cdc, zs):
    """Verdict parity with libzstd.  Frames carrying a content checksum: identical verdicts and bytes.  Frames
    without one: we never accept what libzstd rejects and agree byte for byte whenever both accept; this libzstd
    build (1.5.5, BMI2 fast Huffman loop) skips the end-of-bitstream checks RFC 8878 asks for and decodes some
    damaged literal / sequence streams to garbage, which this decoder rejects instead (zstd)."""
    rng = np.random.default_rng(9)
    raw = corpus()["svb_like"] + corpus()["text"][:3000]
    for checksum, base in ((True, [zs.compress(raw, 1, True), zs.compress(raw, 19, True)]),
                           (False, [zs.compress(raw, 1), zs.compress(raw, 19)])):
        frames = []
        for f in base:
            for _ in range(200):
                b = bytearray(f)
                k = int(rng.integers(0, len(b)))
                b[k] ^= 1 << int(rng.integers(0, 8))
                frames.append(bytes(b))
            frames += [f[:k] for k in (0, 1, 4, 5, 6, 9, 20, len(f) // 2, len(f) - 1)]
            frames.append(f + b"\0")
        want = [zs.depress(f) for f in frames]
        res, st, _ = gpu_zstd(cdc, frames, [len(raw) + 64] * len(frames))
        stricter = 0
        for i, (r, w) in enumerate(zip(res, want)):
            if w is None:
                assert st[i] != 0, (checksum, i)
            elif st[i] == 0:
                assert r == w, (checksum, i)
            else:
                stricter += 1
        if checksum:
            assert stricter == 0
        else:
            assert stricter < len(frames) // 4


def test_slot_too_small_reports_content_size(cdc, zs):
    raw = bytes(10_000)
    f = zs.compress(raw)
    res, st, lh = gpu_zstd(cdc, [f, f], [100, len(raw)])
    assert st.tolist() == [s5.ERR.NOSPACE, 0] and lh[0] == len(raw) and res[1] == raw


def test_view_reads_the_reference_zstd_fixtures(tmp_path):
    """test/test_view.sh:216-229: zstd(+svb-zd) BLOW5 -> SLOW5 must equal the golden text."""
    for name in ("exp_1_lossless_zstd_svb_v0.2.0.blow5", "exp_1_lossless_zstd_v0.2.0.blow5"):
        out = tmp_path / (name + ".slow5")
        r = subprocess.run([CLI, "view", os.path.join(FIX, name), "-o", str(out)], stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        assert filecmp.cmp(out, os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5"), shallow=False), name
        # and through the device-resident blow5 -> blow5 path, re-compressed as zlib + svb-zd
        z = tmp_path / (name + ".z.blow5")
        assert subprocess.run([CLI, "view", os.path.join(FIX, name), "-o", str(z)]).returncode == 0
        back = tmp_path / (name + ".back.slow5")
        assert subprocess.run([CLI, "view", str(z), "-o", str(back)]).returncode == 0
        assert filecmp.cmp(back, os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5"), shallow=False), name
