"""The CUDA zstd frame encoder (zstd_encode_kernel, through the C-ABI), the twin of ptr_compress_zstd
(slow5_press.c:1183-1202).  The reference does not pin compressed zstd bytes (its encode goldens are commented out,
test/test_view.sh:204-214), so the contract is the one its decoder imposes (slow5_press.c:1205-1230): system libzstd
regenerates exactly the input from every frame, ZSTD_getFrameContentSize knows the size, the reference binary reads
files written here, and the size stays within the stated tolerance of libzstd level 1 on BLOW5-like records."""
import filecmp
import os
import subprocess

import numpy as np
import pytest
import torch

import slow5tools_b200 as s5
from slow5tools_b200 import codec, synth
from slow5tools_b200._capi import METHOD
from test_zstd_gpu import LibZstd
from test_deflate_gpu import svb_records

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")
RATIO_TOLERANCE = 1.03      # compressed bytes <= 1.03 x libzstd level 1 (SLOW5_ZSTD_COMPRESS_LEVEL) on svb-zd records
RATIO_TOLERANCE_RAW = 1.10  # same for records holding the raw int16 signal (-s none)


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


def gpu_zstd_encode(cdc, bufs, splits=None, misalign=0):
    n = len(bufs)
    lens = np.array([len(b) for b in bufs], np.uint32)
    ioff = np.zeros(n + 1, np.uint64)
    pos = misalign
    for i, b in enumerate(bufs):
        ioff[i] = pos
        pos += len(b) + misalign
    ioff[-1] = pos
    din = np.full((pos + 15) // 16 * 16 + 16, 0x33, np.uint8)
    for b, o in zip(bufs, ioff):
        din[int(o):int(o) + len(b)] = np.frombuffer(b, np.uint8)
    ooff = np.zeros(n + 1, np.uint64)
    pos = misalign
    for i, b in enumerate(bufs):
        ooff[i] = pos
        pos += int(s5.lib.s5b_zstd_bound(len(b))) + misalign
    ooff[-1] = pos
    out = torch.full((pos + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    status = torch.full((n,), 99, dtype=torch.int32, device="cuda")
    split = dev(np.asarray(splits, np.uint32).view(np.int32)) if splits is not None else None
    cdc.zstd_encode_dev(dev(din), dev(ioff.view(np.int64)), dev(lens.view(np.int32)), out, dev(ooff.view(np.int64)),
                        out_len, status, split=split)
    torch.cuda.synchronize()
    oh, lh, st = out.cpu().numpy(), out_len.cpu().numpy(), status.cpu().numpy()
    res = []
    mask = np.ones(oh.size, bool)
    for i in range(n):
        o = int(ooff[i])
        res.append(oh[o:o + lh[i]].tobytes() if st[i] == 0 else None)
        mask[o:o + int(lh[i])] = False
    assert (oh[mask] == 0x5A).all(), "encoder wrote outside its frames"
    return res, st


def check_frames(zs, frames, raws):
    for i, (f, raw) in enumerate(zip(frames, raws)):
        assert f[:4] == b"\x28\xb5\x2f\xfd", i
        assert zs.z.ZSTD_getFrameContentSize(f, len(f)) == len(raw), i   # what ptr_depress_zstd sizes its buffer from
        assert zs.depress(f) == raw, i


def test_libzstd_regenerates_every_input(cdc, zs, oracle):
    rng = np.random.default_rng(3)
    text = b"the quick brown fox jumps over the lazy dog. " * 300
    bufs = [b"", b"a", b"ab", b"aaa", bytes(10), bytes(255), bytes(256), bytes(100_000), text, b"abc" * 5000,
            rng.integers(0, 256, 20_000).astype(np.uint8).tobytes(),          # incompressible -> raw blocks
            rng.integers(0, 4, 50_000).astype(np.uint8).tobytes(),
            rng.integers(0, 2, 300).astype(np.uint8).tobytes(),
            synth.nanopore_signal(30_000, seed=1).numpy().tobytes(),
            bytes(rng.integers(0, 256, 6144).astype(np.uint8)) + bytes(6144) + b"x" * 6143,
            np.clip(rng.exponential(50, 300_000), 0, 255).astype(np.uint8).tobytes()]   # all 256 values, skewed
    bufs += [bytes(rng.integers(0, 200, int(k)).astype(np.uint8)[: int(k)] // 8)
             for k in (15, 16, 17, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 6143, 6144, 6145, 65791, 65792, 65793)]
    recs, splits = svb_records(oracle, 40, 4096, seed=5)
    for misalign in (0, 7):
        res, st = gpu_zstd_encode(cdc, bufs + recs, misalign=misalign)
        assert (st == 0).all()
        check_frames(zs, res, bufs + recs)
    res, st = gpu_zstd_encode(cdc, recs, splits=splits)
    assert (st == 0).all()
    check_frames(zs, res, recs)


def test_skewed_frequencies_hit_the_11_bit_limit(cdc, zs):
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    data = b"".join(bytes([i * 9]) * f for i, f in enumerate(fib))[:6100]
    rng = np.random.default_rng(1)
    data = bytes(rng.permutation(np.frombuffer(data, np.uint8)))
    bufs = [data, data * 3]
    for k in range(300):   # many different weight distributions -> exercises the FSE weight coder
        p = rng.dirichlet(np.full(256, 0.05 + 0.02 * (k % 7)))
        bufs.append(rng.choice(256, size=int(rng.integers(200, 14000)), p=p).astype(np.uint8).tobytes())
    res, st = gpu_zstd_encode(cdc, bufs)
    assert (st == 0).all()
    check_frames(zs, res, bufs)


def test_ratio_within_tolerance_of_libzstd_level1(cdc, zs, oracle):
    recs, splits = svb_records(oracle, 300, 4096, seed=42)
    ref = sum(len(zs.compress(r, 1)) for r in recs)
    res, st = gpu_zstd_encode(cdc, recs, splits=splits)
    assert (st == 0).all()
    ours = sum(len(z) for z in res)
    res1, _ = gpu_zstd_encode(cdc, recs)
    ours1 = sum(len(z) for z in res1)
    raw = sum(len(r) for r in recs)
    print("\nratio: libzstd-1 %.4f  ours(split) %.4f  ours(no hint) %.4f" % (ref / raw, ours / raw, ours1 / raw))
    check_frames(zs, res, recs)
    assert ours <= RATIO_TOLERANCE * ref, (ours, ref)
    # raw int16 signal records (-s none)
    sig = synth.nanopore_signal(100 * 4096, seed=7).numpy().reshape(100, 4096)
    raws = [s.tobytes() for s in sig]
    ref = sum(len(zs.compress(r, 1)) for r in raws)
    res, st = gpu_zstd_encode(cdc, raws)
    ours = sum(len(z) for z in res)
    print("raw int16: libzstd-1 %.4f  ours %.4f" % (ref / (100 * 8192), ours / (100 * 8192)))
    check_frames(zs, res, raws)
    assert ours <= RATIO_TOLERANCE_RAW * ref, (ours, ref)


def test_gpu_decoder_reads_gpu_encoder(cdc, oracle):
    """Our decoder on our encoder's output at a size where every lane-level path of both kernels is exercised,
    long (multi-block, treeless) records included."""
    recs, splits = svb_records(oracle, 2000, 4096, seed=9)
    long_recs, long_splits = svb_records(oracle, 6, 150_000, seed=4)
    res, st = gpu_zstd_encode(cdc, recs + long_recs, splits=splits + long_splits)
    assert (st == 0).all()
    rc, back = cdc.depress_batch(METHOD.ZSTD, res)
    assert rc == 0 and back == recs + long_recs


def test_pointer_array_and_solo_forms(cdc, zs):
    items = [b"", b"hello", bytes(70_000), b"12345\0" * 700]
    rc, fr = cdc.compress_batch(METHOD.ZSTD, items)
    assert rc == 0
    check_frames(zs, fr, items)
    f = codec.ptr_compress_solo(METHOD.ZSTD, b"hello hello hello hello hello")
    assert zs.depress(f) == b"hello hello hello hello hello"
    assert codec.ptr_depress_solo(METHOD.ZSTD, f) == b"hello hello hello hello hello"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built")
def test_reference_binary_reads_gpu_written_zstd_blow5(tmp_path):
    """view -c zstd here, then the unmodified reference decodes the file to the golden SLOW5 (both signal methods,
    generic and device-resident paths)."""
    gold = os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5")
    for src in ("exp_1_lossless_v0.2.0.slow5", "exp_1_lossless_zlib_svb_v0.2.0.blow5"):
        for sig in ("svb-zd", "none"):
            z = tmp_path / ("%s.%s.zstd.blow5" % (src, sig))
            r = subprocess.run([CLI, "view", os.path.join(FIX, src), "-c", "zstd", "-s", sig, "-o", str(z)],
                               stderr=subprocess.PIPE)
            assert r.returncode == 0, r.stderr.decode()
            back = tmp_path / "back.slow5"
            r = subprocess.run([REF, "view", str(z), "-o", str(back)], stderr=subprocess.PIPE)
            assert r.returncode == 0, r.stderr.decode()
            assert filecmp.cmp(back, gold, shallow=False), (src, sig)
            ours = tmp_path / "ours.slow5"
            assert subprocess.run([CLI, "view", str(z), "-o", str(ours)]).returncode == 0
            assert filecmp.cmp(ours, gold, shallow=False), (src, sig)
