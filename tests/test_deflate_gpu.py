"""The CUDA zlib encoder (deflate_kernel, through the C-ABI): every stream it emits must inflate with
system zlib -- the reference's decoder (slow5_press.c:973-1010) -- to exactly the input, Adler-32 included,
and its size must stay within the stated tolerance of zlib level 6 on BLOW5-like records."""
import zlib

import numpy as np
import pytest
import torch

import slow5tools_b200 as s5
from slow5tools_b200 import codec, synth
from slow5tools_b200._capi import METHOD

pytestmark = pytest.mark.gpu
RATIO_TOLERANCE = 1.03   # compressed bytes <= 1.03 x zlib level 6, summed over the calibrated synthetic records


@pytest.fixture(scope="module")
def cdc():
    c = s5.Codec(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_deflate(cdc, bufs, splits=None, misalign=0):
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
        pos += int(s5.lib.s5b_zlib_bound(len(b))) + misalign
    ooff[-1] = pos
    out = torch.full((pos + 64,), 0x5A, dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    status = torch.full((n,), 99, dtype=torch.int32, device="cuda")
    split = dev(np.asarray(splits, np.uint32).view(np.int32)) if splits is not None else None
    cdc.zlib_deflate_dev(dev(din), dev(ioff.view(np.int64)), dev(lens.view(np.int32)), out, dev(ooff.view(np.int64)),
                         out_len, status, split=split)
    torch.cuda.synchronize()
    oh, lh, st = out.cpu().numpy(), out_len.cpu().numpy(), status.cpu().numpy()
    res = []
    mask = np.ones(oh.size, bool)
    for i in range(n):
        o = int(ooff[i])
        res.append(oh[o:o + lh[i]].tobytes() if st[i] == 0 else None)
        mask[o:o + int(lh[i])] = False
    assert (oh[mask] == 0x5A).all(), "deflate wrote outside its streams"
    return res, st


def svb_records(oracle, n_reads, n_samples, seed):
    """BLOW5-like packed records: ~70 bytes of fixed fields + the svb-zd stream (slow5.c:3928-4074)."""
    sig = synth.nanopore_signal(n_reads * n_samples, seed=seed).numpy().reshape(n_reads, n_samples)
    recs, splits = [], []
    rng = np.random.default_rng(seed)
    for r in range(n_reads):
        rid = ("%08x-%04x-%04x-%04x-%012x" % tuple(int(v) for v in rng.integers(0, 2**16, 5))).encode()
        svb = oracle.compress(sig[r])
        head = (len(rid).to_bytes(2, "little") + rid + (0).to_bytes(4, "little")
                + np.array([8192.0, 9.0, 1444.86, 4000.0]).tobytes() + len(svb).to_bytes(8, "little"))
        recs.append(head + svb)
        splits.append(len(head) + 4 + (n_samples + 3) // 4)
    return recs, splits


def test_roundtrip_through_system_zlib(cdc, oracle):
    rng = np.random.default_rng(3)
    text = b"the quick brown fox jumps over the lazy dog. " * 300
    bufs = [b"", b"a", b"ab", b"aaa", bytes(10), bytes(100_000), text, b"abc" * 5000,
            rng.integers(0, 256, 20_000).astype(np.uint8).tobytes(),          # incompressible -> stored blocks
            rng.integers(0, 4, 50_000).astype(np.uint8).tobytes(),
            synth.nanopore_signal(30_000, seed=1).numpy().tobytes(),
            bytes(rng.integers(0, 256, 6144).astype(np.uint8)) + bytes(6144) + b"x" * 6143]
    bufs += [bytes(rng.integers(0, 256, int(k)).astype(np.uint8)[: int(k)]) for k in (31, 32, 33, 6143, 6144, 6145)]
    recs, splits = svb_records(oracle, 40, 4096, seed=5)
    for misalign in (0, 7):
        res, st = gpu_deflate(cdc, bufs + recs, misalign=misalign)
        assert (st == 0).all()
        for i, (z, raw) in enumerate(zip(res, bufs + recs)):
            assert z[:2] == b"\x78\x9c"
            d = zlib.decompressobj(15)
            assert d.decompress(z) == raw and d.eof and d.unused_data == b"", i      # complete stream, Adler-32 verified


def test_skewed_frequencies_hit_the_length_limit(cdc):
    # Fibonacci-like symbol frequencies force code lengths > 15 before limiting
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    data = b"".join(bytes([i]) * f for i, f in enumerate(fib))[:6100]
    rng = np.random.default_rng(1)
    data = bytes(rng.permutation(np.frombuffer(data, np.uint8)))
    res, st = gpu_deflate(cdc, [data, data * 3])
    assert (st == 0).all()
    assert zlib.decompress(res[0]) == data and zlib.decompress(res[1]) == data * 3


def test_code_length_alphabet_hits_its_7_bit_limit(cdc):
    """The 19-symbol code-length code is limited to 7 bits; skewed header statistics must still give a complete
    code (regression: 1 stream in 100k was rejected by zlib with 'invalid code lengths set')."""
    rng = np.random.default_rng(17)
    bufs = []
    for k in range(400):
        # byte distributions that give literal code lengths spread over many different values
        p = rng.dirichlet(np.full(256, 0.05 + 0.02 * (k % 7)))
        bufs.append(rng.choice(256, size=int(rng.integers(200, 6000)), p=p).astype(np.uint8).tobytes())
    res, st = gpu_deflate(cdc, bufs)
    assert (st == 0).all()
    for z, raw in zip(res, bufs):
        assert zlib.decompress(z) == raw


def test_ratio_within_tolerance_of_zlib_level6(cdc, oracle):
    recs, splits = svb_records(oracle, 300, 4096, seed=42)
    ref = sum(len(zlib.compress(r, 6)) for r in recs)
    res, st = gpu_deflate(cdc, recs, splits=splits)
    assert (st == 0).all()
    ours = sum(len(z) for z in res)
    res1, _ = gpu_deflate(cdc, recs)                       # without the split hint
    ours1 = sum(len(z) for z in res1)
    raw = sum(len(r) for r in recs)
    print("\nratio: zlib-6 %.4f  ours(split) %.4f  ours(single block) %.4f" % (ref / raw, ours / raw, ours1 / raw))
    assert all(zlib.decompress(z) == r for z, r in zip(res, recs))
    assert ours <= RATIO_TOLERANCE * ref, (ours, ref)


def test_gpu_inflate_reads_gpu_deflate(cdc, oracle):
    """Our decoder on our encoder's output (the view recode path), at a size where every lane-level path of both
    kernels is exercised."""
    recs, splits = svb_records(oracle, 2000, 4096, seed=9)
    res, st = gpu_deflate(cdc, recs, splits=splits)
    assert (st == 0).all()
    rc, back = cdc.depress_batch(METHOD.ZLIB, res)
    assert rc == 0 and back == recs


def test_pointer_array_and_solo_forms(cdc):
    items = [b"", b"hello", bytes(70_000), b"12345\0"]
    rc, zs = cdc.compress_batch(METHOD.ZLIB, items)
    assert rc == 0 and [zlib.decompress(z) for z in zs] == items
    z = codec.ptr_compress_solo(METHOD.ZLIB, b"hello")
    assert zlib.decompress(z) == b"hello"


def test_header_kernel_and_in_warp_headers_give_the_same_streams(tmp_path):
    """deflate_header_kernel (one thread per block) took over what the emit kernel's warp did per block; the older path is still
    there behind S5B_DEFLATE_HDR=warp (read once per process).  Both must produce the same bytes, with and without the canned
    code for the front part of the records."""
    import os
    import subprocess
    import sys
    script = tmp_path / "one.py"
    script.write_text(
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "sys.path.insert(0, %r)\n"
        "import torch, slow5tools_b200 as s5\n"
        "from slow5tools_b200 import synth\n"
        "import test_deflate_gpu as t\n"
        "from conftest import Oracle, build_oracle\n"
        "cdc = s5.Codec(0)\n"
        "bufs, splits = t.svb_records(Oracle(build_oracle()), 300, 4096, 5)\n"
        "rng = np.random.default_rng(3)\n"
        "bufs += [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (0, 1, 2, 70, 5000, 20000)]\n"
        "bufs += [bytes(20000), b'ab' * 9000, bytes(rng.integers(0, 3, 30000, dtype=np.uint8))]\n"
        "splits = list(splits) + [0, 0, 1, 30, 100, 7000, 1000, 0, 12000]\n"
        "out, st = t.gpu_deflate(cdc, bufs, splits)\n"
        "assert (st == 0).all()\n"
        "h = hashlib.sha256()\n"
        "for o in out: h.update(o)\n"
        "print('DIGEST', h.hexdigest(), sum(len(o) for o in out))\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                      os.path.dirname(os.path.abspath(__file__))))
    got = {}
    for hdr in ("thread", "warp"):
        for canned in ("1", "0"):
            env = dict(os.environ, S5B_DEFLATE_HDR=hdr, S5B_DEFLATE_CANNED=canned)
            r = subprocess.run([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
            assert r.returncode == 0, r.stderr.decode()[-2000:]
            got[(hdr, canned)] = [l for l in r.stdout.decode().splitlines() if l.startswith("DIGEST")][0]
    assert got[("thread", "1")] == got[("warp", "1")]
    assert got[("thread", "0")] == got[("warp", "0")]
    assert got[("thread", "1")] != got[("thread", "0")]
