"""`slow5tools-b200 view` with the codec on the GPU against the reference binary and its goldens
(test/test_view.sh:90-200): SLOW5 text and uncompressed BLOW5 byte-identical; GPU-compressed BLOW5 must be
read back by the REFERENCE to the identical SLOW5 and stay within the size tolerance."""
import ctypes as C
import filecmp
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")
have_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/slow5tools_ref not present")


def ours(*args, env=None):
    r = subprocess.run([CLI, "view"] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=dict(os.environ, **env) if env else None)
    assert r.returncode == 0, r.stderr.decode()
    return r


def ref(*args):
    subprocess.check_call([REF, "view"] + list(args), stderr=subprocess.DEVNULL)


def test_zlib_svb_to_slow5_matches_reference_golden(tmp_path):
    out = tmp_path / "a.slow5"
    ours(os.path.join(FIX, "exp_1_lossless_zlib_svb_v0.2.0.blow5"), "-o", str(out))
    assert filecmp.cmp(out, os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5"), shallow=False)   # test_view.sh golden
    for name in ("zlib_svb-zd_multi_rg_v0.2.0", "exp_1_lossy_zlib"):
        out = tmp_path / (name + ".slow5")
        ours(os.path.join(FIX, name + ".blow5"), "-o", str(out), "-t", "4", "-K", "3")
        assert filecmp.cmp(out, os.path.join(FIX, name + ".expected.slow5"), shallow=False), name


@have_ref
def test_to_uncompressed_blow5_byte_identical(tmp_path):
    for name in ("exp_1_lossless_zlib_svb_v0.2.0", "zlib_svb-zd_multi_rg_v0.2.0", "zlib_svb-zd_v0.2.0"):
        a, b = tmp_path / "ours.blow5", tmp_path / "ref.blow5"
        ours(os.path.join(FIX, name + ".blow5"), "-o", str(a), "-c", "none", "-s", "none")
        ref(os.path.join(FIX, name + ".blow5"), "-o", str(b), "-c", "none", "-s", "none")
        assert filecmp.cmp(a, b, shallow=False), name
        # svb-zd only: bit-exact signal codec, no zlib involved -> byte-identical file
        ours(str(b), "-o", str(a), "-c", "none", "-s", "svb-zd")
        c = tmp_path / "ref_svb.blow5"
        ref(str(b), "-o", str(c), "-c", "none", "-s", "svb-zd")
        assert filecmp.cmp(a, c, shallow=False), name


@have_ref
@pytest.mark.parametrize("flags", [[], ["-c", "zlib", "-s", "none"], ["-c", "zlib", "-s", "svb-zd", "-K", "2", "-t", "2"]])
def test_gpu_compressed_files_are_read_by_the_reference(tmp_path, flags):
    for name in ("exp_1_lossless_zlib_svb_v0.2.0", "zlib_svb-zd_multi_rg_v0.2.0"):
        src = os.path.join(FIX, name + ".blow5")
        mine, theirs = tmp_path / "mine.blow5", tmp_path / "theirs.blow5"
        ours(src, "-o", str(mine), *flags)
        ref(src, "-o", str(theirs), *[f for f in flags if f not in ("-K", "2", "-t")])
        back, want = tmp_path / "back.slow5", tmp_path / "want.slow5"
        ref(str(mine), "-o", str(back))            # the REFERENCE decodes the GPU-written file
        ref(src, "-o", str(want))
        assert filecmp.cmp(back, want, shallow=False), (name, flags)
        # size tolerance: +3 % of zlib-6 on svb-zd records, the default and north-star path.  Raw
        # int16 records (-s none, the pre-0.2.0 layout) are where zlib's LZ77 matching pays and a Huffman + run
        # encoder does not: a known, stated gap (+15 % bound) -- the output is still a valid file the reference reads.
        tol = 1.15 if "none" in flags else 1.03
        assert os.path.getsize(mine) <= tol * os.path.getsize(theirs), (os.path.getsize(mine), os.path.getsize(theirs))
        again = tmp_path / "again.slow5"
        ours(str(mine), "-o", str(again))          # and so does our own reader
        assert filecmp.cmp(again, want, shallow=False)


def test_slow5_input_to_compressed_blow5_roundtrip(tmp_path):
    src = os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5")
    z, back = tmp_path / "z.blow5", tmp_path / "back.slow5"
    ours(src, "-o", str(z))
    ours(str(z), "-o", str(back))
    assert filecmp.cmp(back, src, shallow=False)


def test_low_level_file_api_roundtrip(tmp_path):
    """slow5_open / slow5_get_next_mem / slow5_decode / slow5_encode / slow5_write_bytes twins."""
    L = C.CDLL(os.path.join(ROOT, "slow5tools_b200", "libslow5b200.so"))
    vp, sz = C.c_void_p, C.c_size_t
    L.s5b_open.restype = vp
    L.s5b_open.argtypes = [C.c_char_p, C.c_char_p]
    L.s5b_get_next_mem.restype = vp
    L.s5b_get_next_mem.argtypes = [C.POINTER(sz), vp]
    L.s5b_decode.argtypes = [C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), vp]
    L.s5b_encode.argtypes = [C.POINTER(vp), C.POINTER(sz), vp, vp]
    L.s5b_write_bytes.argtypes = [vp, sz, vp]
    for f in (L.s5b_close, L.s5b_hdr_write, L.s5b_rec_free):
        f.argtypes = [vp]
    L.s5b_hdr_copy.argtypes = [vp, vp]
    L.s5b_set_press.argtypes = [vp, C.c_int, C.c_int]
    libc = C.CDLL(None)
    libc.free.argtypes = [vp]
    src = os.path.join(FIX, "zlib_svb-zd_multi_rg_v0.2.0.blow5")
    out = tmp_path / "api.blow5"
    fin = L.s5b_open(src.encode(), b"r")
    fout = L.s5b_open(str(out).encode(), b"w")
    assert fin and fout
    assert L.s5b_hdr_copy(fout, fin) == 0 and L.s5b_set_press(fout, 1, 2) == 0 and L.s5b_hdr_write(fout) > 0
    n_rec = 0
    while True:
        n = sz()
        mem = vp(L.s5b_get_next_mem(C.byref(n), fin))
        if not mem:
            assert L.s5b_errno_value() == -1      # SLOW5_ERR_EOF
            break
        rec = vp()
        assert L.s5b_decode(C.byref(mem), C.byref(n), C.byref(rec), fin) == 0
        libc.free(mem)
        enc, en = vp(), sz()
        assert L.s5b_encode(C.byref(enc), C.byref(en), rec, fout) == 0
        assert L.s5b_write_bytes(enc, en, fout) == 0   # like slow5_write_bytes (slow5.c:3785-3794)
        libc.free(enc)
        L.s5b_rec_free(rec)
        n_rec += 1
    assert n_rec == 7
    assert L.s5b_close(fin) == 0 and L.s5b_close(fout) == 0
    back = tmp_path / "api.slow5"
    ours(str(out), "-o", str(back))
    assert filecmp.cmp(back, os.path.join(FIX, "zlib_svb-zd_multi_rg_v0.2.0.expected.slow5"), shallow=False)


# ---- ex-zd signal compression (test/test_view.sh:102-163)
def test_exzd_golden_to_slow5_matches_reference_golden(tmp_path):
    out = tmp_path / "a.slow5"
    ours(os.path.join(FIX, "exp_1_lossless_zlib_ex_zd.blow5"), "-o", str(out))        # test_view.sh:158-160
    assert filecmp.cmp(out, os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5"), shallow=False)


@have_ref
def test_exzd_files_byte_identical_without_zlib_and_read_by_the_reference(tmp_path):
    src = os.path.join(FIX, "exp_1_lossless_zlib_svb_v0.2.0.blow5")
    # ex-zd is deterministic: with uncompressed records our file must equal the reference's byte for byte
    a, b = tmp_path / "ours.blow5", tmp_path / "ref.blow5"
    ours(src, "-o", str(a), "-c", "none", "-s", "ex-zd")
    ref(src, "-o", str(b), "-c", "none", "-s", "ex-zd")
    assert filecmp.cmp(a, b, shallow=False)
    # zlib records around ex-zd signals: the reference reads ours back to the golden text, and we read the reference's
    mine = tmp_path / "mine.blow5"
    ours(os.path.join(FIX, "exp_1_lossless.slow5"), "-o", str(mine), "-s", "ex-zd")   # test_view.sh:102-104 (bytes: zlib's)
    t1, t2 = tmp_path / "t1.slow5", tmp_path / "t2.slow5"
    ref(str(mine), "-o", str(t1))
    assert filecmp.cmp(t1, os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5"), shallow=False)
    ours(str(b), "-o", str(t2))
    assert filecmp.cmp(t2, os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5"), shallow=False)
    # ex-zd -> ex-zd and ex-zd -> svb-zd re-encodes (test_view.sh:161-163)
    c = tmp_path / "c.blow5"
    ours(os.path.join(FIX, "exp_1_lossless_zlib_ex_zd.blow5"), "-o", str(c), "-c", "none", "-s", "ex-zd")
    assert filecmp.cmp(c, b, shallow=False)
    d, e = tmp_path / "d.blow5", tmp_path / "e.blow5"
    ours(os.path.join(FIX, "exp_1_lossless_zlib_ex_zd.blow5"), "-o", str(d), "-c", "none", "-s", "svb-zd")
    ref(os.path.join(FIX, "exp_1_lossless_zlib_ex_zd.blow5"), "-o", str(e), "-c", "none", "-s", "svb-zd")
    assert filecmp.cmp(d, e, shallow=False)


@have_ref
def test_exzd_fast_and_general_paths_agree(tmp_path):
    """blow5 -> blow5 with ex-zd on either side: the device-resident path and the host parse/pack path write the same file."""
    src = os.path.join(FIX, "exp_1_lossless_zlib_svb_v0.2.0.blow5")
    exz = os.path.join(FIX, "exp_1_lossless_zlib_ex_zd.blow5")
    for inp, flags in ((src, ["-c", "none", "-s", "ex-zd"]), (exz, ["-c", "none", "-s", "svb-zd"]),
                       (exz, ["-c", "none", "-s", "none"]), (src, ["-c", "zstd", "-s", "ex-zd"]), (exz, ["-c", "zstd", "-s", "svb-zd"])):
        a, b = tmp_path / "fast.blow5", tmp_path / "slow.blow5"
        ours(inp, "-o", str(a), *flags)
        ours(inp, "-o", str(b), *flags, env={"S5B_VIEW_SLOW_PATH": "1"})
        assert filecmp.cmp(a, b, shallow=False), flags
        t = tmp_path / "t.slow5"
        ref(str(a), "-o", str(t))
        assert filecmp.cmp(t, os.path.join(FIX, "exp_1_lossless_v0.2.0.slow5"), shallow=False), flags


def test_mt_batch_api_roundtrip(tmp_path):
    """slow5_mt.h twins: slow5_init_mt / slow5_init_batch / slow5_get_next_batch / slow5_write_batch (pyslow5's path)."""
    L = C.CDLL(os.path.join(ROOT, "slow5tools_b200", "libslow5b200.so"))
    vp = C.c_void_p

    class Batch(C.Structure):  # slow5_batch_t, slow5_mt.h:23-33
        _fields_ = [("n_rec", C.c_int32), ("capacity_rec", C.c_int32), ("mem_records", C.POINTER(vp)),
                    ("mem_bytes", C.POINTER(C.c_size_t)), ("slow5_rec", C.POINTER(vp)), ("rid", C.POINTER(vp))]

    L.s5b_open.restype = vp
    L.s5b_open.argtypes = [C.c_char_p, C.c_char_p]
    L.s5b_init_mt.restype = vp
    L.s5b_init_mt.argtypes = [C.c_int, vp]
    L.s5b_init_batch.restype = C.POINTER(Batch)
    L.s5b_init_batch.argtypes = [C.c_int]
    for f in (L.s5b_get_next_batch, L.s5b_write_batch):
        f.argtypes = [vp, C.POINTER(Batch), C.c_int]
    L.s5b_free_batch.argtypes = [C.POINTER(Batch)]
    for f in (L.s5b_close, L.s5b_hdr_write, L.s5b_free_mt):
        f.argtypes = [vp]
    L.s5b_hdr_copy.argtypes = [vp, vp]
    L.s5b_set_press.argtypes = [vp, C.c_int, C.c_int]
    src = os.path.join(FIX, "zlib_svb-zd_multi_rg_v0.2.0.blow5")
    for rec_press, sig_press in ((1, 2), (3, 4), (0, 0)):      # zlib+svb-zd, zstd+ex-zd, none/none
        out = tmp_path / ("mt_%d_%d.blow5" % (rec_press, sig_press))
        fin, fout = L.s5b_open(src.encode(), b"r"), L.s5b_open(str(out).encode(), b"w")
        assert fin and fout
        assert L.s5b_hdr_copy(fout, fin) == 0 and L.s5b_set_press(fout, rec_press, sig_press) == 0 and L.s5b_hdr_write(fout) > 0
        mt_in, mt_out = L.s5b_init_mt(4, fin), L.s5b_init_mt(4, fout)
        batch = L.s5b_init_batch(3)
        total, sizes = 0, []
        while True:
            n = L.s5b_get_next_batch(mt_in, batch, 3)
            assert n >= 0
            if n == 0:
                break
            assert batch.contents.n_rec == n
            assert L.s5b_write_batch(mt_out, batch, n) == n
            total += n
            sizes.append(n)
            if n < 3:
                break
        assert total == 7 and sizes == [3, 3, 1]
        L.s5b_free_batch(batch)
        L.s5b_free_mt(mt_in)
        L.s5b_free_mt(mt_out)
        assert L.s5b_close(fin) == 0 and L.s5b_close(fout) == 0
        back = tmp_path / "mt.slow5"
        ours(str(out), "-o", str(back))
        assert filecmp.cmp(back, os.path.join(FIX, "zlib_svb-zd_multi_rg_v0.2.0.expected.slow5"), shallow=False)
        if os.path.exists(REF):
            back2 = tmp_path / "mt_ref.slow5"
            ref(str(out), "-o", str(back2))
            assert filecmp.cmp(back2, os.path.join(FIX, "zlib_svb-zd_multi_rg_v0.2.0.expected.slow5"), shallow=False)


def test_fast_path_small_last_chunk_lands_after_the_big_ones(tmp_path):
    """Regression: the chunk writer mixes parallel pwrite()s (chunks > 8 MiB) with a sequential write for small chunks;
    a small LAST chunk used to be written at the descriptor's stale offset, over the first records of the file."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_view
    from slow5tools_b200 import synth
    R, N = 5600, 4096            # 46 MB of records: one 40 MB input chunk (18 MB out) + a 4 MB tail (1.8 MB out)
    raw = tmp_path / "raw.blow5"
    bench_view.write_blow5(str(raw), synth.nanopore_signal(R * N, seed=5).numpy(), R, N)
    for flags in (["-c", "zlib", "-s", "svb-zd"], ["-c", "zstd", "-s", "ex-zd"]):
        z, back = tmp_path / "z.blow5", tmp_path / "back.blow5"
        ours(str(raw), "-o", str(z), *flags)
        ours(str(z), "-o", str(back), "-c", "none", "-s", "none")
        assert filecmp.cmp(raw, back, shallow=False), flags
        slow = tmp_path / "slow.blow5"
        ours(str(raw), "-o", str(slow), *flags, env={"S5B_VIEW_SLOW_PATH": "1"})
        assert filecmp.cmp(z, slow, shallow=False), flags


@have_ref
def test_record_larger_than_a_chunk(tmp_path):
    """blow5 -> blow5 fast path with chunks far smaller than one stored record (S5B_VIEW_CHUNK_KB test hook): the reader
    must grow its buffer from the record's own size prefix instead of carrying the same bytes forever (round-1 advisor
    finding: a record >= the 40 MiB chunk made `view` spin)."""
    src = os.path.join(FIX, "exp_1_lossless_zlib_svb_v0.2.0.blow5")
    raw, small, normal = tmp_path / "raw.blow5", tmp_path / "small.blow5", tmp_path / "normal.blow5"
    ref(src, "-o", str(raw), "-c", "none", "-s", "none")       # uncompressed records: ~70 KB each
    ours(str(raw), "-o", str(normal), "-c", "none", "-s", "svb-zd")
    r = subprocess.run([CLI, "view", str(raw), "-o", str(small), "-c", "none", "-s", "svb-zd"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=dict(os.environ, S5B_VIEW_CHUNK_KB="4"), timeout=120)
    assert r.returncode == 0, r.stderr.decode()
    assert filecmp.cmp(small, normal, shallow=False)


def test_corrupt_size_prefix_fails_cleanly(tmp_path):
    """a record size prefix of 2^39 must give the reference's clean failure (exit 1), not an uncaught bad_alloc (rc 134)"""
    src = os.path.join(FIX, "exp_1_lossless_zlib_svb_v0.2.0.blow5")
    data = bytearray(open(src, "rb").read())
    hsize = int.from_bytes(data[64:68], "little")
    at = 68 + hsize
    data[at:at + 8] = (1 << 39).to_bytes(8, "little")
    bad = tmp_path / "bad.blow5"
    bad.write_bytes(bytes(data))
    for extra in ([], ["-K", "7"]):   # fast path and the -K slow path
        r = subprocess.run([CLI, "view", str(bad), "-o", str(tmp_path / "o.slow5")] + extra, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, timeout=60)
        assert r.returncode == 1, (r.returncode, r.stderr.decode())


@have_ref
def test_text_paths_through_the_gpu_formatter_and_parser(tmp_path):
    """blow5 -> slow5 formats the raw_signal column on the GPU and slow5 -> blow5 parses it there (ascii_kernels.cu): both
    must be byte-identical to the reference binary, for svb-zd, ex-zd and uncompressed signals under zlib"""
    for name, flags in (("exp_1_lossless_zlib_svb_v0.2.0", []), ("zlib_svb-zd_multi_rg_v0.2.0", ["-t", "3", "-K", "2"]),
                        ("exp_1_lossless_zlib_ex_zd", [])):
        src = os.path.join(FIX, name + ".blow5")
        if not os.path.exists(src):
            continue
        mine, theirs = tmp_path / "m.slow5", tmp_path / "t.slow5"
        ours(src, "-o", str(mine), *flags)
        ref(src, "-o", str(theirs))
        assert filecmp.cmp(mine, theirs, shallow=False), name
        # and back: text in, compressed blow5 out; the reference reads it to the same text
        back = tmp_path / "back.blow5"
        ours(str(theirs), "-o", str(back), "-c", "zlib", "-s", "svb-zd")
        again = tmp_path / "again.slow5"
        ref(str(back), "-o", str(again))
        assert filecmp.cmp(again, theirs, shallow=False), name
        # svb-zd only output from text is byte-identical to the reference's
        a, b = tmp_path / "a.blow5", tmp_path / "b.blow5"
        ours(str(theirs), "-o", str(a), "-c", "none", "-s", "svb-zd")
        ref(str(theirs), "-o", str(b), "-c", "none", "-s", "svb-zd")
        assert filecmp.cmp(a, b, shallow=False), name


def test_inconsistent_aux_section_fails_on_the_device_path_too(tmp_path):
    """A record whose auxiliary section does not match the header's columns (here: the count of the channel_number string is one too
    large) stops a conversion with exit 1 on the -K host path (record_parse_binary) and on the device-resident blow5 -> blow5 path
    (rec_locate_kernel walks the section against the layout `view` hands to s5b_ctx_set_aux_layout), like slow5_rec_aux_parse."""
    src = os.path.join(FIX, "exp_1_lossless.blow5")          # uncompressed, five auxiliary fields, channel_number (char*) first
    data = bytearray(open(src, "rb").read())
    hsize = int.from_bytes(data[64:68], "little")
    at = 68 + hsize
    size = int.from_bytes(data[at:at + 8], "little")
    rec = at + 8
    idl = int.from_bytes(data[rec:rec + 2], "little")
    o = rec + 2 + idl + 4 + 32
    ns = int.from_bytes(data[o:o + 8], "little")
    aux = o + 8 + 2 * ns
    cnt = int.from_bytes(data[aux:aux + 8], "little")
    assert 0 < cnt < 16 and aux + 8 + cnt + 8 + 4 + 1 + 8 == rec + size      # the layout this test assumes
    good = tmp_path / "good.blow5"
    ours(src, "-o", str(good), "-c", "zlib", "-s", "svb-zd")                 # the intact file converts
    data[aux:aux + 8] = (cnt + 1).to_bytes(8, "little")
    bad = tmp_path / "bad.blow5"
    bad.write_bytes(bytes(data))
    for extra in (["-c", "zlib", "-s", "svb-zd"], ["-c", "none", "-s", "svb-zd"], ["-c", "zlib", "-s", "svb-zd", "-K", "3"]):
        env = dict(os.environ, S5B_VIEW_SLOW_PATH="1") if "-K" in extra else None
        r = subprocess.run([CLI, "view", str(bad), "-o", str(tmp_path / "o.blow5")] + extra, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, timeout=60, env=env)
        assert r.returncode == 1, (extra, r.returncode, r.stderr.decode())
