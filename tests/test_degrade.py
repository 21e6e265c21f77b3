"""`degrade` (src/degrade.c; the per-sample step slow5_arr_qts_round, slow5lib/src/slow5_press.c:1965-2005)
against the reference's own inputs and expected outputs (test/test_degrade.sh: tests/golden/degrade_fixtures.tar.xz, packed by
make_degrade_fixtures.sh) and, where it has been built in this container, against the compiled reference.

CPU part: the oracle restatement (oracle/qts_oracle.c) is pinned to the golden pair example2.slow5 -> example2_b1.slow5 and to
the compiled reference; option parsing and the header's dataset detection of the CLI need no device.  GPU part: the kernel
through the C-ABI against the oracle for every bit count, and the CLI against the goldens -- SLOW5 text byte-identical, BLOW5
compared after both files went through `view -c none` (the record compression is ours, the ex-zd / svb-zd streams inside must be
the reference's byte for byte)."""
import ctypes as C
import filecmp
import os
import subprocess
import tarfile

import numpy as np
import pytest

from conftest import build_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libslow5_ref.so")
TARBALL = os.path.join(ROOT, "tests", "golden", "degrade_fixtures.tar.xz")


@pytest.fixture(scope="module")
def fx(tmp_path_factory):
    d = tmp_path_factory.mktemp("degrade_fx")
    with tarfile.open(TARBALL) as t:
        t.extractall(d, filter="data")
    return os.path.join(str(d), "degrade")


@pytest.fixture(scope="module")
def orc():
    L = C.CDLL(build_oracle())
    L.orc_qts_round.restype = None
    L.orc_qts_round.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
    L.orc_qts_round_sample.restype = C.c_int
    L.orc_qts_round_sample.argtypes = [C.c_int, C.c_int]
    return L


def orc_round(orc, x, bits):
    y = np.ascontiguousarray(x, dtype=np.int16).copy()
    orc.orc_qts_round(y.ctypes.data, y.size, bits)
    return y


def run(args, env=None):
    return subprocess.run([CLI] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300,
                          env=dict(os.environ, **env) if env else None)


def slow5_signals(path):
    """raw_signal columns of a SLOW5 text file as int16 arrays, and everything else of every line"""
    sigs, rest = [], []
    for l in open(path, "rb").read().split(b"\n"):
        if not l or l[:1] in (b"#", b"@"):
            rest.append(l)
            continue
        c = l.split(b"\t")
        sigs.append(np.array(c[7].split(b","), dtype=np.int64).astype(np.int16))
        rest.append(b"\t".join(c[:7] + c[8:]))
    return sigs, rest


def adversarial(rng, n):
    x = rng.integers(-32768, 32768, n).astype(np.int16)
    edge = np.array([-32768, -32767, -1, 0, 1, 2, 3, 4, 7, 8, 15, 16, 127, 128, 255, 256, 16383, 16384, 32766, 32767], np.int16)
    x[:min(n, edge.size)] = edge[:min(n, edge.size)]
    return x


# ---------------------------------------------------------------- CPU: oracle pinning, host logic
def test_oracle_matches_reference_golden_pair(fx, orc):
    """test/test_degrade.sh testcase 1: example2.slow5 --bits=1 -> example2_b1.slow5"""
    raw, raw_rest = slow5_signals(os.path.join(fx, "raw", "example2.slow5"))
    exp, exp_rest = slow5_signals(os.path.join(fx, "exp", "example2_b1.slow5"))
    assert raw_rest == exp_rest and len(raw) == len(exp) == 8
    changed = 0
    for a, b in zip(raw, exp):
        got = orc_round(orc, a, 1)
        assert np.array_equal(got, b)
        changed += int((a != b).sum())
    assert changed > 100000  # the pair is not a trivial one


def test_oracle_closed_form_and_edges(orc):
    """(x + 2^(b-1)) & ~(2^b - 1) in 16-bit wrap-around arithmetic -- the form the kernel uses -- is the reference's rule"""
    x = np.arange(-32768, 32768, dtype=np.int32)
    for b in range(1, 17):
        want = orc_round(orc, x.astype(np.int16), b)
        got = (((x + (1 << (b - 1))) & ~((1 << b) - 1)) & 0xffff).astype(np.uint16).view(np.int16)
        assert np.array_equal(got, want), b
    assert orc.orc_qts_round_sample(32767, 1) == 32768 and orc_round(orc, np.array([32767], np.int16), 1)[0] == -32768
    assert orc.orc_qts_round_sample(-1, 16) == 0 and orc.orc_qts_round_sample(5, 3) == 8 and orc.orc_qts_round_sample(3, 3) == 0
    y = np.array([1, 2, 3], np.int16)
    assert np.array_equal(orc_round(orc, y, 0), y)  # slow5_press.c:1995-1996


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libslow5_ref.so not present")
def test_oracle_matches_compiled_reference(orc):
    R = C.CDLL(REF_SO)
    R.slow5_arr_qts_round.restype = None
    R.slow5_arr_qts_round.argtypes = [C.c_void_p, C.c_uint64, C.c_uint8]
    rng = np.random.default_rng(7)
    for b in range(0, 17):
        x = adversarial(rng, 5000 + b)
        y = x.copy()
        R.slow5_arr_qts_round(y.ctypes.data, y.size, b)
        assert np.array_equal(orc_round(orc, x, b), y), b


def test_cli_rejects_bad_bits_and_arguments(fx):
    for bad in ("0", "17", "-3", "x", "3x", ""):
        r = run(["degrade", "-b", bad, os.path.join(fx, "raw", "example2.slow5")])
        assert r.returncode == 1 and b"Invalid bits argument" in r.stderr, bad
    r = run(["degrade"])
    assert r.returncode == 1 and b"Usage" in r.stderr
    r = run(["degrade", "-b", "2"])
    assert r.returncode == 1 and b"missing input file" in r.stderr
    r = run(["degrade", "-b", "2", "a.slow5", "b.slow5"])
    assert r.returncode == 1 and b"more than 1 input file" in r.stderr
    r = run(["degrade", "-b", "2", "-c", "zlib", os.path.join(fx, "raw", "example2.slow5")])
    assert r.returncode == 1 and b"only valid for blow5" in r.stderr
    r = run(["degrade", "--help"])
    assert r.returncode == 0 and b"--bits" in r.stdout and b"[ex-zd]" in r.stdout


@pytest.mark.parametrize("name", ["promr10dna_badhdr", "promr10dna_badhdr_sample_freq", "promr10dna_badhdr_sample_rate",
                                  "promr10dna_badhdr_sample_rate2"])
def test_cli_auto_bits_refuses_unknown_headers(fx, name):
    """test/test_degrade.sh testcases 8-11: no dataset matches, nothing is written, exit 1 (decided before any device work)"""
    r = run(["degrade", os.path.join(fx, "raw", name + ".slow5")])
    assert r.returncode == 1 and r.stdout == b"" and b"No suitable bits suggestion" in r.stderr


DETECT = [("minir10dna.blow5", "DNA lsk114 5kHz MinION", 3), ("promr10dna4khz.blow5", "DNA lsk114 4kHz PromethION", 3),
          ("PRPN119035_read1.blow5", "RNA rna002 3kHz PromethION", 2),
          ("na12878_prom_merged_r9.4.1_chr22_read1.blow5", "DNA lsk109 4kHz PromethION", 2),
          ("promr10dna_badrec.slow5", "DNA lsk114 5kHz PromethION", 3)]


@pytest.mark.parametrize("name,dataset,bits", DETECT)
def test_cli_detects_the_dataset_like_the_reference(fx, tmp_path, name, dataset, bits):
    """src/degrade.c:124-148: the header names the dataset and with it the bit count (printed before the conversion starts)"""
    r = run(["degrade", os.path.join(fx, "raw", name), "-o", str(tmp_path / "o.blow5")])
    assert ("Detected: %s" % dataset).encode() in r.stderr and ("Eliminating %d bits" % bits).encode() in r.stderr
    if os.path.exists(REF):
        q = subprocess.run([REF, "degrade", os.path.join(fx, "raw", name), "-o", str(tmp_path / "r.blow5")],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert ("Detected: %s" % dataset).encode() in q.stderr and ("Eliminating %d bits" % bits).encode() in q.stderr


# ---------------------------------------------------------------- GPU: kernel and CLI parity
@pytest.mark.gpu
def test_kernel_matches_oracle_all_bit_counts(orc):
    import torch
    from slow5tools_b200.codec import Codec
    cd = Codec(0)
    rng = np.random.default_rng(11)
    for b in range(1, 17):
        for n, skew in ((0, 0), (1, 0), (7, 1), (8, 0), (9, 3), (4099, 5), (1 << 20, 0), ((1 << 20) + 13, 7)):
            x = adversarial(rng, n + skew)
            t = torch.from_numpy(x.copy()).cuda()
            cd.qts_round_dev(t[skew:], b, n)  # pointer `skew` samples past a 16-byte boundary: head / body / tail paths
            torch.cuda.synchronize()
            want = x.copy()
            want[skew:] = orc_round(orc, x[skew:], b)
            assert np.array_equal(t.cpu().numpy(), want), (b, n, skew)
    # bits = 0 leaves the samples alone; the batch form returns degraded copies of unaligned host arrays
    x = adversarial(rng, 1000)
    t = torch.from_numpy(x.copy()).cuda()
    cd.qts_round_dev(t, 0)
    assert np.array_equal(t.cpu().numpy(), x)
    bufs = [adversarial(rng, n).tobytes() for n in (0, 1, 5, 4096, 70001)]
    rc, out = cd.qts_round_batch(3, bufs)
    assert rc == 0
    for a, o in zip(bufs, out):
        assert o == orc_round(orc, np.frombuffer(a, np.int16), 3).tobytes()
    rc, out = cd.qts_round_batch(3, [b"\x01\x02\x03"])  # an odd byte count is not an int16 array
    assert rc == -2


@pytest.mark.gpu
def test_transcoder_degrades_between_decode_and_encode(orc):
    """s5b_ctx_set_degrade: every method pair of the whole-batch transcoder (also same -> same) stores round(signal)"""
    from slow5tools_b200.codec import Codec
    from recode_helpers import make_records, walk_image, M_NONE, M_ZLIB, M_SVB_ZD, M_ZSTD, M_EX_ZD
    cd = Codec(0)
    lengths = [1, 5, 4096, 777, 30000, 12, 4097, 2]  # (ex-zd has no empty stream: slow5_press.c:1723 asserts)
    recs, sigs = make_records(lengths, seed=5)
    plain = Codec(0)
    for bits in (1, 3, 16):
        cd.set_degrade(bits)
        want_recs = []
        for r, s in zip(recs, sigs):
            body = len(r) - 2 * len(s)
            want_recs.append(r[:body] + orc_round(orc, s, bits).tobytes())
        for in_m, out_m in (((M_NONE, M_NONE), (M_ZLIB, M_EX_ZD)), ((M_ZLIB, M_EX_ZD), (M_ZLIB, M_EX_ZD)),
                            ((M_ZLIB, M_SVB_ZD), (M_NONE, M_SVB_ZD)), ((M_NONE, M_SVB_ZD), (M_ZSTD, M_NONE)),
                            ((M_NONE, M_NONE), (M_NONE, M_NONE))):
            # the input in its stored form, made by the undegrading context
            rc, img_in = plain.blow5_recode(M_NONE, M_NONE, in_m[0], in_m[1], recs)
            assert rc == 0
            rc, img = cd.blow5_recode(in_m[0], in_m[1], out_m[0], out_m[1], walk_image(img_in))
            assert rc == 0, (bits, in_m, out_m)
            # back to uncompressed records without touching the samples
            rc, flat = plain.blow5_recode(out_m[0], out_m[1], M_NONE, M_NONE, walk_image(img))
            assert rc == 0
            assert walk_image(flat) == want_recs, (bits, in_m, out_m)
    # the dataset rule: records carry digitisation 8192 / sampling rate 4000 (recode_helpers.make_records)
    cd.set_degrade(3, True, 8192.0, 4000.0)
    rc, img = cd.blow5_recode(M_NONE, M_NONE, M_NONE, M_EX_ZD, recs)
    assert rc == 0
    cd.set_degrade(3, True, 2048.0, 4000.0)
    rc, img = cd.blow5_recode(M_NONE, M_NONE, M_NONE, M_EX_ZD, recs)
    assert rc == -42
    cd.set_degrade(0)
    rc, img = cd.blow5_recode(M_NONE, M_NONE, M_NONE, M_NONE, recs)
    assert rc == 0 and walk_image(img) == recs


def view_none(src, dst):
    """any BLOW5 -> records uncompressed, signal streams carried as stored (view keeps a signal whose method does not change)"""
    name = {0: "none", 1: "svb-zd", 2: "ex-zd"}[open(src, "rb").read(64)[14]]  # slow5_press.c:107-161
    r = run(["view", src, "-o", dst, "-c", "none", "-s", name])
    assert r.returncode == 0, r.stderr.decode()


def same_signal_method(a, b):
    return open(a, "rb").read(64)[14] == open(b, "rb").read(64)[14]


@pytest.mark.gpu
def test_cli_text_golden_byte_identical(fx, tmp_path):
    """testcase 1: SLOW5 in, SLOW5 out, --bits=1"""
    out = tmp_path / "b1.slow5"
    r = run(["degrade", os.path.join(fx, "raw", "example2.slow5"), "-o", str(out), "--bits=1"])
    assert r.returncode == 0, r.stderr.decode()
    assert filecmp.cmp(out, os.path.join(fx, "exp", "example2_b1.slow5"), shallow=False)
    r = run(["degrade", os.path.join(fx, "raw", "example2.slow5"), "-b", "1", "-K", "3", "-t", "2"])  # to stdout, small batches
    assert r.returncode == 0 and r.stdout == open(os.path.join(fx, "exp", "example2_b1.slow5"), "rb").read()


BLOW5_CASES = [  # (input, flags, expected) -- test/test_degrade.sh testcases 2, 3, 5, 13, 14
    ("example2.slow5", ["-b", "4"], "example2_b4.blow5"),
    ("minir10dna.blow5", [], "minir10dna_b3.blow5"),
    ("promr10dna4khz.blow5", ["-s", "svb-zd"], "promr10dna4khz_b3.blow5"),
    ("na12878_prom_merged_r9.4.1_chr22_read1.blow5", [], "na12878_prom_merged_r9.4.1_chr22_read1_b2.blow5"),
    ("PRPN119035_read1.blow5", [], "PRPN119035_read1_b2.blow5"),
]


@pytest.mark.gpu
@pytest.mark.parametrize("src,flags,exp", BLOW5_CASES)
def test_cli_blow5_goldens(fx, tmp_path, src, flags, exp):
    src, exp = os.path.join(fx, "raw", src), os.path.join(fx, "exp", exp)
    want = str(tmp_path / "want.blow5")
    view_none(exp, want)
    # default record compression (our deflate): identical once the records are unwrapped
    out, flat = str(tmp_path / "out.blow5"), str(tmp_path / "flat.blow5")
    r = run(["degrade", src, "-o", out] + flags)
    assert r.returncode == 0, r.stderr.decode()
    assert same_signal_method(out, exp)
    view_none(out, flat)
    assert filecmp.cmp(flat, want, shallow=False)
    assert os.path.getsize(out) < 1.05 * os.path.getsize(exp)  # and about as small as the reference's file
    # without record compression the file itself is the reference's, byte for byte
    r = run(["degrade", src, "-o", out, "-c", "none"] + flags)
    assert r.returncode == 0, r.stderr.decode()
    assert filecmp.cmp(out, want, shallow=False)
    # the per-record path (parse on the host, codec calls batched) writes the same file as the device-resident one
    r = run(["degrade", src, "-o", flat, "-c", "none", "-K", "2"] + flags, env={"S5B_VIEW_SLOW_PATH": "1"})
    assert r.returncode == 0, r.stderr.decode()
    assert filecmp.cmp(flat, want, shallow=False)
    if os.path.exists(REF):  # the reference reads our compressed file to the same text as its own
        a, b = str(tmp_path / "a.slow5"), str(tmp_path / "b.slow5")
        r = run(["degrade", src, "-o", out] + flags)
        subprocess.check_call([REF, "view", out, "-o", a], stderr=subprocess.DEVNULL)
        subprocess.check_call([REF, "view", exp, "-o", b], stderr=subprocess.DEVNULL)
        assert filecmp.cmp(a, b, shallow=False)


@pytest.mark.gpu
def test_cli_same_method_and_text_output(fx, tmp_path):
    """ex-zd in, ex-zd out (the transcoder must not pass the stored streams through), and BLOW5 -> degraded SLOW5 text"""
    src = os.path.join(fx, "exp", "minir10dna_b3.blow5")  # zlib + ex-zd
    fast, slow, text = str(tmp_path / "f.blow5"), str(tmp_path / "s.blow5"), str(tmp_path / "t.slow5")
    assert run(["degrade", "-b", "6", src, "-o", fast, "-c", "none"]).returncode == 0
    assert run(["degrade", "-b", "6", src, "-o", slow, "-c", "none"], env={"S5B_VIEW_SLOW_PATH": "1"}).returncode == 0
    assert filecmp.cmp(fast, slow, shallow=False)
    assert run(["degrade", "-b", "6", src, "-o", text]).returncode == 0
    back = str(tmp_path / "back.slow5")
    assert run(["view", fast, "-o", back]).returncode == 0
    assert filecmp.cmp(text, back, shallow=False)
    raw, _ = slow5_signals(text)
    assert all((s.astype(np.int32) & 63).max() == 0 for s in raw if s.size)  # six low bits are gone


@pytest.mark.gpu
def test_cli_bad_record_fails_on_both_paths(fx, tmp_path):
    """testcase 12: the header names a 5 kHz dataset, the last record says 4 kHz"""
    bad = os.path.join(fx, "raw", "promr10dna_badrec.slow5")
    r = run(["degrade", bad])
    assert r.returncode == 1 and b"0d624d4b-671f-40b8-9798-84f2ccc4d7fc" in r.stderr and b"does not match" in r.stderr
    # the same records as BLOW5: the device-resident transcoder applies the rule (S5B_ERR_DATASET)
    blow = str(tmp_path / "bad.blow5")
    assert run(["view", bad, "-o", blow]).returncode == 0
    r = run(["degrade", blow, "-o", str(tmp_path / "o.blow5")])
    assert r.returncode == 1 and b"does not match" in r.stderr
    # with the bit count given the records are not held to a dataset (src/degrade.c:249)
    r = run(["degrade", "-b", "3", blow, "-o", str(tmp_path / "o.blow5")])
    assert r.returncode == 0, r.stderr.decode()
