"""BASELINE config[0] -- `view` plumbing on the reference's small fixtures, CPU only (no codec involved:
none/none files), byte-identical to the reference's goldens / its own binary."""
import filecmp
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")


def run(args, **kw):
    return subprocess.run([CLI] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


def test_cli_is_built():
    assert os.path.exists(CLI), "run __graft_entry__.build()"
    assert b"slow5tools-b200" in run(["--version"]).stdout


def test_blow5_to_slow5_matches_reference_golden(tmp_path):
    out = tmp_path / "a.slow5"
    assert run(["view", os.path.join(FIX, "exp_1_lossless.blow5"), "-o", str(out)]).returncode == 0
    assert filecmp.cmp(out, os.path.join(FIX, "exp_1_lossless.slow5"), shallow=False)
    # default output is SLOW5 on stdout (src/misc.c:53, view.c:160-162)
    r = run(["view", os.path.join(FIX, "exp_1_lossless.blow5")])
    assert r.returncode == 0 and r.stdout == open(os.path.join(FIX, "exp_1_lossless.slow5"), "rb").read()


def test_slow5_to_blow5_and_back(tmp_path):
    b = tmp_path / "b.blow5"
    assert run(["view", os.path.join(FIX, "exp_1_lossless.slow5"), "-o", str(b), "-c", "none", "-s", "none"]).returncode == 0
    assert filecmp.cmp(b, os.path.join(FIX, "exp_1_lossless.blow5"), shallow=False)
    c = tmp_path / "c.blow5"
    assert run(["view", str(b), "-o", str(c), "-c", "none", "-s", "none", "-t", "3", "-K", "1"]).returncode == 0
    assert filecmp.cmp(c, b, shallow=False)


def test_against_reference_binary_multi_read_group(tmp_path):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/slow5tools_ref not built")
    src = os.path.join(FIX, "zlib_svb-zd_multi_rg_v0.2.0.blow5")
    none = tmp_path / "none.blow5"
    subprocess.check_call([REF, "view", src, "-c", "none", "-s", "none", "-o", str(none)], stderr=subprocess.DEVNULL)
    ours = tmp_path / "ours.slow5"
    assert run(["view", str(none), "-o", str(ours)]).returncode == 0
    assert filecmp.cmp(ours, os.path.join(FIX, "zlib_svb-zd_multi_rg_v0.2.0.expected.slow5"), shallow=False)
    back_ref, back_ours = tmp_path / "r.blow5", tmp_path / "o.blow5"
    subprocess.check_call([REF, "view", str(ours), "-c", "none", "-s", "none", "-o", str(back_ref)], stderr=subprocess.DEVNULL)
    assert run(["view", str(ours), "-c", "none", "-s", "none", "-o", str(back_ours)]).returncode == 0
    assert filecmp.cmp(back_ref, back_ours, shallow=False)


def test_failure_cases(tmp_path):
    f = os.path.join(FIX, "exp_1_lossless.blow5")
    assert run(["view", "/nonexistent.blow5"]).returncode == 1
    assert run(["view", f, "--to", "slow5", "-o", str(tmp_path / "x.blow5")]).returncode == 1      # test_view.sh:236
    assert run(["view", f, "-c", "lz4", "-o", str(tmp_path / "x.blow5")]).returncode == 1           # unknown method name
    assert run(["view", f, "-c", "zlib"]).returncode == 1                                           # -c with ASCII output
    assert run(["view"]).returncode == 1
    trunc = tmp_path / "t.blow5"
    trunc.write_bytes(open(f, "rb").read()[:-9])
    assert run(["view", str(trunc)]).returncode == 1                                                # no EOF marker


def test_compressed_input_needs_a_gpu_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run(["view", os.path.join(FIX, "exp_1_lossless_zlib_svb_v0.2.0.blow5")])
    assert r.returncode == 1 and b"GPU" in r.stderr
