"""`slow5tools-b200 get` against the reference binary (`slow5tools get`, src/get.c): same records in the same order, text and
uncompressed BLOW5 output byte-identical; compressed output is read back by the reference to the same text."""
import filecmp
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")
have_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/slow5tools_ref not present")


def ids_of(path):
    out = subprocess.run([REF, "view", path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode()
    return [l.split("\t")[0] for l in out.splitlines() if l and l[0] not in "#@"]


def run(exe, args, stdin=None):
    return subprocess.run([exe, "get"] + args, input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE)


def copies(tmp_path, name):
    a, b = tmp_path / "ours" / name, tmp_path / "ref" / name
    for p in (a, b):
        os.makedirs(p.parent, exist_ok=True)
        shutil.copy(os.path.join(FIX, name), p)
    return str(a), str(b)


def check_file(tmp_path, name):
    a, b = copies(tmp_path, name)
    ids = ids_of(a)
    pick = [ids[-1], ids[0]] + ids[1:3]                      # out of file order, as given
    # ids on the command line -> SLOW5 text on stdout (index built on the fly by both)
    mine, theirs = run(CLI, [a, "--to", "slow5"] + pick), run(REF, [b, "--to", "slow5"] + pick)
    assert mine.returncode == 0, mine.stderr.decode()
    assert mine.stdout == theirs.stdout
    assert filecmp.cmp(a + ".idx", b + ".idx", shallow=False)
    # ids from a list file and from stdin, uncompressed BLOW5 out: byte-identical files
    lst = tmp_path / "ids.txt"
    lst.write_text("\n".join(pick) + "\n\n")
    o1, o2, o3 = tmp_path / "o1.blow5", tmp_path / "o2.blow5", tmp_path / "o3.blow5"
    assert run(CLI, [a, "-l", str(lst), "-c", "none", "-s", "none", "-o", str(o1)]).returncode == 0
    assert run(REF, [b, "-l", str(lst), "-c", "none", "-s", "none", "-o", str(o2)]).returncode == 0
    assert filecmp.cmp(o1, o2, shallow=False)
    assert run(CLI, [a, "--to", "blow5", "-c", "none", "-s", "none", "-o", str(o3), "-K", "2"],
               stdin=("\r\n".join(pick) + "\n").encode()).returncode == 0
    assert filecmp.cmp(o3, o2, shallow=False)
    return a, b, pick


@have_ref
@pytest.mark.parametrize("name", ["exp_1_lossless.blow5", "exp_1_lossless.slow5"])
def test_uncompressed_inputs_match_the_reference(tmp_path, name):
    check_file(tmp_path, name)


@have_ref
def test_missing_ids(tmp_path):
    a, b = copies(tmp_path, "exp_1_lossless.blow5")
    ids = ids_of(a)
    r = run(CLI, [a, "not-a-read", ids[0]])
    assert r.returncode != 0                                   # src/get.c:398-402
    r = run(CLI, [a, "--to", "slow5", "--skip", "not-a-read", ids[0]])
    t = run(REF, [b, "--to", "slow5", "--skip", "not-a-read", ids[0]])
    assert r.returncode == 0 and r.stdout == t.stdout


@pytest.mark.gpu
@have_ref
@pytest.mark.parametrize("name", ["exp_1_lossless_zlib_svb_v0.2.0.blow5", "exp_1_lossless_zstd_svb_v0.2.0.blow5",
                                  "exp_1_lossless_zlib_ex_zd.blow5", "zlib_svb-zd_multi_rg_v0.2.0.blow5"])
def test_compressed_inputs_match_the_reference(tmp_path, name):
    a, b, pick = check_file(tmp_path, name)
    # compressed output: the reference reads ours back to the text it produces itself
    z = tmp_path / "z.blow5"
    assert run(CLI, [a] + pick + ["-o", str(z), "-c", "zlib", "-s", "svb-zd"]).returncode == 0
    text = subprocess.run([REF, "view", str(z)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    assert text == run(REF, [b, "--to", "slow5"] + pick).stdout
    # default output of get is blow5 zlib+svb-zd on stdout
    d = run(CLI, [a] + pick)
    assert d.returncode == 0 and d.stdout[:6] == b"BLOW5\x01" and d.stdout[-5:] == b"5WOLB"


@have_ref
def test_custom_and_malformed_index(tmp_path):
    a, b = copies(tmp_path, "exp_1_lossless.blow5")
    ids = ids_of(a)
    subprocess.check_call([CLI, "index", a], stderr=subprocess.DEVNULL)
    moved = tmp_path / "elsewhere.idx"
    shutil.move(a + ".idx", moved)
    r = run(CLI, [a, "--index", str(moved), "--to", "slow5", ids[0]])
    t = run(REF, [b, "--to", "slow5", ids[0]])
    assert r.returncode == 0 and r.stdout == t.stdout and not os.path.exists(a + ".idx")
    bad = tmp_path / "bad.idx"
    raw = open(moved, "rb").read()
    bad.write_bytes(b"X" + raw[1:])                                  # wrong magic number
    assert run(CLI, [a, "--index", str(bad), "--to", "slow5", ids[0]]).returncode != 0
    bad.write_bytes(raw[:-8])                                        # end marker missing
    assert run(CLI, [a, "--index", str(bad), "--to", "slow5", ids[0]]).returncode != 0
    bad.write_bytes(raw[:80] + raw[-8:])                             # entry cut short
    assert run(CLI, [a, "--index", str(bad), "--to", "slow5", ids[0]]).returncode != 0
