"""`split -x` -- demultiplexing by a read id -> category table (src/demux.c, driven from src/split.c:195-225,369)
-- against the reference's own inputs and expected output directories (test/test_split.sh "demux" cases:
tests/golden/demux_fixtures.tar.xz, packed by make_demux_fixtures.sh).  SLOW5 text in and out involves no codec and runs on the
CPU; the cases that read or write compressed BLOW5 are marked gpu.  Compressed outputs are compared after `view -c none`
(record compression is ours; the svb-zd streams inside are the reference's byte for byte)."""
import filecmp
import os
import subprocess
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
TARBALL = os.path.join(ROOT, "tests", "golden", "demux_fixtures.tar.xz")


@pytest.fixture(scope="module")
def fx(tmp_path_factory):
    d = tmp_path_factory.mktemp("demux_fx")
    with tarfile.open(TARBALL) as t:
        t.extractall(d, filter="data")
    return os.path.join(str(d), "demux")


def split(fx, out, table, src, *flags):
    return subprocess.run([CLI, "split", "-x", os.path.join(fx, "raw", table), os.path.join(fx, "raw", src), "-d", str(out)] + list(flags),
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)


def same_dir(out, exp):
    c = filecmp.dircmp(out, exp)
    return not c.left_only and not c.right_only and all(filecmp.cmp(os.path.join(out, f), os.path.join(exp, f), shallow=False)
                                                          for f in c.common_files) and c.common_files


TEXT_CASES = [  # (name of the expected directory, table, input, flags) -- test/test_split.sh:205-325
    ("demux1", "demux1/barcode_summary.txt", "demux1/example2_0.slow5", []),
    ("demux1", "demux1/barcode_summary.txt", "demux1/example2_0.slow5", ["-u", "mixed"]),
    ("demux9", "demux9/barcode_summary.txt", "demux2/example2_0.slow5", ["-u", "vmixed", "--demux-rid", "rid", "--demux-code", "code"]),
    ("demux10", "demux9/barcode_summary.txt", "demux10/example2_0_multi.slow5", ["--demux-rid", "rid", "--demux-code", "code"]),
    ("demux10-uniq", "demux9/barcode_summary.txt", "demux10/example2_0_multi.slow5",
     ["--demux-rid", "rid", "--demux-code", "code", "-u", "vmixed"]),
    ("demux10-uniq-lossy", "demux9/barcode_summary.txt", "demux10/example2_0_multi.slow5",
     ["--demux-rid", "rid", "--demux-code", "code", "-u", "vmixed", "--lossless", "false"]),
    ("demux11", "demux11/onemissing.txt", "demux10/example2_0_multi.slow5",
     ["--demux-rid", "rid", "--demux-code", "code", "-m", "missing"]),
]


@pytest.mark.parametrize("exp,table,src,flags", TEXT_CASES)
def test_text_goldens(fx, tmp_path, exp, table, src, flags):
    out = tmp_path / "out"
    r = split(fx, out, table, src, "--to", "slow5", *flags)
    assert r.returncode == 0, r.stderr.decode()
    assert same_dir(str(out), os.path.join(fx, "exp", exp))


def test_category_name_collisions_and_bad_tables(fx, tmp_path):
    """test/test_split.sh:292-336: -u / -m naming a category of the table, or each other"""
    for i, flags in enumerate((["-u", "unclassified"], ["-m", "unclassified"], ["-u", "mixed", "-m", "mixed"])):
        r = split(fx, tmp_path / ("c%d" % i), "demux1/barcode_summary.txt", "demux1/example2_0.slow5", "--to", "slow5", *flags)
        assert r.returncode == 1 and b"already exists" in r.stderr, flags
    r = split(fx, tmp_path / "h", "demux1/barcode_summary.txt", "demux1/example2_0.slow5", "--to", "slow5", "--demux-rid", "nope")
    assert r.returncode == 1 and b"Invalid demux TSV header: missing 'nope'" in r.stderr
    r = split(fx, tmp_path / "t", "demux1/no_such_table", "demux1/example2_0.slow5", "--to", "slow5")
    assert r.returncode == 1 and b"Failed to open" in r.stderr
    # a read of the file that the table does not list: dropped with a warning unless -m catches it
    r = split(fx, tmp_path / "w", "demux11/onemissing.txt", "demux10/example2_0_multi.slow5", "--to", "slow5",
              "--demux-rid", "rid", "--demux-code", "code")
    assert r.returncode == 0 and b"is missing from demux TSV" in r.stderr
    assert sorted(os.listdir(tmp_path / "w")) == ["example2_0_multi_1.slow5"]


REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/slow5tools_ref not present")
@pytest.mark.parametrize("table,src,flags", [
    ("demux3/bs.txt", "demux3/example2_0.slow5", []),
    ("demux3/bs.txt", "demux3/example2_0.slow5", ["-u", "mixed"]),
    ("demux9/barcode_summary.txt", "demux10/example2_0_multi.slow5", ["--demux-rid", "rid", "--demux-code", "code", "-m", "rest"]),
])
def test_uncompressed_blow5_output_equals_the_reference_binary(fx, tmp_path, table, src, flags):
    """no codec involved (-c none -s none): the files must be the reference's byte for byte, names included"""
    ours, theirs = tmp_path / "ours", tmp_path / "theirs"
    r = split(fx, ours, table, src, "--to", "blow5", "-c", "none", "-s", "none", *flags)
    assert r.returncode == 0, r.stderr.decode()
    q = subprocess.run([REF, "split", "-x", os.path.join(fx, "raw", table), os.path.join(fx, "raw", src), "-d", str(theirs),
                        "--to", "blow5", "-c", "none", "-s", "none"] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert q.returncode == 0, q.stderr.decode()
    assert same_dir(str(ours), str(theirs))


def view_none(src, dst):
    r = subprocess.run([CLI, "view", src, "-o", dst, "-c", "none", "-s", "svb-zd"], stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()


@pytest.mark.gpu
@pytest.mark.parametrize("exp,table,src,flags", [
    ("demux2", "demux2/barcode_summary.txt", "demux2/example2_0.slow5", []),
    ("demux2", "demux2/barcode_summary.txt", "demux2/example2_0.slow5", ["--demux-uniq", "mixed"]),
    ("demux3", "demux3/bs.txt", "demux3/example2_0.slow5", []),
    ("demux3-uniq", "demux3/bs.txt", "demux3/example2_0.slow5", ["-u", "mixed"]),
])
def test_blow5_output_goldens(fx, tmp_path, exp, table, src, flags):
    """SLOW5 in, zlib + svb-zd BLOW5 out (test/test_split.sh:213-222, :256-265)"""
    out = tmp_path / "out"
    r = split(fx, out, table, src, "--to", "blow5", *flags)
    assert r.returncode == 0, r.stderr.decode()
    want = os.path.join(fx, "exp", exp)
    assert sorted(os.listdir(out)) == sorted(os.listdir(want))
    for f in os.listdir(want):
        a, b = str(tmp_path / ("a_" + f)), str(tmp_path / ("b_" + f))
        view_none(os.path.join(str(out), f), a)
        view_none(os.path.join(want, f), b)
        assert filecmp.cmp(a, b, shallow=False), f
        assert os.path.getsize(os.path.join(str(out), f)) < 1.03 * os.path.getsize(os.path.join(want, f))


@pytest.mark.gpu
@pytest.mark.parametrize("exp,table,flags", [
    ("demux4", "demux4/summary", []),
    ("demux4-uniq", "demux4/summary", ["--demux-uniq", "mixed"]),
    ("demux5", "demux5/custom", ["--demux-rid=MyCustomId", "--demux-code", "BC0D35!"]),
    ("demux5", "demux5/custom_rev", ["--demux-code", "BC0D35!", "--demux-rid=MyCustomId"]),
    ("demux5-uniq", "demux5/custom", ["-u", "abc", "--demux-rid=MyCustomId", "--demux-code", "BC0D35!"]),
    ("demux7", "demux7/barcode_summary.txt", []),
    ("demux7-uniq", "demux7/barcode_summary.txt", ["--demux-uniq", "hodjbodj"]),
])
def test_blow5_input_goldens(fx, tmp_path, exp, table, flags):
    """zlib + svb-zd BLOW5 in, SLOW5 out (test/test_split.sh:225-245, :268-288, :339)"""
    out = tmp_path / "out"
    r = split(fx, out, table, table.split("/")[0] + "/example2_0.blow5", "--to", "slow5", *flags)
    assert r.returncode == 0, r.stderr.decode()
    assert same_dir(str(out), os.path.join(fx, "exp", exp))


@pytest.mark.gpu
def test_table_with_extra_reads_fails(fx, tmp_path):
    """test/test_split.sh:237-241 (demux6): the table lists reads the file does not hold"""
    for i, flags in enumerate(([], ["-u", "what"])):
        r = split(fx, tmp_path / ("x%d" % i), "demux6/barcode_summary.txt", "demux6/example2_0.blow5", "--to", "slow5", *flags)
        assert r.returncode == 1 and b"Extra read(s) in demux TSV" in r.stderr
