"""`merge` and `split` (src/merge.c, src/split.c) against the reference's own expected outputs
(test/test_merge.sh, test/test_split.sh: tests/golden/merge_split_fixtures.tar.xz, packed by make_merge_split_fixtures.sh)
and, where it has been built in this container, against the reference binary itself.  SLOW5-text and uncompressed-BLOW5
cases involve no codec and run on the CPU; the compressed ones are marked gpu."""
import filecmp
import os
import subprocess
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")
TARBALL = os.path.join(ROOT, "tests", "golden", "merge_split_fixtures.tar.xz")


@pytest.fixture(scope="module")
def fx(tmp_path_factory):
    d = tmp_path_factory.mktemp("merge_split_fx")
    with tarfile.open(TARBALL) as t:
        t.extractall(d, filter="data")
    return str(d)


def run(args):
    return subprocess.run([CLI] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)


def raw(fx, *names):
    return [os.path.join(fx, "merge", "raw", n) for n in names]


def records(path):
    return [l for l in open(path, "rb").read().split(b"\n") if l and l[:1] not in (b"#", b"@")]


def many_reads(fx, path, n=11):
    """a single read group SLOW5 file of n records: the header and first record of rg0.slow5, the record repeated with its own
    read id, a shortened signal and different auxiliary values"""
    lines = open(os.path.join(fx, "merge", "raw", "rg0.slow5"), "rb").read().split(b"\n")
    head = [l for l in lines if l[:1] in (b"#", b"@")]
    col = [l for l in lines if l and l[:1] not in (b"#", b"@")][0].split(b"\t")
    out = []
    for i in range(n):
        c = list(col)
        c[0] = b"read-%04d" % i
        sig = c[7].split(b",")[:50 + 7 * i]
        c[6] = str(len(sig)).encode()
        c[7] = b",".join(sig)
        out.append(b"\t".join(c))
    with open(path, "wb") as f:
        f.write(b"\n".join(head + out) + b"\n")
    return str(path)


MERGE_GOLDENS = [   # (expected file, extra flags, inputs)  -- the numbering is test/test_merge.sh's
    ("same_rg.slow5", [], ["rg0.slow5", "rg0_1.slow5"]),                                                    # 1.9
    ("same_rg_aux_order.slow5", [], ["rg0.slow5", "rg0_2_aux_order.slow5"]),                                # 1.11
    ("merged_output_enum.slow5", [], ["aux_no_enum.slow5", "aux_enum.slow5"]),                              # 2.1
    ("asic_id_missing_expected.slow5", ["-a"], ["rg0_asic_id_missing.slow5", "rg0.slow5"]),                 # 3.2
    ("asic_id_missing_expected.slow5", ["-a"], ["rg0.slow5", "rg0_asic_id_missing.slow5"]),                 # 3.4
    ("same_run_id_different_attribute_values.slow5", ["--allow"], ["rg0.slow5", "rg0_diff_attr.slow5"]),    # 3.6
    ("diff_rg_diff_aux_field.slow5", [], ["rg0.slow5", "rg1_1_new_aux_field.slow5"]),                       # 4.4
]


@pytest.mark.parametrize("exp,flags,inputs", MERGE_GOLDENS)
def test_merge_matches_reference_goldens(fx, tmp_path, exp, flags, inputs):
    out = tmp_path / "out.slow5"
    r = run(["merge"] + flags + raw(fx, *inputs) + ["-o", str(out), "-t", "2"])
    assert r.returncode == 0, r.stderr.decode()
    assert filecmp.cmp(out, os.path.join(fx, "merge", "exp", exp), shallow=False)
    # to standard output with --to (test 1.3)
    r = run(["merge"] + flags + raw(fx, *inputs) + ["--to", "slow5"])
    assert r.returncode == 0 and r.stdout == open(os.path.join(fx, "merge", "exp", exp), "rb").read()


@pytest.mark.parametrize("inputs,message", [
    (["aux_enum.slow5", "aux_enum_diff_label.slow5"], b"Attribute end_reason has different order/name of the enum labels in different files"),   # 2.3
    (["aux_enum.slow5", "aux_enum_new_label.slow5"], b"Attribute end_reason has different number of enum labels in different files"),           # 2.4
    (["aux_enum.slow5", "aux_enum_uint8_t.slow5"], b"ERROR"),
    (["aux_enum_uint8_t.slow5", "aux_enum.slow5"], b"ERROR"),
    (["rg0_asic_id_missing.slow5", "rg0.slow5"], b"Attributes are different for the same run_id"),                                             # 3.1
    (["rg0.slow5", "rg0_asic_id_missing.slow5"], b"Attributes are different for the same run_id"),                                             # 3.3
    (["rg0.slow5", "rg0_diff_attr.slow5"], b"Attributes are different for the same run_id"),                                                   # 3.5
])
def test_merge_refusals(fx, tmp_path, inputs, message):
    r = run(["merge"] + raw(fx, *inputs) + ["-o", str(tmp_path / "x.slow5")])
    assert r.returncode != 0 and message in r.stderr


def test_merge_argument_errors(fx, tmp_path):
    assert run(["merge"]).returncode != 0
    assert run(["merge", str(tmp_path)]).returncode != 0                                      # a directory without slow5 files
    assert run(["merge"] + raw(fx, "rg0.slow5") + ["-o", str(tmp_path / "x.txt")]).returncode != 0
    assert run(["merge"] + raw(fx, "rg0.slow5") + ["--to", "slow5", "-c", "zlib"]).returncode != 0
    assert run(["merge"] + raw(fx, "rg0.slow5") + ["--lossless", "maybe"]).returncode != 0


def test_merge_to_uncompressed_blow5_and_back(fx, tmp_path):
    b = tmp_path / "m.blow5"
    r = run(["merge"] + raw(fx, "rg0.slow5", "rg1_1_new_aux_field.slow5") + ["-o", str(b), "-c", "none", "-s", "none"])
    assert r.returncode == 0, r.stderr.decode()
    s = tmp_path / "m.slow5"
    assert run(["view", str(b), "-o", str(s)]).returncode == 0
    assert filecmp.cmp(s, os.path.join(fx, "merge", "exp", "diff_rg_diff_aux_field.slow5"), shallow=False)
    # a directory as input: both files are found (readdir order decides which comes first, as in the reference)
    d = tmp_path / "in"
    d.mkdir()
    for n in ("rg0.slow5", "rg1_1_new_aux_field.slow5"):
        os.symlink(os.path.join(fx, "merge", "raw", n), d / n)
    r = run(["merge", str(d), "--to", "slow5"])
    assert r.returncode == 0 and b"#num_read_groups\t2" in r.stdout


def test_split_by_groups_inverts_merge(fx, tmp_path):
    merged = os.path.join(fx, "merge", "exp", "diff_rg_diff_aux_field.slow5")     # two read groups
    out = tmp_path / "g"
    r = run(["split", "-g", merged, "-d", str(out), "--to", "slow5"])
    assert r.returncode == 0, r.stderr.decode()
    files = sorted(os.listdir(out))
    assert files == ["diff_rg_diff_aux_field_0.slow5", "diff_rg_diff_aux_field_1.slow5"]
    all_in = records(merged)
    per_group = [records(out / f) for f in files]
    assert sum(len(g) for g in per_group) == len(all_in)
    for g, recs in enumerate(per_group):
        want = [l for l in all_in if l.split(b"\t")[1] == str(g).encode()]
        # the same records, read_group rewritten to 0 (split.c:88-89)
        assert [l.split(b"\t")[:1] + l.split(b"\t")[2:] for l in recs] == [l.split(b"\t")[:1] + l.split(b"\t")[2:] for l in want]
        assert all(l.split(b"\t")[1] == b"0" for l in recs)
        head = open(out / files[g], "rb").read()
        assert b"#num_read_groups\t1\n" in head
    # merging the parts again gives the merged file back (test/test_merge_split_integrity.sh)
    again = tmp_path / "again.slow5"
    assert run(["merge", str(out / files[0]), str(out / files[1]), "-o", str(again)]).returncode == 0
    assert filecmp.cmp(again, merged, shallow=False)
    # refusals (split.c:331-344)
    assert b"already has a single read group" in run(["split", "-g", str(out / files[0]), "-d", str(tmp_path / "h")]).stderr
    assert b"contains multiple read groups" in run(["split", "-r", "2", merged, "-d", str(tmp_path / "i")]).stderr
    assert b"is not empty" in run(["split", "-g", merged, "-d", str(out)]).stderr
    assert run(["split", "-g", merged]).returncode != 0                            # no output directory
    assert run(["split", merged, "-d", str(tmp_path / "j")]).returncode != 0       # no -r count


@pytest.mark.parametrize("how,count", [("-r", 2), ("-r", 5), ("-f", 3), ("-f", 1)])
def test_split_by_reads_and_files(fx, tmp_path, how, count):
    src = many_reads(fx, tmp_path / "eleven.slow5")
    recs = records(src)
    n = len(recs)
    out = tmp_path / "o"
    r = run(["split", how, str(count), src, "-d", str(out), "--to", "slow5", "-K", "3"])
    assert r.returncode == 0, r.stderr.decode()
    files = sorted(os.listdir(out), key=lambda f: int(f.rsplit("_", 1)[1].split(".")[0]))
    got = [records(out / f) for f in files]
    if how == "-r":
        assert [len(g) for g in got] == [count] * (n // count) + ([n % count] if n % count else [])
    else:  # the first n % count files get one record more (split.c:379-400)
        assert [len(g) for g in got] == [n // count + (1 if i < n % count else 0) for i in range(count)]
    assert [l for g in got for l in g] == recs
    header = [l for l in open(src, "rb").read().split(b"\n") if l[:1] in (b"#", b"@")]
    for f in files:
        assert [l for l in open(out / f, "rb").read().split(b"\n") if l[:1] in (b"#", b"@")] == header


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/slow5tools_ref not built")
def test_merge_and_split_against_reference_binary(fx, tmp_path):
    for k, (flags, inputs) in enumerate([([], ["rg0.slow5", "aux_enum.slow5"]), (["--lossless", "false"], ["rg0.slow5", "rg0_1.slow5"]),
                                         ([], ["aux_enum.slow5", "rg0_2_aux_order.slow5", "rg1_1_new_aux_field.slow5"])]):
        a, b = tmp_path / ("ref%d.slow5" % k), tmp_path / ("ours%d.slow5" % k)
        subprocess.check_call([REF, "merge"] + flags + raw(fx, *inputs) + ["-o", str(a)], stderr=subprocess.DEVNULL, timeout=60)
        assert run(["merge"] + flags + raw(fx, *inputs) + ["-o", str(b)]).returncode == 0
        assert filecmp.cmp(a, b, shallow=False), (flags, inputs)
        # uncompressed BLOW5 output is byte-identical as well
        a2, b2 = tmp_path / ("ref%d.blow5" % k), tmp_path / ("ours%d.blow5" % k)
        subprocess.check_call([REF, "merge"] + flags + raw(fx, *inputs) + ["-o", str(a2), "-c", "none", "-s", "none"], stderr=subprocess.DEVNULL, timeout=60)
        assert run(["merge"] + flags + raw(fx, *inputs) + ["-o", str(b2), "-c", "none", "-s", "none"]).returncode == 0
        assert filecmp.cmp(a2, b2, shallow=False), (flags, inputs)
    merged = str(tmp_path / "ref2.slow5")
    for k, how in enumerate((["-g"], ["-g", "-l", "false"])):
        da, db = tmp_path / ("sref%d" % k), tmp_path / ("sours%d" % k)
        subprocess.check_call([REF, "split"] + how + [merged, "-d", str(da), "--to", "slow5"], stderr=subprocess.DEVNULL, timeout=60)
        assert run(["split"] + how + [merged, "-d", str(db), "--to", "slow5"]).returncode == 0
        assert sorted(os.listdir(da)) == sorted(os.listdir(db))
        for f in os.listdir(da):
            assert filecmp.cmp(da / f, db / f, shallow=False), f
    # (never more files than records: the reference then creates empty files without end, split.c:379-456)
    single = many_reads(fx, tmp_path / "eleven.slow5")
    for k, how in enumerate((["-r", "3"], ["-f", "4"], ["-r", "3", "--to", "blow5", "-c", "none", "-s", "none"])):
        da, db = tmp_path / ("rref%d" % k), tmp_path / ("rours%d" % k)
        tail = [] if "--to" in how else ["--to", "slow5"]
        subprocess.check_call([REF, "split"] + how + [single, "-d", str(da)] + tail, stderr=subprocess.DEVNULL, timeout=60)
        assert run(["split"] + how + [single, "-d", str(db)] + tail).returncode == 0
        assert sorted(os.listdir(da)) == sorted(os.listdir(db))
        for f in os.listdir(da):
            assert filecmp.cmp(da / f, db / f, shallow=False), f


# ---- compressed inputs / outputs: the codec runs on the GPU -------------------------------------------------------------------
@pytest.mark.gpu
def test_merge_mixed_formats_golden(fx, tmp_path):
    """test_merge.sh 1.5: SLOW5 text, BLOW5 v0.1.0 uncompressed, zlib + svb-zd v0.2.0 and zlib v0.2.0 inputs in one merge."""
    out = tmp_path / "out.slow5"
    r = run(["merge"] + raw(fx, "aux_no_enum.slow5", "none_v0.1.0.blow5", "zlib_svb-zd_v0.2.0.blow5", "zlib_v0.2.0.blow5") + ["-o", str(out)])
    assert r.returncode == 0, r.stderr.decode()
    assert filecmp.cmp(out, os.path.join(fx, "merge", "exp", "merged_output_formats.slow5"), shallow=False)


@pytest.mark.gpu
@pytest.mark.parametrize("rec,sig", [("zlib", "svb-zd"), ("zstd", "ex-zd"), ("zlib", "none")])
def test_merge_to_compressed_blow5_round_trip(fx, tmp_path, rec, sig):
    b = tmp_path / "m.blow5"
    r = run(["merge"] + raw(fx, "rg0.slow5", "rg0_2_aux_order.slow5") + ["-o", str(b), "-c", rec, "-s", sig])
    assert r.returncode == 0, r.stderr.decode()
    s = tmp_path / "m.slow5"
    assert run(["view", str(b), "-o", str(s)]).returncode == 0
    assert filecmp.cmp(s, os.path.join(fx, "merge", "exp", "same_rg_aux_order.slow5"), shallow=False)
    # and a compressed file as merge input again
    s2 = tmp_path / "m2.slow5"
    assert run(["merge", str(b), "--to", "slow5", "-o", str(s2)]).returncode == 0
    assert filecmp.cmp(s2, s, shallow=False)


@pytest.mark.gpu
def test_split_groups_of_compressed_blow5_golden(fx, tmp_path):
    """test_split.sh testcase 4: a zlib BLOW5 v0.1.0 file with five read groups."""
    out = tmp_path / "g"
    r = run(["split", "-g", os.path.join(fx, "split", "raw", "example_multi_rg_v0.1.0.blow5"), "-d", str(out), "--to", "slow5"])
    assert r.returncode == 0, r.stderr.decode()
    exp = os.path.join(fx, "split", "exp", "expected_group_split_blow5_input")
    assert sorted(os.listdir(out)) == sorted(os.listdir(exp))
    for f in os.listdir(exp):
        assert filecmp.cmp(out / f, os.path.join(exp, f), shallow=False), f


@pytest.mark.gpu
def test_split_to_compressed_outputs_round_trip(fx, tmp_path):
    src = os.path.join(fx, "split", "raw", "example_multi_rg_v0.1.0.blow5")
    out = tmp_path / "g"
    assert run(["split", "-g", src, "-d", str(out), "-c", "zlib", "-s", "svb-zd"]).returncode == 0
    exp = os.path.join(fx, "split", "exp", "expected_group_split_blow5_input")
    for f in sorted(os.listdir(out)):
        assert f.endswith(".blow5")
        s = tmp_path / (f + ".slow5")
        assert run(["view", str(out / f), "-o", str(s)]).returncode == 0
        assert filecmp.cmp(s, os.path.join(exp, f.replace(".blow5", ".slow5")), shallow=False), f


@pytest.mark.gpu
@pytest.mark.parametrize("rec,sig", [("zlib", "svb-zd"), ("none", "svb-zd"), ("zlib", "none")])
def test_merge_of_blow5_inputs_on_the_device_path(fx, tmp_path, rec, sig):
    """BLOW5 inputs whose auxiliary columns need no re-laying stay on the device (read groups renumbered by rec_rg_remap_kernel):
    same records as the host record loop (S5B_VIEW_SLOW_PATH=1) and as the merge of the SLOW5 originals; the reference reads
    the result."""
    a_txt = os.path.join(fx, "merge", "raw", "rg0.slow5")
    # a second input with the same columns but another run_id: a new read group, 0 -> 1
    b_txt = tmp_path / "other.slow5"
    lines = open(os.path.join(fx, "merge", "raw", "rg0_1.slow5"), "rb").read().split(b"\n")
    b_txt.write_bytes(b"\n".join(l + b"_other" if l.startswith(b"@run_id\t") else l for l in lines))
    want = tmp_path / "want.slow5"
    assert run(["merge", a_txt, str(b_txt), "-o", str(want)]).returncode == 0
    assert b"#num_read_groups\t2" in open(want, "rb").read()
    ins = []
    for k, t in enumerate((a_txt, str(b_txt))):
        b = tmp_path / ("in%d.blow5" % k)
        assert run(["view", t, "-o", str(b), "-c", rec, "-s", sig]).returncode == 0
        ins.append(str(b))
    outs = {}
    for name, env in (("device", {}), ("host", {"S5B_VIEW_SLOW_PATH": "1"})):
        out = tmp_path / (name + ".blow5")
        r = subprocess.run([CLI, "merge"] + ins + ["-o", str(out), "-c", "zlib", "-s", "svb-zd"], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, env=dict(os.environ, **env), timeout=120)
        assert r.returncode == 0, r.stderr.decode()
        txt = tmp_path / (name + ".slow5")
        assert run(["view", str(out), "-o", str(txt)]).returncode == 0
        outs[name] = txt
        if os.path.exists(REF):
            rtxt = tmp_path / (name + ".ref.slow5")
            subprocess.check_call([REF, "view", str(out), "-o", str(rtxt)], stderr=subprocess.DEVNULL, timeout=60)
            assert filecmp.cmp(rtxt, want, shallow=False)
    assert filecmp.cmp(outs["device"], want, shallow=False)
    assert filecmp.cmp(outs["host"], want, shallow=False)
    # same methods in and out: the records still have to be rewritten (their read groups change)
    same = tmp_path / "same.blow5"
    assert run(["merge"] + ins + ["-o", str(same), "-c", rec, "-s", sig]).returncode == 0
    stxt = tmp_path / "same.slow5"
    assert run(["view", str(same), "-o", str(stxt)]).returncode == 0
    assert filecmp.cmp(stxt, want, shallow=False)
    # a record whose read group is not in its file's header stops the merge
    data = bytearray(open(ins[1], "rb").read())
    if rec == "none":
        hsize = int.from_bytes(data[64:68], "little")
        at = 68 + hsize + 8
        idl = int.from_bytes(data[at:at + 2], "little")
        data[at + 2 + idl:at + 2 + idl + 4] = (7).to_bytes(4, "little")
        bad = tmp_path / "bad.blow5"
        bad.write_bytes(bytes(data))
        r = run(["merge", ins[0], str(bad), "-o", str(tmp_path / "x.blow5")])
        assert r.returncode != 0
