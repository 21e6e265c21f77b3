"""`slow5tools-b200 index` against the reference binary (`slow5tools index`, src/index.c + slow5lib/src/slow5_idx.c):
the .idx file must be byte-identical.  Uncompressed BLOW5 and SLOW5 text need no GPU; zlib / zstd records do."""
import filecmp
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")
have_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/slow5tools_ref not present")


def both_indexes(tmp_path, src_file):
    a, b = tmp_path / "ours" / os.path.basename(src_file), tmp_path / "ref" / os.path.basename(src_file)
    for p in (a, b):
        os.makedirs(p.parent, exist_ok=True)
        shutil.copy(src_file, p)
    r = subprocess.run([CLI, "index", str(a)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    subprocess.check_call([REF, "index", str(b)], stderr=subprocess.DEVNULL)
    return str(a) + ".idx", str(b) + ".idx"


def parse_idx(path):
    b = open(path, "rb").read()
    assert b[:9] == b"SLOW5IDX\x01" and b[-8:] == b"XDI5WOLS"
    pos, out = 64, []
    while pos < len(b) - 8:
        n = struct.unpack_from("<H", b, pos)[0]
        rid = b[pos + 2:pos + 2 + n]
        off, size = struct.unpack_from("<QQ", b, pos + 2 + n)
        out.append((rid, off, size))
        pos += 2 + n + 16
    return b[9:12], out


@have_ref
@pytest.mark.parametrize("name", ["exp_1_lossless.blow5", "exp_1_lossless.slow5", "exp_1_lossless_v0.2.0.slow5"])
def test_uncompressed_and_text_files_match_the_reference(tmp_path, name):
    mine, theirs = both_indexes(tmp_path, os.path.join(FIX, name))
    assert filecmp.cmp(mine, theirs, shallow=False)
    ver, entries = parse_idx(mine)
    assert len(entries) >= 1 and all(size > 0 for _, _, size in entries)


def test_failure_cases(tmp_path):
    r = subprocess.run([CLI, "index"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0
    r = subprocess.run([CLI, "index", str(tmp_path / "missing.blow5")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0
    # truncated file: no end-of-file marker (slow5_idx.c:264-272)
    src = open(os.path.join(FIX, "exp_1_lossless.blow5"), "rb").read()
    cut = tmp_path / "cut.blow5"
    open(cut, "wb").write(src[:-5])
    r = subprocess.run([CLI, "index", str(cut)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and not os.path.exists(str(cut) + ".idx")


@pytest.mark.gpu
@have_ref
@pytest.mark.parametrize("name", ["exp_1_lossless_zlib_svb_v0.2.0.blow5", "exp_1_lossless_zstd_svb_v0.2.0.blow5",
                                  "exp_1_lossless_zstd_v0.2.0.blow5", "exp_1_lossless_zlib_ex_zd.blow5", "exp_1_lossy_zlib.blow5",
                                  "zlib_svb-zd_multi_rg_v0.2.0.blow5", "zlib_svb-zd_v0.2.0.blow5"])
def test_compressed_fixtures_match_the_reference(tmp_path, name):
    mine, theirs = both_indexes(tmp_path, os.path.join(FIX, name))
    assert filecmp.cmp(mine, theirs, shallow=False)


@pytest.mark.gpu
@have_ref
@pytest.mark.parametrize("method", ["zlib", "zstd"])
def test_many_records_and_long_ids(tmp_path, method):
    """6000 records in several chunks, with read ids from 1 byte to 700 bytes: the long ones do not fit in the 256-byte
    prefix that is decompressed first (slow5_idx.c:290-320) and take the full-decompression path."""
    rng = np.random.default_rng(9)
    text = ("@asic_id\t0\n#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n"
            "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n").encode()
    raw = tmp_path / "raw.blow5"
    with open(raw, "wb") as f:
        hdr = bytearray(68)
        hdr[0:6] = b"BLOW5\x01"
        hdr[6:9] = bytes([0, 2, 0])
        hdr[10:14] = struct.pack("<I", 1)
        hdr[64:68] = struct.pack("<I", len(text))
        f.write(hdr)
        f.write(text)
        fixed = struct.pack("<I4d", 0, 8192.0, 9.0, 1444.86, 4000.0)
        for r in range(6000):
            if r % 500 == 7:
                rid = ("long-%d-" % r).encode() + bytes(rng.integers(97, 123, int(rng.integers(300, 700)), dtype=np.uint8))
            elif r % 500 == 8:
                rid = b"%d" % r  # short ids
            else:
                rid = ("%08x-%04x-%04x-%04x-%012x" % (r, r & 0xffff, 7, 9, r * 7919)).encode()
            n = int(rng.integers(1, 5000))
            sig = (500 + rng.integers(-40, 40, n)).astype(np.int16)
            body = struct.pack("<H", len(rid)) + rid + fixed + struct.pack("<Q", n) + sig.tobytes()
            f.write(struct.pack("<Q", len(body)))
            f.write(body)
        f.write(b"5WOLB")
    z = tmp_path / ("z_%s.blow5" % method)
    subprocess.check_call([REF, "view", str(raw), "-c", method, "-s", "svb-zd", "-o", str(z)], stderr=subprocess.DEVNULL)
    mine, theirs = both_indexes(tmp_path, str(z))
    assert filecmp.cmp(mine, theirs, shallow=False)
    ver, entries = parse_idx(mine)
    assert len(entries) == 6000 and max(len(e[0]) for e in entries) > 300
