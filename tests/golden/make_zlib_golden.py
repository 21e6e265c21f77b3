"""Cuts real zlib-compressed records out of the reference's own BLOW5 fixtures (read-only tree under
/root/reference) and stores them in tests/golden/zlib_records.npz (the tests inflate them with system zlib, the reference's
own dependency, as the oracle).  Also stores the reference's press golden
slow5lib/test/data/exp/unit_test_exp_press (4 concatenated zlib streams).  Build container only.

    python tests/golden/make_zlib_golden.py
"""
import os
import struct
import zlib

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = [
    ("merge", "test/data/raw/merge/zlib_svb-zd_v0.2.0.blow5", 1),
    ("lossless", "test/data/exp/one_fast5/exp_1_lossless_zlib_svb_v0.2.0.blow5", 1),
    ("lossy_zlib_only", "test/data/exp/one_fast5/exp_1_lossy_zlib.blow5", 1),
    ("multi_rg", "test/data/raw/stats/zlib_svb-zd_multi_rg_v0.2.0.blow5", 7),
    ("example3", "slow5lib/examples/adv/example3.blow5", 10),
    ("multi100", "test/data/exp/f2s/retain_dir_structure/single_small_multifast5/multi.blow5", 4),
]


def records(path, limit):
    b = open(path, "rb").read()
    assert b[:6] == b"BLOW5\x01", path
    rec_method = b[9]
    assert rec_method == 1, (path, rec_method)     # zlib (slow5_press.c:58-104 file byte map)
    hdr_size = struct.unpack_from("<I", b, 64)[0]
    pos = 68 + hdr_size
    out = []
    while len(out) < limit and b[pos:pos + 5] != b"5WOLB":
        (size,) = struct.unpack_from("<Q", b, pos)
        out.append(b[pos + 8:pos + 8 + size])
        pos += 8 + size
    return out


def main():
    d = {}
    n = 0
    for name, rel, limit in FIXTURES:
        path = os.path.join(REF, rel)
        if not os.path.exists(path):
            print("skip", rel)
            continue
        for i, rec in enumerate(records(path, limit)):
            d["z__%s_%d" % (name, i)] = np.frombuffer(rec, np.uint8)
            zlib.decompress(rec)                      # must be a complete, valid stream
            n += 1
    d["unit_test_exp_press"] = np.frombuffer(
        open(os.path.join(REF, "slow5lib/test/data/exp/unit_test_exp_press"), "rb").read(), np.uint8)
    out = os.path.join(HERE, "zlib_records.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, n, "records", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
