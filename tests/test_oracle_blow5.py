"""Pins oracle/blow5_oracle.c (the CPU restatement of view's per-record worker, slow5.c:2580-2950 / :3928-4074) against
the compiled reference run on the same records (slow5_decode + slow5_encode through oracle/ref_driver.c) and against
records cut from the reference's own BLOW5 fixtures."""
import os
import zlib

import numpy as np
import pytest

from conftest import build_oracle
from recode_helpers import (M_EX_ZD, M_NONE, M_SVB_ZD, M_ZLIB, REF_SO, RecordOracle, make_records, walk_image)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
have_ref = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ro():
    return RecordOracle(build_oracle())


LENGTHS = [0, 1, 2, 3, 4, 5, 31, 32, 33, 255, 256, 257, 1000, 4096, 4097, 10000, 70001]


@have_ref
@pytest.mark.parametrize("out_m", [(M_ZLIB, M_SVB_ZD), (M_NONE, M_SVB_ZD), (M_ZLIB, M_NONE), (M_NONE, M_EX_ZD), (M_ZLIB, M_EX_ZD)])
def test_port_matches_reference_bytes(ro, out_m):
    """same system zlib, same svb-zd / ex-zd bytes: the restatement's output equals the reference's, byte for byte"""
    # ex-zd: an empty read is undefined in the reference and a 1-sample read makes it free() an invalid pointer
    lens = [n for n in LENGTHS if not (out_m[1] == M_EX_ZD and n < 2)]
    recs, _ = make_records(lens, seed=3)
    rc_r, img_r = ro.batch((M_NONE, M_NONE) + out_m, recs, use_ref=True)
    rc_p, img_p = ro.batch((M_NONE, M_NONE) + out_m, recs, use_ref=False)
    assert rc_r == 0 and rc_p == 0
    assert img_r == img_p
    # and back: both decode the reference's image to the original records
    stored = walk_image(img_r)
    rc_r, back_r = ro.batch(out_m + (M_NONE, M_NONE), stored, use_ref=True)
    rc_p, back_p = ro.batch(out_m + (M_NONE, M_NONE), stored, use_ref=False)
    assert rc_r == 0 and rc_p == 0 and back_r == back_p
    assert walk_image(back_r) == recs


def test_port_round_trip_and_layout(ro):
    recs, sigs = make_records([0, 7, 4096, 12345], seed=5, aux=b"\x01\x02\x03\x04\x05")
    for r, s in zip(recs, sigs):
        rc, enc = ro.recode((M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), r)
        assert rc == 0
        assert int.from_bytes(enc[:8], "little") == len(enc) - 8
        packed = zlib.decompress(enc[8:])
        idlen = int.from_bytes(packed[:2], "little")
        head = 2 + idlen + 4 + 32
        assert packed[:head] == r[:head]
        nbytes = int.from_bytes(packed[head:head + 8], "little")             # compressed signal: byte count (slow5.c:3983)
        assert int.from_bytes(packed[head + 8:head + 12], "little") == len(s)  # svb-zd header = sample count
        assert packed[head + 8 + nbytes:] == b"\x01\x02\x03\x04\x05"          # aux carried over
        rc, back = ro.recode((M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), enc[8:])
        assert rc == 0 and back[8:] == r


def test_port_rejects_malformed(ro):
    recs, _ = make_records([500], seed=9)
    rc, enc = ro.recode((M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs[0])
    bad = bytearray(enc[8:])
    bad[len(bad) // 2] ^= 0x40
    rc, _ = ro.recode((M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), bytes(bad))
    assert rc != 0
    rc, _ = ro.recode((M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs[0][:40])      # cut inside the fixed fields
    assert rc == -4   # SLOW5_ERR_RECPARSE, slow5_defs.h:140


def test_fixture_records(ro):
    """records cut from the reference's own zlib + svb-zd file (tests/golden/zlib_records.npz): decode, re-encode, decode"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "zlib_records.npz"))
    names = [k for k in z.files if k.startswith("z__merge") or k.startswith("z__multi_rg") or k.startswith("z__example3")]
    if not names:
        pytest.skip("no record fixtures in zlib_records.npz")
    for k in names[:8]:
        stored = z[k].tobytes()
        rc, raw = ro.recode((M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), stored)
        assert rc == 0
        rc, again = ro.recode((M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), raw[8:])
        assert rc == 0 and again[8:] == stored                                 # same zlib, same svb-zd: same bytes
