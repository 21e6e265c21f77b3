"""The pipelined whole-batch record transcoder (s5b_blow5_recode_batch_host / s5b_blow5_recode_dev, recode_engine.cu)
against the oracle's restatement of view's per-record worker (oracle/blow5_oracle.c, pinned to the compiled reference by
tests/test_oracle_blow5.py): every record of every batch is compared, not a sample."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import build_oracle  # noqa: E402
from recode_helpers import (M_EX_ZD, M_NONE, M_SVB_ZD, M_ZLIB, M_ZSTD, REF_SO, RecordOracle, make_records, slab,  # noqa: E402
                            walk_image)

have_ref = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ro():
    return RecordOracle(build_oracle())


@pytest.fixture(scope="module")
def codecs():
    """two contexts: default chunking, and 37-record chunks so that small batches exercise the 3-lane pipeline of the host form
    and the running image offset of the device-resident form"""
    import slow5tools_b200 as s5
    big = s5.Codec(0)
    os.environ["S5B_RECODE_CHUNK"] = "37"
    os.environ["S5B_RECODE_DEV_CHUNK"] = "37"   # (the device-resident form sizes its chunks from the free memory otherwise)
    small = s5.Codec(0)
    del os.environ["S5B_RECODE_CHUNK"]
    del os.environ["S5B_RECODE_DEV_CHUNK"]
    yield big, small
    big.close()
    small.close()


def host_recode(cdc, methods, records, gap=0, align=1, cap=None, want_off=True):
    buf, off, ln, used = slab(records, align=align, gap=gap)
    cap = cap if cap is not None else int(ln.sum()) * 5 + 1100 * len(records) + 4096
    out = np.zeros(cap, np.uint8)
    img_off = np.zeros(len(records) + 1, np.uint64) if want_off else None
    rc, nb = cdc.blow5_recode_batch_host(*methods, buf, used, off, ln, out, img_off, check=False)
    return rc, out[:nb].tobytes() if rc == 0 else nb, img_off


def dev_recode(cdc, methods, records, gap=0, cap=None):
    import torch
    buf, off, ln, used = slab(records, align=1, gap=gap)
    d_in = torch.from_numpy(np.concatenate([buf, np.zeros(32, np.uint8)])).cuda()
    cap = cap if cap is not None else int(ln.sum()) * 5 + 1100 * len(records) + 4096
    d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_off = torch.zeros(len(records) + 1, dtype=torch.int64, device="cuda")
    cdc.blow5_recode_dev(*methods, d_in, used, off, ln, d_out, d_res, d_off)
    cdc.sync()
    res = d_res.cpu().numpy()
    return int(res[1]), d_out[:int(res[0])].cpu().numpy().tobytes(), d_off.cpu().numpy().view(np.uint64)


LENS = [4096] * 90 + [0, 1, 2, 3, 4, 5, 31, 32, 33, 255, 256, 257, 1023, 1025, 30000, 70001, 8, 9, 4095, 4097] + [4096] * 60


def expect(ro, methods, records):
    out = []
    for r in records:
        rc, e = ro.recode(methods, r)
        assert rc == 0
        out.append(e[8:])
    return out


@pytest.mark.parametrize("which", [0, 1])
def test_encode_every_record_matches_oracle(ro, codecs, which):
    """none/none -> none/svb-zd is fully deterministic: the image must equal the oracle's byte for byte; with zlib on top
    the records must inflate (system zlib) to exactly the oracle's packed records."""
    cdc = codecs[which]
    recs, _ = make_records(LENS, seed=11, aux=b"")
    want = expect(ro, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs)
    rc, img, off = host_recode(cdc, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs, gap=8)
    assert rc == 0
    got = walk_image(img)
    assert got == want
    assert int(off[-1]) == len(img) and all(int(off[i + 1] - off[i]) == 8 + len(want[i]) for i in range(len(want)))
    rc, img, off = host_recode(cdc, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs, gap=3)
    assert rc == 0
    got = walk_image(img)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert zlib.decompress(g) == w
    total, ztotal = sum(len(w) for w in want), sum(len(g) for g in got)
    zref = sum(len(zlib.compress(w, 6)) for w in want)
    assert ztotal <= 1.03 * zref, (ztotal, zref)          # size tolerance of the record codec
    assert int(off[-1]) == len(img)


@pytest.mark.parametrize("which", [0, 1])
def test_decode_every_record_matches_original(ro, codecs, which):
    """records compressed by the ORACLE (system zlib + svb-zd, byte-identical to the reference's) -> our decode"""
    cdc = codecs[which]
    recs, _ = make_records(LENS, seed=12, aux=b"\x07" * 11)
    stored = expect(ro, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs)
    rc, img, off = host_recode(cdc, (M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), stored, gap=8)
    assert rc == 0
    assert walk_image(img) == recs
    # signal-only and record-only conversions, both ways
    rc, img, _ = host_recode(cdc, (M_ZLIB, M_SVB_ZD, M_NONE, M_SVB_ZD), stored)
    assert rc == 0 and walk_image(img) == expect(ro, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs)
    rc, img, _ = host_recode(cdc, (M_ZLIB, M_SVB_ZD, M_ZLIB, M_SVB_ZD), stored)
    assert rc == 0 and walk_image(img) == stored             # nothing to do: stored records pass through


def test_device_form_equals_host_form(ro, codecs):
    cdc = codecs[1]
    recs, _ = make_records(LENS, seed=13)
    for methods in [(M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), (M_NONE, M_NONE, M_NONE, M_EX_ZD), (M_NONE, M_NONE, M_ZSTD, M_SVB_ZD)]:
        use = [r for r in recs if not (methods[3] == M_EX_ZD and len(r) < 2 + 18 + 44 + 4)]
        rc, img, off = host_recode(cdc, methods, use, gap=8)
        assert rc == 0
        err, dimg, doff = dev_recode(cdc, methods, use, gap=8)
        assert err == 0 and dimg == img and np.array_equal(doff, off)
        back = (methods[2], methods[3], M_NONE, M_NONE)
        stored = walk_image(img)
        rc, raw, _ = host_recode(cdc, back, stored)
        err, draw, _ = dev_recode(cdc, back, stored)
        assert rc == 0 and err == 0 and raw == draw and walk_image(raw) == use
    # a workspace budget far below what one chunk over the batch needs: the default context cuts the pass into many chunks
    big = codecs[0]
    big.set_recode_workspace(1 << 20)
    err, dimg2, doff2 = dev_recode(big, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs, gap=8)
    big.set_recode_workspace(0)
    rc, img, off = host_recode(big, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs, gap=8)
    assert err == 0 and rc == 0 and dimg2 == img and np.array_equal(doff2, off)


def test_exzd_and_zstd_records(ro, codecs):
    cdc = codecs[1]
    lens = [n for n in LENS if n >= 2]
    recs, _ = make_records(lens, seed=14)
    want = expect(ro, (M_NONE, M_NONE, M_NONE, M_EX_ZD), recs)
    rc, img, _ = host_recode(cdc, (M_NONE, M_NONE, M_NONE, M_EX_ZD), recs)
    assert rc == 0 and walk_image(img) == want
    rc, img2, _ = host_recode(cdc, (M_NONE, M_EX_ZD, M_ZLIB, M_SVB_ZD), want)
    assert rc == 0
    assert [zlib.decompress(g) for g in walk_image(img2)] == expect(ro, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs)


def test_inflate_slot_overflow_takes_the_careful_path(ro, codecs):
    """a constant signal compresses far beyond the 4x + 1 KiB slot the fast path gives a record to inflate into: the host
    form settles that chunk with the careful transcoder, the device form reports S5B_ERR_NOSPACE"""
    cdc = codecs[1]
    lens = [4096] * 50 + [60000] + [4096] * 50
    recs, _ = make_records(lens, seed=15, constant={50})
    stored = expect(ro, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs)
    assert len(stored[50]) * 4 + 1024 < 60000
    rc, img, off = host_recode(cdc, (M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), stored)
    assert rc == 0 and walk_image(img) == recs
    assert int(off[-1]) == len(img) and [int(off[i + 1] - off[i]) for i in range(len(recs))] == [len(r) + 8 for r in recs]
    err, _, _ = dev_recode(cdc, (M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), stored)
    assert err == -40


def test_malformed_record_verdicts(ro, codecs):
    cdc = codecs[1]
    recs, _ = make_records([4096] * 120, seed=16)
    stored = expect(ro, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs)
    bad = list(stored)
    b = bytearray(bad[77])
    b[len(b) // 2] ^= 0x10
    bad[77] = bytes(b)
    rc, _, _ = host_recode(cdc, (M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), bad)
    assert rc == -13
    err, _, _ = dev_recode(cdc, (M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), bad)
    assert err == -13
    # an svb-zd stream that claims more samples than it has bytes
    plain = expect(ro, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs[:40])
    p = bytearray(plain[5])
    idlen = int.from_bytes(p[:2], "little")
    at = 2 + idlen + 4 + 32 + 8
    p[at:at + 4] = (10 ** 9).to_bytes(4, "little")
    plain[5] = bytes(p)
    rc, _, _ = host_recode(cdc, (M_NONE, M_SVB_ZD, M_NONE, M_NONE), plain)
    assert rc == -13
    # truncated fixed fields
    cut = list(recs[:40])
    cut[3] = cut[3][:20]
    rc, _, _ = host_recode(cdc, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), cut)
    assert rc == -13


def test_output_capacity(ro, codecs):
    cdc = codecs[1]
    recs, _ = make_records([4096] * 100, seed=17)
    rc, img, _ = host_recode(cdc, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs)
    assert rc == 0
    rc, need, _ = host_recode(cdc, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs, cap=len(img) - 1)
    assert rc == -40 and need == len(img)
    err, _, _ = dev_recode(cdc, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs, cap=len(img) - 1)
    assert err == -40
    rc, img2, _ = host_recode(cdc, (M_NONE, M_NONE, M_NONE, M_SVB_ZD), recs, cap=len(img))
    assert rc == 0 and img2 == img


@have_ref
def test_reference_reads_our_records_and_we_read_its(ro, codecs):
    """the compiled reference (slow5_decode + slow5_encode) on OUR compressed records and vice versa, every record"""
    cdc = codecs[0]
    recs, _ = make_records([4096] * 300 + [123, 7000, 65536], seed=18)
    rc, img, _ = host_recode(cdc, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs)
    assert rc == 0
    rc, back = ro.batch((M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), walk_image(img), use_ref=True)
    assert rc == 0 and walk_image(back) == recs
    rc, theirs = ro.batch((M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs, use_ref=True)
    assert rc == 0
    rc, mine, _ = host_recode(cdc, (M_ZLIB, M_SVB_ZD, M_NONE, M_NONE), walk_image(theirs))
    assert rc == 0 and walk_image(mine) == recs


def test_stage_timing_reports(codecs):
    cdc = codecs[1]
    recs, _ = make_records([4096] * 200, seed=19)
    cdc.stage_timing(True)
    cdc.stage_report(reset=True)
    rc, img, _ = host_recode(cdc, (M_NONE, M_NONE, M_ZLIB, M_SVB_ZD), recs)
    rep = cdc.stage_report(reset=True)
    cdc.stage_timing(False)
    assert rc == 0
    assert rep["record_press"][1] >= 5 and rep["record_press"][0] > 0 and rep["signal_press"][0] > 0 and rep["d2h"][1] >= 5
    assert rep["record_depress"][1] == 0
