"""BASELINE.json's configurations as parity cases (sizes scaled where the full size needs hours of reference CPU time;
the full-size runs check size-independent properties: exact round trips, stream trailers, reference-readable output).

  configs[2]  full BLOW5 encode (svb-zd + zlib), 1M reads x 4096 int16, 1 GPU
  configs[3]  BLOW5 decode -> re-encode, lognormal read lengths 500-200k samples, byte-balanced shards (8 ranks)
  configs[4]  svb-zd + zstd path, 4096-sample reads
"""
import ctypes as C
import os
import struct
import subprocess
import sys
import zlib

import numpy as np
import pytest
import torch

import slow5tools_b200 as s5
from slow5tools_b200 import synth
from slow5tools_b200._capi import METHOD
from slow5tools_b200.dist import shard_bounds

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
CLI = os.path.join(ROOT, "slow5tools_b200", "bin", "slow5tools-b200")
REF = os.path.join(ROOT, "oracle", "_ref", "slow5tools_ref")
have_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/slow5tools_ref not present")


@pytest.fixture(scope="module")
def cdc():
    c = s5.Codec(0)
    yield c
    c.close()


def _full_record_roundtrip(cdc, R, N, entropy):
    """svb-zd encode -> zlib|zstd encode -> decode -> svb-zd decode, device resident, in slabs of 100k reads."""
    slab = 100000
    slot = int(s5.lib.s5b_svbzd_slot(N))
    zslot = int(s5.lib.s5b_zlib_bound(slot) if entropy == "zlib" else s5.lib.s5b_zstd_bound(slot))
    n = torch.full((slab,), N, dtype=torch.int32, device="cuda")
    soff = torch.arange(slab + 1, dtype=torch.int64, device="cuda") * N
    ooff = torch.arange(slab + 1, dtype=torch.int64, device="cuda") * slot
    zoff = torch.arange(slab + 1, dtype=torch.int64, device="cuda") * zslot
    split = torch.full((slab,), 4 + (N + 3) // 4, dtype=torch.int32, device="cuda")
    svb = torch.zeros(slab * slot + 16, dtype=torch.uint8, device="cuda")
    svb2 = torch.zeros_like(svb)
    zbuf = torch.zeros(slab * zslot + 16, dtype=torch.uint8, device="cuda")
    l1, l2, zl = (torch.zeros(slab, dtype=torch.int32, device="cuda") for _ in range(3))
    st = [torch.ones(slab, dtype=torch.int32, device="cuda") for _ in range(4)]
    n2 = torch.zeros_like(n)
    raw = comp = svb_total = 0
    enc = cdc.zlib_deflate_dev if entropy == "zlib" else cdc.zstd_encode_dev
    dec = cdc.zlib_inflate_dev if entropy == "zlib" else cdc.zstd_decode_dev
    sample = None
    for k in range(R // slab):
        sig = synth.nanopore_signal(slab * N, seed=1000 + k, device="cuda")
        back = torch.zeros_like(sig)
        cdc.svbzd_encode_dev(sig, soff, n, svb, ooff, l1, st[0])
        enc(svb, ooff, l1, zbuf, zoff, zl, st[1], split=split)
        dec(zbuf, zoff, zl, svb2, ooff, l2, st[2])
        cdc.svbzd_decode_dev(svb2, ooff, l2, back, soff, n2, st[3])
        torch.cuda.synchronize()
        assert all(int(s.abs().sum()) == 0 for s in st), k
        assert torch.equal(l1, l2) and torch.equal(back, sig) and torch.equal(n2, n), k
        raw += slab * N * 2
        svb_total += int(l1.sum())
        comp += int(zl.sum())
        if k == 0:
            zl_h, l1_h = zl[:64].cpu().numpy(), l1[:64].cpu().numpy()
            sample = [(zbuf[i * zslot:i * zslot + int(zl_h[i])].cpu().numpy().tobytes(),
                       svb[i * slot:i * slot + int(l1_h[i])].cpu().numpy().tobytes()) for i in range(64)]
        del sig, back
    return raw, svb_total, comp, sample


def test_config2_full_encode_svbzd_zlib_1m_reads(cdc):
    raw, svb_total, comp, sample = _full_record_roundtrip(cdc, 1000000, 4096, "zlib")
    assert 1.24 < svb_total / (raw / 2) < 1.30          # SURVEY 8d calibration: svb-zd 1.26 +- 0.03 B/sample
    assert 0.64 < comp / svb_total < 0.70               # zlib-6 on the same bytes: 0.68 +- 0.03
    ours = theirs = 0
    for z, plain in sample:                             # any zlib reads our streams; size within 3 % of level 6
        assert z[:2] == b"\x78\x9c" and zlib.decompress(z) == plain
        assert struct.unpack(">I", z[-4:])[0] == zlib.adler32(plain)
        ours += len(z)
        theirs += len(zlib.compress(plain, 6))
    assert ours <= 1.03 * theirs


def test_config4_svbzd_zstd_1m_reads(cdc):
    raw, svb_total, comp, sample = _full_record_roundtrip(cdc, 1000000, 4096, "zstd")
    assert 0.64 < comp / svb_total < 0.70
    for z, plain in sample:
        assert z[:4] == b"\x28\xb5\x2f\xfd"            # one zstd frame per record, content size in the header
        nb = C.c_uint64()
        assert s5.lib.s5b_zstd_content_size(z, len(z), C.byref(nb)) == 0 and nb.value == len(plain)


def _lognormal_file(path, n_reads, seed):
    """none/none BLOW5 with lognormal read lengths (BASELINE configs[3]: 500 - 200k samples)."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.round(rng.lognormal(np.log(30000), 1.0, n_reads)), 500, 200000).astype(np.int64)
    sig = synth.nanopore_signal(int(lens.sum()), seed=seed).numpy()
    text = ("@asic_id\t0\n@exp_start_time\t2026-01-01T00:00:00Z\n@flow_cell_id\tSYNTH\n@sample_frequency\t4000\n"
            "#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n"
            "#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n").encode()
    with open(path, "wb") as f:
        hdr = bytearray(68)
        hdr[0:6] = b"BLOW5\x01"
        hdr[6:9] = bytes([0, 2, 0])
        hdr[10:14] = struct.pack("<I", 1)
        hdr[64:68] = struct.pack("<I", len(text))
        f.write(hdr)
        f.write(text)
        fixed = struct.pack("<I4d", 0, 8192.0, 9.0, 1444.86, 4000.0)
        pos = 0
        for r, n in enumerate(lens):
            rid = ("read-%08d-%06d" % (r, n)).encode()
            body = struct.pack("<H", len(rid)) + rid + fixed + struct.pack("<Q", int(n)) + sig[pos:pos + n].tobytes()
            f.write(struct.pack("<Q", len(body)))
            f.write(body)
            pos += n
        f.write(b"5WOLB")
    return lens


def _records(path):
    b = open(path, "rb").read()
    pos = 68 + struct.unpack_from("<I", b, 64)[0]
    head = b[:pos]
    out = []
    while b[pos:pos + 5] != b"5WOLB":
        sz = struct.unpack_from("<Q", b, pos)[0]
        out.append(b[pos + 8:pos + 8 + sz])
        pos += 8 + sz
    return head, out


@have_ref
def test_config3_recode_lognormal_lengths_sharded(cdc, tmp_path):
    raw = tmp_path / "raw.blow5"
    lens = _lognormal_file(str(raw), 600, seed=77)
    assert lens.min() >= 500 and lens.max() <= 200000 and lens.max() > 30 * lens.min()
    ref_z = tmp_path / "ref_zlib_svb.blow5"          # reference-written zlib+svb-zd input, as the config says
    subprocess.check_call([REF, "view", str(raw), "-o", str(ref_z), "-t", "16"], stderr=subprocess.DEVNULL)
    head, recs = _records(str(ref_z))
    # whole batch in one call
    rc, image = cdc.blow5_recode(METHOD.ZLIB, METHOD.SVB_ZD, METHOD.ZLIB, METHOD.SVB_ZD, recs)
    assert rc == 0
    # 8 contiguous byte-balanced shards (one per rank / GPU), processed independently, concatenated in rank order
    bounds = shard_bounds([len(r) for r in recs], 8)
    share = [sum(len(r) for r in recs[bounds[k]:bounds[k + 1]]) for k in range(8)]
    assert max(share) <= sum(share) / 8 + max(len(r) for r in recs)
    parts = []
    for k in range(8):
        rc, img = cdc.blow5_recode(METHOD.ZLIB, METHOD.SVB_ZD, METHOD.ZLIB, METHOD.SVB_ZD, recs[bounds[k]:bounds[k + 1]])
        assert rc == 0
        parts.append(img)
    assert b"".join(parts) == image                 # no cross-record dependency: sharding does not change a byte
    # the re-encoded file is read by the reference to the same text as its own file
    mine = tmp_path / "mine.blow5"
    with open(mine, "wb") as f:
        f.write(head)
        f.write(image)
        f.write(b"5WOLB")
    t1, t2 = tmp_path / "t1.slow5", tmp_path / "t2.slow5"
    subprocess.check_call([REF, "view", str(mine), "-o", str(t1)], stderr=subprocess.DEVNULL)
    subprocess.check_call([REF, "view", str(ref_z), "-o", str(t2)], stderr=subprocess.DEVNULL)
    assert open(t1, "rb").read() == open(t2, "rb").read()
    assert os.path.getsize(mine) <= 1.03 * os.path.getsize(ref_z)
    # decode side of the config: zlib+svb-zd -> none/none must equal the raw records byte for byte
    rc, plain = cdc.blow5_recode(METHOD.ZLIB, METHOD.SVB_ZD, METHOD.NONE, METHOD.NONE, recs)
    assert rc == 0
    _, raw_recs = _records(str(raw))
    assert plain == b"".join(struct.pack("<Q", len(r)) + r for r in raw_recs)
    # and the CLI does the same conversion
    out = tmp_path / "cli.blow5"
    subprocess.check_call([CLI, "view", str(ref_z), "-o", str(out), "-c", "none", "-s", "none"], stderr=subprocess.DEVNULL)
    assert _records(str(out))[1] == raw_recs


def _refdrv():
    from conftest import build_oracle
    L = C.CDLL(build_oracle())
    L.refdrv_record_pass.restype = C.c_int
    L.refdrv_record_pass.argtypes = [C.c_char_p, C.c_char_p] + [C.c_int] * 4 + [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
        C.c_void_p, C.POINTER(C.c_double)]
    return L


def _ref_pass(L, ref_so, methods, src, off, length, out, out_off=None):
    nb, sec = C.c_uint64(), C.c_double()
    rc = L.refdrv_record_pass(ref_so, b"/tmp", *methods, src.ctypes.data, off.ctypes.data, length.ctypes.data, len(length),
                              os.cpu_count() or 1, out.ctypes.data, out.size, C.byref(nb),
                              out_off.ctypes.data if out_off is not None else None, C.byref(sec))
    assert rc == 0, rc
    return nb.value


@pytest.mark.parametrize("R", [100000, 1000000])
def test_config2_every_record_bit_exact_at_full_size(cdc, R):
    """BASELINE configs[1]/[2] at their full sizes, EVERY read: (a) the svb-zd records the GPU packs equal the oracle's
    byte for byte (the whole image is compared, not a sample); (b) the zlib + svb-zd image the GPU writes is decoded by
    the compiled reference (slow5_decode + slow5_encode per record) back to exactly the input records; (c) the GPU decodes
    its own image to the input."""
    N = 4096
    rl = synth.record_bytes(N)
    sig = synth.nanopore_signal(R * N, seed=4321, device="cuda")
    d_raw = synth.blow5_records(sig, R, N, seed=4321).view(-1)
    del sig
    h_raw = d_raw.cpu().numpy()
    off = np.arange(R, dtype=np.uint64) * np.uint64(rl)
    ln = np.full(R, rl, np.uint32)
    L = _refdrv()
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libslow5_ref.so")
    ref_so = ref_so.encode() if os.path.exists(ref_so) else None
    d_res = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_off = torch.zeros(R + 1, dtype=torch.int64, device="cuda")
    # (a) none/none -> none/svb-zd: deterministic, whole image against the oracle's restatement
    d_img = torch.zeros(R * (rl * 3 // 4) + 4096, dtype=torch.uint8, device="cuda")
    cdc.blow5_recode_dev(METHOD.NONE, METHOD.NONE, METHOD.NONE, METHOD.SVB_ZD, d_raw, R * rl, off, ln, d_img, d_res, None)
    cdc.sync()
    nb, err = (int(x) for x in d_res.cpu().numpy())
    assert err == 0
    want = np.zeros(d_img.numel(), np.uint8)
    nb_o = _ref_pass(L, None, (0, 0, 0, 2), h_raw, off, ln, want)          # oracle port (blow5_oracle.c + svbzd_oracle.c)
    assert nb == nb_o
    got = d_img[:nb].cpu().numpy()
    assert np.array_equal(got, want[:nb])
    del got, want, d_img
    # (b) full encode on the GPU, decoded by the reference library, all R records
    d_z = torch.zeros(R * (rl // 2 + 512), dtype=torch.uint8, device="cuda")
    cdc.blow5_recode_dev(METHOD.NONE, METHOD.NONE, METHOD.ZLIB, METHOD.SVB_ZD, d_raw, R * rl, off, ln, d_z, d_res, d_off)
    cdc.sync()
    nbz, err = (int(x) for x in d_res.cpu().numpy())
    assert err == 0
    h_z = d_z[:nbz].cpu().numpy()
    zoff_all = d_off.cpu().numpy().view(np.uint64)
    zo = zoff_all[:-1] + np.uint64(8)
    zl = (zoff_all[1:] - zoff_all[:-1] - np.uint64(8)).astype(np.uint32)
    back = np.zeros(R * (rl + 8) + 64, np.uint8)
    nb_back = _ref_pass(L, ref_so, (1, 2, 0, 0), h_z, zo, zl, back)
    assert nb_back == R * (rl + 8)
    b = back[:nb_back].reshape(R, rl + 8)
    assert np.array_equal(b[:, 8:], h_raw.reshape(R, rl))
    assert (b[:, :8] == np.frombuffer(np.uint64(rl).tobytes(), np.uint8)).all()
    del back, b
    # (c) and by the GPU itself
    d_back = torch.zeros(R * (rl + 8) + 64, dtype=torch.uint8, device="cuda")
    cdc.blow5_recode_dev(METHOD.ZLIB, METHOD.SVB_ZD, METHOD.NONE, METHOD.NONE, d_z, nbz, zo, zl, d_back, d_res, None)
    cdc.sync()
    nbb, err = (int(x) for x in d_res.cpu().numpy())
    assert err == 0 and nbb == R * (rl + 8)
    assert torch.equal(d_back[:nbb].view(R, rl + 8)[:, 8:], d_raw.view(R, rl))
