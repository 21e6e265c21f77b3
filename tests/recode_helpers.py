"""Helpers shared by the record-path tests: synthetic uncompressed records, the oracle's / the reference's per-record
transcoder (oracle/blow5_oracle.c, oracle/ref_driver.c), image walking."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M_NONE, M_ZLIB, M_SVB_ZD, M_ZSTD, M_EX_ZD = 0, 1, 2, 3, 4
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libslow5_ref.so")


def make_records(lengths, seed=1, aux=b"", constant=()):
    """List of uncompressed records (bytes, no size prefix) with the given sample counts.  Signal: nanopore-like random
    walk; reads whose index is in `constant` get a constant signal (compresses > 100x: the inflate-slot overflow case)."""
    rng = np.random.default_rng(seed)
    recs, sigs = [], []
    for i, n in enumerate(lengths):
        if i in constant:
            s = np.full(n, 512, np.int16)
        else:
            lv = np.repeat(rng.normal(500, 70, n // 10 + 2), 10)[:n]
            s = np.clip(np.rint(lv + rng.normal(0, 9, n)), 0, 2047).astype(np.int16)
        rid = ("read_%06d_%s" % (i, "x" * (i % 17))).encode()
        head = np.uint16(len(rid)).tobytes() + rid + np.uint32(i % 3).tobytes() + \
            np.array([8192.0, 9.0 + i, 1444.86, 4000.0], "<f8").tobytes() + np.uint64(n).tobytes()
        recs.append(head + s.tobytes() + aux)
        sigs.append(s)
    return recs, sigs


def slab(records, align=1, gap=0):
    """records -> (uint8 slab, uint64 offsets, uint32 lengths); `gap` bytes of filler before every record"""
    ln = np.array([len(r) for r in records], np.uint32)
    off = np.zeros(len(records), np.uint64)
    at = 0
    for i, r in enumerate(records):
        at += gap
        at = (at + align - 1) // align * align
        off[i] = at
        at += len(r)
    buf = np.full(at + 64, 0xEE, np.uint8)
    for r, o in zip(records, off):
        buf[int(o):int(o) + len(r)] = np.frombuffer(r, np.uint8)
    return buf, off, ln, at


def walk_image(img):
    """file image bytes -> list of records (without their size prefixes)"""
    out, at = [], 0
    img = bytes(img)
    while at < len(img):
        sz = int.from_bytes(img[at:at + 8], "little")
        out.append(img[at + 8:at + 8 + sz])
        at += 8 + sz
    assert at == len(img)
    return out


class RecordOracle:
    def __init__(self, liboracle_path):
        self.L = L = C.CDLL(liboracle_path)
        L.orc_blow5_recode_record.restype = C.c_int
        L.orc_blow5_recode_record.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.refdrv_record_pass.restype = C.c_int
        L.refdrv_record_pass.argtypes = [C.c_char_p, C.c_char_p] + [C.c_int] * 4 + [
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
            C.c_void_p, C.POINTER(C.c_double)]
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]

    def recode(self, methods, rec):
        """one record through the oracle's restatement; returns (rc, bytes incl. the size prefix)"""
        buf = np.frombuffer(rec, np.uint8) if len(rec) else np.zeros(1, np.uint8)
        p, n = C.c_void_p(), C.c_size_t()
        rc = self.L.orc_blow5_recode_record(*methods, buf.ctypes.data, len(rec), C.byref(p), C.byref(n))
        if rc != 0:
            return rc, None
        out = C.string_at(p, n.value)
        self.libc.free(p)
        return 0, out

    def batch(self, methods, records, use_ref, threads=4):
        """whole batch through the pthread pool driver; use_ref: the compiled reference (slow5_decode + slow5_encode)"""
        buf, off, ln, used = slab(records)
        out = np.zeros(int(ln.sum()) * 4 + 64 * len(records) + 4096, np.uint8)
        oo = np.zeros(len(records) + 1, np.uint64)
        nb, sec = C.c_uint64(), C.c_double()
        rc = self.L.refdrv_record_pass(REF_SO.encode() if use_ref else None, b"/tmp", *methods, buf.ctypes.data,
                                       off.ctypes.data, ln.ctypes.data, len(records), threads, out.ctypes.data, out.size,
                                       C.byref(nb), oo.ctypes.data, C.byref(sec))
        return rc, out[:nb.value].tobytes()
