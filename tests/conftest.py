import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a CUDA device skips the gpu tests instead of failing them one by one.  (With a
    device present nothing is skipped: a missing or broken library must fail loudly there.)"""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if not have:
        skip = pytest.mark.skip(reason="no CUDA device on this machine (the gpu tests run on the B200 box)")
        for it in gpu_items:
            it.add_marker(skip)


class Oracle:
    """ctypes view of oracle/liboracle.so -- the CPU restatement used ONLY as the checker."""

    def __init__(self, path):
        self.lib = L = C.CDLL(path)
        vp, sz, u32, u64 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64
        L.orc_svbzd_bound.restype = sz
        L.orc_svbzd_bound.argtypes = [u32]
        L.orc_svbzd_compress.restype = sz
        L.orc_svbzd_compress.argtypes = [vp, sz, vp]
        L.orc_svbzd_depress.restype = C.c_int
        L.orc_svbzd_depress.argtypes = [vp, sz, vp, sz, C.POINTER(u32)]
        L.orc_svbzd_size.restype = sz
        L.orc_svbzd_size.argtypes = [vp, u32]
        L.orc_svbzd_compress_batch.restype = None
        L.orc_svbzd_compress_batch.argtypes = [vp, vp, vp, u64, vp, vp, vp]
        L.orc_svbzd_depress_batch.restype = C.c_int
        L.orc_svbzd_depress_batch.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp]

        L.orc_exzd_bound.restype = sz
        L.orc_exzd_bound.argtypes = [u64]
        L.orc_exzd_compress.restype = sz
        L.orc_exzd_compress.argtypes = [vp, sz, vp]
        L.orc_exzd_depress.restype = C.c_int
        L.orc_exzd_depress.argtypes = [vp, sz, vp, sz, C.POINTER(u64)]

    def exzd_compress(self, x):
        x = np.ascontiguousarray(x, dtype=np.int16)
        out = np.empty(int(self.lib.orc_exzd_bound(x.size)) + 16, np.uint8)
        k = self.lib.orc_exzd_compress(x.ctypes.data, x.nbytes, out.ctypes.data)
        return out[:k].tobytes()

    def exzd_depress(self, b, cap=None):
        buf = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8)
        n = int.from_bytes(b[1:9], "little") if len(b) >= 9 else 0
        cap = min(n, 1 << 28) if cap is None else cap
        out = np.zeros(max(cap, 1), np.int16)
        nn = C.c_uint64()
        rc = self.lib.orc_exzd_depress(buf.ctypes.data, len(b), out.ctypes.data, cap, C.byref(nn))
        return rc, out[:nn.value].copy() if rc == 0 else None

    def compress(self, x):
        x = np.ascontiguousarray(x, dtype=np.int16)
        out = np.empty(int(self.lib.orc_svbzd_bound(x.size)) + 16, np.uint8)
        k = self.lib.orc_svbzd_compress(x.ctypes.data, x.nbytes, out.ctypes.data)
        return out[:k].tobytes()

    def depress(self, b, cap=None):
        buf = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8)
        n = int.from_bytes(b[:4], "little") if len(b) >= 4 else 0
        cap = n if cap is None else cap
        out = np.zeros(max(cap, 1), np.int16)
        nn = C.c_uint32()
        rc = self.lib.orc_svbzd_depress(buf.ctypes.data, len(b), out.ctypes.data, cap, C.byref(nn))
        return rc, out[:nn.value].copy() if rc == 0 else None

    def compress_batch(self, sig, sig_off, n_samples, out_off):
        """numpy in; returns (out slab, out_len)."""
        sig = np.ascontiguousarray(sig, dtype=np.int16)
        sig_off = np.ascontiguousarray(sig_off, dtype=np.uint64)
        n_samples = np.ascontiguousarray(n_samples, dtype=np.uint32)
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        out = np.zeros(int(out_off[-1]) + 16, np.uint8)
        out_len = np.zeros(len(n_samples), np.uint32)
        self.lib.orc_svbzd_compress_batch(sig.ctypes.data, sig_off.ctypes.data, n_samples.ctypes.data,
                                          len(n_samples), out.ctypes.data, out_off.ctypes.data, out_len.ctypes.data)
        return out, out_len


def build_oracle():
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith((".c", ".h"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    return path


@pytest.fixture(scope="session")
def oracle():
    return Oracle(build_oracle())


@pytest.fixture(scope="session")
def reflib():
    """The UNMODIFIED reference slow5lib compiled by oracle/Makefile (None when not built)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libslow5_ref.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    for f in (L.slow5_ptr_compress_solo, L.slow5_ptr_depress_solo):
        f.restype = C.c_void_p
        f.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    return L


def ref_call(fn, method, data):
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
    n = C.c_size_t()
    p = fn(method, buf.ctypes.data, len(data), C.byref(n))
    if not p:
        return None
    out = C.string_at(p, n.value)
    C.CDLL(None).free(C.c_void_p(p))
    return out
