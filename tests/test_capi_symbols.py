"""The C-ABI library loads, exports every symbol include/slow5b200.h declares, and refuses to compute
without a GPU (no CPU fallback).  CPU only -- no compute calls."""
import ctypes as C
import os
import re

import numpy as np
import torch

import slow5tools_b200 as s5
from slow5tools_b200 import codec
from slow5tools_b200._capi import METHOD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "slow5b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s5b_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    syms = declared_symbols()
    assert len(syms) >= 18
    for name in syms:
        assert hasattr(s5.lib, name), name


def test_bound_matches_formula():
    for n in (0, 1, 3, 4, 5, 4096, 200000):
        assert s5.lib.s5b_svbzd_bound(n) == 4 + (n + 3) // 4 + 3 * n
        assert s5.lib.s5b_svbzd_slot(n) % 16 == 0 and s5.lib.s5b_svbzd_slot(n) >= s5.lib.s5b_svbzd_bound(n)


def test_layout_helpers():
    n = np.array([0, 1, 8, 9, 4096], np.uint32)
    off = s5.sig_layout(n)
    assert off.tolist() == [0, 0, 8, 16, 32, 32 + 4096]
    so = s5.svb_slot_layout(n)
    assert all(int(so[i + 1] - so[i]) == s5.lib.s5b_svbzd_slot(int(n[i])) for i in range(len(n)))


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        return  # exercised by the gpu tests instead
    h = C.c_void_p()
    assert s5.lib.s5b_device_count() == 0
    assert s5.lib.s5b_ctx_create(0, C.byref(h)) == s5.ERR.DEVICE
    x = np.arange(100, dtype=np.int16).tobytes()
    assert codec.ptr_compress_solo(METHOD.SVB_ZD, x) is None
    assert s5.lib.s5b_last_error() == s5.ERR.DEVICE
    assert codec.ptr_depress_solo(METHOD.SVB_ZD, b"\0\0\0\0") is None
    assert s5.lib.s5b_last_error() == s5.ERR.DEVICE


def test_strerror():
    assert s5.strerror(0) == "ok"
    assert "stream" in s5.strerror(s5.ERR.PRESS)
