"""The raw_signal column of SLOW5 text records on the GPU (ascii_kernels.cu; slow5.c:3866-3878 and :2754-2778):
formatting against Python's own str(), parsing against the reference's acceptance rules (slow5_int_check + strtol +
int16 range, slow5_misc.c:122-139, :303-319)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import slow5tools_b200 as s5  # noqa: E402
from slow5tools_b200 import _capi  # noqa: E402
from slow5tools_b200._capi import METHOD  # noqa: E402

lib = s5.lib
lib.s5b_signal_to_ascii_batch_host.restype = C.c_int
lib.s5b_signal_to_ascii_batch_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t,
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
lib.s5b_ascii_to_signal_batch_host.restype = C.c_int
lib.s5b_ascii_to_signal_batch_host.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_uint64),
                                               C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]


@pytest.fixture(scope="module")
def cdc():
    c = s5.Codec(0)
    yield c
    c.close()


def to_ascii(cdc, method, bufs):
    n = len(bufs)
    keep = [np.frombuffer(b, np.uint8) if len(b) else np.zeros(1, np.uint8) for b in bufs]
    ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
    counts = (C.c_size_t * n)(*[len(b) for b in bufs])
    outp, outn = (C.c_void_p * n)(), (C.c_size_t * n)()
    rc = lib.s5b_signal_to_ascii_batch_host(cdc._h, method, ptrs, counts, n, outp, outn)
    res = []
    for i in range(n):
        res.append(C.string_at(outp[i], outn[i]) if outp[i] else None)
        if outp[i]:
            _capi.free(outp[i])
    return rc, res


def from_ascii(cdc, texts, expect):
    n = len(texts)
    keep = [np.frombuffer(t, np.uint8) if len(t) else np.zeros(1, np.uint8) for t in texts]
    ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
    counts = (C.c_size_t * n)(*[len(t) for t in texts])
    ex = (C.c_uint64 * n)(*expect)
    outp, outn = (C.c_void_p * n)(), (C.c_size_t * n)()
    rc = lib.s5b_ascii_to_signal_batch_host(cdc._h, ptrs, counts, ex, n, outp, outn)
    res = []
    for i in range(n):
        if outp[i]:
            res.append(np.frombuffer(C.string_at(outp[i], outn[i] * 2), np.int16).copy())
            _capi.free(outp[i])
        else:
            res.append(None)
    return rc, res


def arrays():
    rng = np.random.default_rng(5)
    out = [np.array([], np.int16), np.array([0], np.int16), np.array([-32768, 32767, -1, 0, 9, 10, 99, 100, 999, 1000, 9999, 10000], np.int16)]
    for n in (1, 2, 7, 8, 9, 255, 256, 257, 4096, 30001):
        out.append(np.clip(np.rint(rng.normal(500, 90, n)), 0, 2047).astype(np.int16))
    out.append(rng.integers(-32768, 32768, 5000).astype(np.int16))
    out.append(np.zeros(1000, np.int16))
    return out


def text_of(a):
    return ",".join(str(int(v)) for v in a).encode()


def test_format_matches_sprintf(cdc):
    arrs = arrays()
    rc, got = to_ascii(cdc, METHOD.NONE, [a.tobytes() for a in arrs])
    assert rc == 0
    for a, t in zip(arrs, got):
        assert t == text_of(a)


def test_format_from_stored_streams(cdc):
    arrs = [a for a in arrays() if a.size >= 2]
    rc, svb = cdc.compress_batch(METHOD.SVB_ZD, [a.tobytes() for a in arrs])
    assert rc == 0
    rc, got = to_ascii(cdc, METHOD.SVB_ZD, svb)
    assert rc == 0 and got == [text_of(a) for a in arrs]
    # ex-zd: signal-like arrays only (the reference's encoder aborts on streams that outgrow its buffer, and so does ours)
    sane = [a for a in arrs if int(np.abs(np.diff(a.astype(np.int32))).max(initial=0)) < 1000 and a.min() >= 0]
    rc, ex = cdc.compress_batch(METHOD.EX_ZD, [a.tobytes() for a in sane])
    assert rc == 0
    rc, got = to_ascii(cdc, METHOD.EX_ZD, ex)
    assert rc == 0 and got == [text_of(a) for a in sane]
    # a corrupt stream is an error, not text
    bad = bytearray(svb[5])
    bad[0:4] = (len(bad) * 2).to_bytes(4, "little")
    rc, _ = to_ascii(cdc, METHOD.SVB_ZD, [bytes(bad)])
    assert rc != 0


def test_parse_round_trip(cdc):
    arrs = arrays()
    rc, got = from_ascii(cdc, [text_of(a) for a in arrs], [a.size for a in arrs])
    assert rc == 0
    for a, g in zip(arrs, got):
        assert g is not None and np.array_equal(a, g)


@pytest.mark.parametrize("text,expect,ok,value", [
    (b"5", 1, True, [5]),
    (b"-", 1, True, [0]),                 # int_check lets '-' through, strtol finds no digits: 0
    (b"1-2", 1, True, [1]),               # strtol stops at the second '-'
    (b"-0", 1, True, [0]),
    (b"0", 1, True, [0]),
    (b"-32768,32767", 2, True, [-32768, 32767]),
    (b"007", 1, False, None),             # leading zero
    (b"", 1, False, None),                # empty token
    (b"1,,2", 3, False, None),
    (b"1,2,", 3, False, None),            # trailing comma: empty last token
    (b"32768", 1, False, None),
    (b"-32769", 1, False, None),
    (b"12a", 1, False, None),
    (b"1 2", 1, False, None),
    (b"+5", 1, False, None),
    (b"1,2,3", 2, False, None),           # more samples than len_raw_signal says
    (b"1,2,3", 4, False, None),           # fewer
    (b"1234567", 1, False, None),
])
def test_parse_acceptance_rules(cdc, text, expect, ok, value):
    rc, got = from_ascii(cdc, [b"1,2,3", text, b"4"], [3, expect, 1])   # good neighbours must not be affected
    assert np.array_equal(got[0], [1, 2, 3]) and np.array_equal(got[2], [4])
    if ok:
        assert rc == 0 and np.array_equal(got[1], value)
    else:
        assert rc == -2 and got[1] is None
