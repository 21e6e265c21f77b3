"""Generates tests/golden/exzd_ref_vectors.npz from the UNMODIFIED reference library (oracle/_ref/libslow5_ref.so):
ex-zd streams produced by slow5_ptr_compress_solo(SLOW5_COMPRESS_EX_ZD = 4) for synthetic inputs, plus the ex-zd
signal streams cut out of the reference's own golden file test/data/exp/one_fast5/exp_1_lossless_zlib_ex_zd.blow5
(the file test/test_view.sh:102-163 compares against) together with the raw signals of the matching uncompressed
golden.  Run in the build container only; the .npz is committed so the GPU box needs no reference tree.

    python tests/golden/make_exzd_golden.py
"""
import ctypes as C
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def blow5_records(path):
    """(record_method, signal_method, [raw record bytes ...]) of a BLOW5 file (slow5.c:794-881, :3233-3281)."""
    b = open(path, "rb").read()
    assert b[:6] == b"BLOW5\x01"
    rec_m = b[9]
    sig_m = b[14]
    hdr = struct.unpack_from("<I", b, 64)[0]
    pos = 68 + hdr
    out = []
    while b[pos:pos + 5] != b"5WOLB":
        sz = struct.unpack_from("<Q", b, pos)[0]
        out.append(b[pos + 8:pos + 8 + sz])
        pos += 8 + sz
    return rec_m, sig_m, out


def signal_of(rec, raw):
    """stored signal bytes of a packed binary record (slow5.c:3928-3987); the length field counts samples for a raw
    signal and bytes for a compressed one (:3983-3987)"""
    idl = struct.unpack_from("<H", rec, 0)[0]
    p = 2 + idl + 4 + 8 * 4
    n = struct.unpack_from("<Q", rec, p)[0] * (2 if raw else 1)
    return rec[:2 + idl], rec[p + 8:p + 8 + n]


def main():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libslow5_ref.so"))
    ref.slow5_ptr_compress_solo.restype = C.c_void_p
    ref.slow5_ptr_compress_solo.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(20261017)
    from slow5tools_b200 import synth
    cases = []
    lens = [1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 255, 256, 257, 258, 1023, 1024, 1025, 4095, 4096, 4097, 5336, 70000]
    sig = synth.nanopore_signal(sum(lens) + 8192, seed=9).numpy()
    pos = 0
    for n in lens:
        cases.append(("nanopore_%d" % n, sig[pos:pos + n].copy()))
        pos += n
    for q in range(1, 7):  # QTS: all samples share q low zero bits (what `slow5tools degrade` produces)
        x = synth.nanopore_signal(3000, seed=20 + q).numpy().astype(np.int32)
        cases.append(("qts_%d" % q, ((x >> q) << q).astype(np.int16)))
    cases.append(("constant_2000", np.full(2000, 437, np.int16)))
    cases.append(("zeros_300", np.zeros(300, np.int16)))
    cases.append(("negative_levels", (synth.nanopore_signal(2500, seed=31).numpy().astype(np.int32) - 900).astype(np.int16)))
    x = synth.nanopore_signal(2600, seed=32).numpy().copy()
    x[1300] = 30000
    cases.append(("one_spike", x))  # two exceptions (up and down)
    x = np.full(1500, 501, np.int16)
    x[700:] = 901
    cases.append(("single_exception", x))  # exactly one exception -> raw (pos, value) pair
    cases.append(("noisy_many_exceptions", (500 + rng.integers(-300, 300, 3000)).astype(np.int16)))  # ~57 % exceptions
    cases.append(("wrap_deltas", np.where(np.arange(400) % 2 == 0, -32768, 32767).astype(np.int16)[:256]))
    cases.append(("uniform_200", rng.integers(-32768, 32768, 200).astype(np.int16)))
    out = {}
    for name, x in cases:
        x = np.ascontiguousarray(x, dtype=np.int16)
        n = C.c_size_t()
        p = ref.slow5_ptr_compress_solo(4, x.ctypes.data, x.nbytes, C.byref(n))
        assert p, name
        out["in__" + name] = x
        out["exzd__" + name] = np.frombuffer(C.string_at(p, n.value), dtype=np.uint8).copy()
        libc.free(p)
    # the reference's own ex-zd golden file and the raw signals of its uncompressed twin
    rm, sm, recs = blow5_records(os.path.join(REF, "test/data/exp/one_fast5/exp_1_lossless_zlib_ex_zd.blow5"))
    assert (rm, sm) == (1, 2), (rm, sm)  # zlib records, ex-zd signal (slow5_press.c:58-161)
    rm0, sm0, raw = blow5_records(os.path.join(REF, "test/data/exp/one_fast5/exp_1_lossless.blow5"))
    assert (rm0, sm0) == (0, 0) or rm0 == 0
    for i, (z, r0) in enumerate(zip(recs, raw)):
        idz, sz = signal_of(zlib.decompress(z), False)
        id0, s0 = signal_of(r0, True)
        assert idz == id0
        out["in__fixture_%d" % i] = np.frombuffer(s0, dtype=np.int16).copy()
        out["exzd__fixture_%d" % i] = np.frombuffer(sz, dtype=np.uint8).copy()
    path = os.path.join(ROOT, "tests", "golden", "exzd_ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out) // 2, "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
