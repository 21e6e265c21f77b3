"""Generates tests/golden/svbzd_ref_vectors.npz from the UNMODIFIED reference library
(oracle/_ref/libslow5_ref.so, built by `make -C oracle ref` from /root/reference).
Run in the build container only; the .npz is committed so the GPU box needs no reference tree.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libslow5_ref.so"))
    ref.slow5_ptr_compress_solo.restype = C.c_void_p
    ref.slow5_ptr_compress_solo.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(20261017)
    from slow5tools_b200 import synth
    cases = []
    # lengths around every structural boundary: key byte (4), lane (8), warp iteration (256), chunk (1024)
    lens = [0, 1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 4097, 5336]
    sig = synth.nanopore_signal(sum(lens) + 8192, seed=7).numpy()
    pos = 0
    for n in lens:
        cases.append(("nanopore_%d" % n, sig[pos:pos + n].copy()))
        pos += n
    for kind in ("uniform", "alternating", "constant", "boundary"):
        for n in (5, 300, 2050):
            cases.append(("%s_%d" % (kind, n), synth.adversarial(kind, n, seed=3)))
    cases.append(("mixed_rare_big_jumps", np.where(rng.random(3000) < 0.01, rng.integers(-32768, 32767, 3000),
                                                    500 + rng.integers(-20, 20, 3000)).astype(np.int16)))
    out = {}
    for name, x in cases:
        x = np.ascontiguousarray(x, dtype=np.int16)
        buf = x if x.size else np.zeros(1, np.int16)
        n = C.c_size_t()
        p = ref.slow5_ptr_compress_solo(2, buf.ctypes.data, x.nbytes, C.byref(n))
        assert p, name
        out["in__" + name] = x
        out["svb__" + name] = np.frombuffer(C.string_at(p, n.value), dtype=np.uint8).copy()
        libc.free(p)
    path = os.path.join(ROOT, "tests", "golden", "svbzd_ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(cases), "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
