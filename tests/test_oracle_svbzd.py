"""Pins the oracle (oracle/svbzd_oracle.c) against (1) the reference's own unit-test vectors and the
SURVEY 8c known answers, (2) golden vectors produced by the compiled reference, (3) the compiled
reference itself on random inputs when oracle/_ref is present.  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import ref_call

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "svbzd_kat.json")))


@pytest.mark.parametrize("case", KAT["encode"], ids=lambda c: c["name"])
def test_kat_encode_and_roundtrip(oracle, case):
    x = np.array(case["in"], dtype=np.int16)
    enc = oracle.compress(x)
    assert enc.hex() == case["hex"]
    assert int.from_bytes(enc[:4], "little") == x.size          # unit_test_press.c:131,161
    rc, dec = oracle.depress(enc)
    assert rc == 0 and np.array_equal(dec, x)                    # unit_test_press.c:140,170


@pytest.mark.parametrize("case", KAT["decode_only"], ids=lambda c: c["name"])
def test_kat_decode_only(oracle, case):
    rc, dec = oracle.depress(bytes.fromhex(case["hex"]))
    assert rc == 0 and dec.tolist() == case["out"]


def test_realistic_vector_is_smaller_than_raw(oracle):
    x = np.array([1039, 588, 588, 593, 586, 574, 570, 585, 588, 586], dtype=np.int16)
    assert len(oracle.compress(x)) < x.nbytes                    # unit_test_press.c:184


def test_golden_vectors_from_reference(oracle):
    g = np.load(os.path.join(HERE, "golden", "svbzd_ref_vectors.npz"))
    names = [k[4:] for k in g.files if k.startswith("in__")]
    assert len(names) >= 30
    for name in names:
        x, want = g["in__" + name], g["svb__" + name].tobytes()
        assert oracle.compress(x) == want, name
        rc, dec = oracle.depress(want)
        assert rc == 0 and np.array_equal(dec, x), name


def test_malformed_streams(oracle):
    good = oracle.compress(np.arange(10, dtype=np.int16))
    assert oracle.depress(good[:-1])[0] == -13         # short data (slow5_press.c:1130-1136)
    assert oracle.depress(good + b"\0")[0] == -13      # trailing byte
    assert oracle.depress(b"\x01\x00")[0] == -2
    assert oracle.depress(b"\x05\x00\x00\x00")[0] == -13


def test_against_compiled_reference_random(oracle, reflib):
    if reflib is None:
        pytest.skip("oracle/_ref/libslow5_ref.so not built (no /root/reference here)")
    rng = np.random.default_rng(5)
    for i in range(400):
        n = int(rng.integers(0, 3000))
        mode = i % 4
        if mode == 0:
            x = rng.integers(-32768, 32768, n)
        elif mode == 1:
            x = 500 + np.cumsum(rng.integers(-9, 10, n))
        elif mode == 2:
            x = np.where(rng.random(n) < 0.02, rng.integers(-32768, 32768, n), 400 + rng.integers(-100, 100, n))
        else:
            x = rng.integers(0, 2048, n)
        x = x.astype(np.int16)
        want = ref_call(reflib.slow5_ptr_compress_solo, 2, x.tobytes())
        assert want is not None
        assert oracle.compress(x) == want
        back = ref_call(reflib.slow5_ptr_depress_solo, 2, want)
        rc, dec = oracle.depress(want)
        assert rc == 0
        if n:
            assert back == x.tobytes() and np.array_equal(dec, x)
