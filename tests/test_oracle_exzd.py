"""Pins the ex-zd oracle (oracle/exzd_oracle.c, a plain-C restatement of slow5_press.c:1236-1848) against golden
streams produced by the compiled reference -- synthetic inputs and the signal streams of the reference's own ex-zd
golden file (tests/golden/make_exzd_golden.py) -- and, when oracle/_ref is present, against that library directly."""
import os

import numpy as np
import pytest

from conftest import ref_call

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "exzd_ref_vectors.npz"))
NAMES = sorted(k[4:] for k in GOLD.files if k.startswith("in__"))


def test_golden_has_the_reference_fixture_and_the_edge_cases():
    assert any(n.startswith("fixture_") for n in NAMES)
    assert {"single_exception", "qts_5", "qts_6", "nanopore_1", "nanopore_257", "wrap_deltas"} <= set(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_encode_matches_the_reference_bytes(oracle, name):
    x, want = GOLD["in__" + name], GOLD["exzd__" + name].tobytes()
    assert oracle.exzd_compress(x) == want


@pytest.mark.parametrize("name", NAMES)
def test_decode_restores_the_signal(oracle, name):
    x, stream = GOLD["in__" + name], GOLD["exzd__" + name].tobytes()
    rc, back = oracle.exzd_depress(stream)
    assert rc == 0 and np.array_equal(back, x)


def test_header_fields(oracle):
    s = oracle.exzd_compress(GOLD["in__qts_3"])
    assert s[0] == 0 and int.from_bytes(s[1:9], "little") == 3000 and s[9] == 3
    s = oracle.exzd_compress(GOLD["in__qts_6"])
    assert s[9] == 5  # QTS never shifts more than 5 bits (slow5_press.c:1753)
    s = oracle.exzd_compress(GOLD["in__single_exception"])
    assert int.from_bytes(s[12:16], "little") == 1 and len(s) == 16 + 8 + 1500 - 1 - 1


def test_malformed_streams(oracle):
    good = GOLD["exzd__one_spike"].tobytes()
    assert oracle.exzd_depress(good)[0] == 0
    assert oracle.exzd_depress(b"\x01" + good[1:])[0] == -13           # unsupported version (:1838-1842)
    assert oracle.exzd_depress(good[:10])[0] == -2                      # header cut short
    assert oracle.exzd_depress(good[:-1])[0] == -13                     # a value byte missing
    assert oracle.exzd_depress(good + b"\x00")[0] == -13                # a byte too many
    bad = bytearray(good)
    bad[16] ^= 0x01                                                     # stated svb length no longer matches (:1492-1500)
    assert oracle.exzd_depress(bytes(bad))[0] == -13
    bad = bytearray(good)
    bad[9] = 6                                                          # q > 5 (SLOW5_ASSERT :1803)
    assert oracle.exzd_depress(bytes(bad))[0] == -13


def test_against_compiled_reference_random(oracle, reflib):
    if reflib is None:
        pytest.skip("oracle/_ref not built")
    from slow5tools_b200 import synth
    rng = np.random.default_rng(5)
    for t in range(150):
        n = int(rng.integers(1, 3000))
        kind = t % 5
        if kind == 0:
            x = synth.nanopore_signal(n, seed=t).numpy()
        elif kind == 1:
            x = (500 + rng.integers(-140, 140, n)).astype(np.int16)       # exceptions around the 8-bit edge
        elif kind == 2:
            x = ((synth.nanopore_signal(n, seed=t).numpy().astype(np.int32) >> 2) << 2).astype(np.int16)
        elif kind == 3:
            x = (rng.integers(-3, 4, n).cumsum() * 8).astype(np.int16)
        else:
            x = rng.integers(-32768, 32768, min(n, 200)).astype(np.int16)  # must fit the reference's count + 1024 buffer
        want = ref_call(reflib.slow5_ptr_compress_solo, 4, x.tobytes())
        assert want is not None and oracle.exzd_compress(x) == want
        back = ref_call(reflib.slow5_ptr_depress_solo, 4, want)
        assert back == x.tobytes()
        rc, mine = oracle.exzd_depress(want)
        assert rc == 0 and np.array_equal(mine, x)
