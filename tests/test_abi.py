"""CPU tier: the C-ABI library loads, exports every symbol include/ofhe_b200.h declares, and the host-side helpers
(setup-time number theory, CKKS chain) agree with the oracle / the reference.  No compute call needs a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

import ofhe_b200


def _declared_symbols():
    txt = open(ofhe_b200.HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ofhe_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(ofhe_b200.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(ofhe_b200.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    # the Python binding mirrors the header one to one
    assert sorted(ofhe_b200.SIGNATURES) == declared


def test_version_and_error_string():
    L = ofhe_b200.lib()
    assert b"sm_100a" in L.ofhe_version()
    assert isinstance(L.ofhe_last_error(), bytes)


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(ofhe_b200.OfheError, match="no CUDA device|CPU fallback"):
        ofhe_b200.Context(0)


def test_host_number_theory_matches_oracle(oracle):
    L, o = ofhe_b200.lib(), oracle.L
    for bits, m in [(59, 1 << 17), (60, 1 << 13), (50, 1 << 13), (28, 16), (16, 16), (60, 1 << 18)]:
        assert L.ofhe_first_prime(bits, m) == o.orc_first_prime(bits, m)
        q = L.ofhe_last_prime(bits, m)
        assert q == o.orc_last_prime(bits, m)
        assert L.ofhe_previous_prime(q, m) == o.orc_previous_prime(q, m)
        assert L.ofhe_next_prime(q, m) == o.orc_next_prime(q, m)
        if m <= (1 << 14):
            assert L.ofhe_root_of_unity(m, q) == o.orc_root_of_unity(m, q)
    assert L.ofhe_root_of_unity(8, 113) == 18


@pytest.mark.parametrize("args", [(64, 4, 40, 50, 2, 0), (128, 5, 45, 55, 3, 2), (1 << 10, 6, 59, 60, 3, 1),
                                  (1 << 9, 1, 45, 50, 0, 0), (1 << 12, 12, 50, 60, 0, 2), (1 << 11, 7, 50, 60, 0, 3),
                                  (256, 3, 40, 50, 2, 3)])
def test_ckks_paramgen_matches_oracle(oracle, args):
    p = ofhe_b200.ckks_paramgen(*args)
    c = oracle.ckks(*args)
    try:
        assert (p["sizeQ"], p["sizeP"], p["dnum"], p["alpha"]) == (c.sizeQ, c.sizeP, c.dnum, c.alpha)
        assert (p["moduli"] == c.moduli).all()
        assert (p["roots"] == c.roots).all()
    finally:
        c.close()


def test_ckks_paramgen_matches_golden(golden):
    for name in ("ckks_n64", "ckks_n128_flex"):
        g = golden(name)
        p = ofhe_b200.ckks_paramgen(*(int(v) for v in g["args"]))
        assert (p["moduli"] == g["moduli"]).all() and (p["roots"] == g["roots"]).all()
        assert [p["N"], p["sizeQ"], p["sizeP"], p["dnum"], p["alpha"]] == g["shape"].tolist()


def test_ckks_paramgen_config2_shape():
    # BASELINE.json configs[1]: N = 2^16, depth 44, scaleModSize 59, firstMod 60, HYBRID -> 45 + 15 limbs, dnum 3
    # (the shape the reference itself reports, SURVEY.md header)
    L = ofhe_b200.lib()
    m = 1 << 17
    q = L.ofhe_first_prime(59, m)
    assert q % m == 1
    # full chain generation at N = 2^16 costs ~60 minimal-root searches of 65536 steps: fine on the CPU tier
    p = ofhe_b200.ckks_paramgen(1 << 16, 44, 59, 60, 0, 0)
    assert (p["sizeQ"], p["sizeP"], p["dnum"], p["alpha"]) == (45, 15, 3, 15)
    assert len(set(p["moduli"].tolist())) == 60


def test_find_automorphism_index():
    # FindAutomorphismIndex2nComplex (core/lib/math/nbtheory2.cpp:243-262)
    m = 1 << 7
    assert ofhe_b200.find_automorphism_index(0, m) == 1
    assert ofhe_b200.find_automorphism_index(1, m) == 5
    assert ofhe_b200.find_automorphism_index(2, m) == 25
    inv5 = pow(5, -1, m)
    assert ofhe_b200.find_automorphism_index(-1, m) == inv5
    assert ofhe_b200.find_automorphism_index(-3, m) == pow(inv5, 3, m)


def test_paramgen_errors_are_reported_not_thrown():
    with pytest.raises(ofhe_b200.OfheError):
        ofhe_b200.ckks_paramgen(64, 4, 40, 50, 2, 4)  # COMPOSITESCALING* are not restated
    with pytest.raises(ofhe_b200.OfheError):
        ofhe_b200.ckks_paramgen(64, 4, 40, 50, 4, 0)  # 5 towers cannot fill 4 digits of ceil(5/4) = 2 (rns-cryptoparameters.cpp:87-93)
