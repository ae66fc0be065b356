"""BFV (HPSPOVERQ) multiplication tables, re-derived from (Q, R, t) alone and checked against the reference's own
CryptoParametersBFVRNS tables in the golden fixtures (bfvrns-cryptoparameters.cpp:140-600): first with exact big-integer
arithmetic straight from the reference's formulas, then through the library's word-size host routine
(ofhe_bfv_tables_host), which is what lets a caller build the BFV multiplication from moduli without the reference."""
import numpy as np
import pytest

import ofhe_b200

NAMES = ["QlHatInvModq", "QlHatModr", "alphaQlModr", "qInv", "mNegRlQlHatInvModq", "qInvModr", "RlHatInvModr", "RlHatModq",
         "alphaRlModq", "rInv", "tQlSlHatInvModsDivsModq", "tQlSlHatInvModsDivsFrac"]


def derive(Q, R, t):
    """Exact restatement with Python integers (level l = sizeQ, sizeR = number of R moduli)."""
    Q, R = [int(x) for x in Q], [int(x) for x in R]
    nQ, nR = len(Q), len(R)
    pq, pr = 1, 1
    for x in Q:
        pq *= x
    for x in R:
        pr *= x
    inv = lambda a, m: pow(a % m, -1, m)
    T = {}
    T["QlHatInvModq"] = np.array([inv(pq // q, q) for q in Q], dtype=np.uint64)
    T["QlHatModr"] = np.array([[(pq // q) % r for q in Q] for r in R], dtype=np.uint64)
    T["alphaQlModr"] = np.array([[(pq % r) * k % r for r in R] for k in range(nQ + 1)], dtype=np.uint64)
    T["qInv"] = np.array([1.0 / float(q) for q in Q])
    T["mNegRlQlHatInvModq"] = np.array([q - (pr * inv(pq // q, q)) % q for q in Q], dtype=np.uint64)
    T["qInvModr"] = np.array([[inv(q, r) for r in R] for q in Q], dtype=np.uint64)
    T["RlHatInvModr"] = np.array([inv(pr // r, r) for r in R], dtype=np.uint64)
    T["RlHatModq"] = np.array([[(pr // r) % q for r in R] for q in Q], dtype=np.uint64)
    T["alphaRlModq"] = np.array([[(pr % q) * k % q for q in Q] for k in range(nR + 1)], dtype=np.uint64)
    T["rInv"] = np.array([1.0 / float(r) for r in R])
    qr = pq * pr
    tab = np.zeros((nQ, nR + 1), dtype=np.uint64)
    frac = np.zeros(nR)
    for j, r in enumerate(R):
        x = t * pq * inv(qr // r, r)
        frac[j] = float(x % r) / float(r)
        for i, q in enumerate(Q):
            tab[i, j] = (x // r) % q
    for i, q in enumerate(Q):
        tab[i, nR] = ((t * pq * inv(qr // q, q)) // q) % q
    T["tQlSlHatInvModsDivsModq"], T["tQlSlHatInvModsDivsFrac"] = tab, frac
    return T


@pytest.mark.parametrize("name", ["bfv_n64", "bfv_n128_d4", "bfv_cfg3_tables"])
def test_tables_from_moduli(golden, name):
    g = golden(name)
    N, nQ, nR = (int(v) for v in g["shape"])
    t = int(g["args"][3])
    Q, R = g["moduli"][:nQ], g["moduli"][nQ:]
    big = derive(Q, R, t)
    lib = ofhe_b200.bfv_tables(Q, R, t)
    for k in NAMES:
        want = g["T_" + k]
        assert big[k].shape == want.shape, k
        assert (big[k] == want).all(), f"big-integer restatement of {k}"
        assert lib[k].shape == want.shape and (lib[k] == want).all(), f"ofhe_bfv_tables_host {k}"
