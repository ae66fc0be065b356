"""The oracle (oracle/ofhe_oracle.c) against the reference's own known-answer tests and the committed golden
vectors generated from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle_lib import seeded_residues


def test_kat_crt_polynomial_mult_q113(oracle):
    # UTTransform.CRT_polynomial_mult, src/core/unittest/UnitTestTransform.cpp:52-90:
    # q = 113, m = 8, a = b = (1,2,4,1): INTT(NTT(a) .* NTT(b)) == (94,109,11,18)
    q, m = 113, 8
    root = oracle.L.orc_root_of_unity(m, q)
    a = np.array([1, 2, 4, 1], dtype=np.uint64)
    A = oracle.ntt_fwd(q, root, a)
    C = (A * A) % np.uint64(q)
    c = oracle.ntt_inv(q, root, C)
    assert c.tolist() == [94, 109, 11, 18]


@pytest.mark.parametrize("m,vals", [(16, [431, 3414, 1234, 7845, 2145, 7415, 5471, 8452]), (8, [4127, 9647, 1987, 5410]),
                                    (4, [9647, 1987])])
def test_switch_format_round_trip_single_crt(oracle, m, vals):
    # UTNTT.switch_format_simple_single_crt, src/core/unittest/UnitTestNTT.cpp:53-90 (16-bit modulus)
    q = oracle.L.orc_last_prime(16, m)
    root = oracle.L.orc_root_of_unity(m, q)
    x = np.array(vals, dtype=np.uint64)
    assert (oracle.ntt_inv(q, root, oracle.ntt_fwd(q, root, x)) == x).all()


def test_switch_format_round_trip_double_crt(oracle):
    # UTNTT.switch_format_simple_double_crt, UnitTestNTT.cpp:92-119: m = 16, two 28-bit towers
    m = 16
    q0 = oracle.L.orc_last_prime(28, m)
    q = [q0, oracle.L.orc_previous_prime(q0, m)]
    for vals in ([431, 3414, 1234, 7845, 2145, 7415, 5471, 8452], [4127, 9647, 1987, 5410, 6541, 7014, 9741, 1256]):
        x = np.array(vals, dtype=np.uint64)
        for qi in q:
            root = oracle.L.orc_root_of_unity(m, qi)
            assert (oracle.ntt_inv(qi, root, oracle.ntt_fwd(qi, root, x)) == x).all()


@pytest.mark.parametrize("tag", ["n64", "n4096", "n8"])
def test_ntt_golden(oracle, golden, tag):
    g = golden("core_vectors")
    q, roots = g[f"{tag}_moduli"], g[f"{tag}_roots"]
    N = g[f"{tag}_coeff"].shape[1]
    for l in range(len(q)):
        assert oracle.L.orc_root_of_unity(2 * N, int(q[l])) == int(roots[l])
    assert (oracle.dcrt_switch_format(q, roots, g[f"{tag}_coeff"], False) == g[f"{tag}_eval"]).all()
    assert (oracle.dcrt_switch_format(q, roots, g[f"{tag}_eval"], True) == g[f"{tag}_coeff"]).all()


def test_twiddles_golden(oracle, golden):
    g = golden("core_vectors")
    T, Tp, Ti, Tip, ni, nip = oracle.twiddles(int(g["n64_moduli"][0]), int(g["n64_roots"][0]), 64)
    assert (T == g["n64_T"]).all() and (Tp == g["n64_Tp"]).all()
    assert (Ti == g["n64_Ti"]).all() and (Tip == g["n64_Tip"]).all()
    assert [ni, nip] == g["n64_nInv"].tolist()


def test_limb_ops_golden(oracle, golden):
    g = golden("core_vectors")
    q = g["n64_moduli"]
    a, b = g["n64_a"], g["n64_b"]
    for op, name in ((0, "add"), (1, "sub"), (2, "mul"), (3, "neg")):
        assert (oracle.dcrt_binop(q, op, a, b) == g[f"n64_{name}"]).all(), name
    assert (oracle.dcrt_automorphism(5, a) == g["n64_auto5"]).all()
    assert (oracle.dcrt_automorphism(127, a) == g["n64_auto127"]).all()
    assert (oracle.switch_modulus(q[0], q[1], a[0]) == g["n64_swmod_down"]).all()
    assert (oracle.switch_modulus(q[1], q[0], b[1]) == g["n64_swmod_up"]).all()


@pytest.mark.parametrize("name", ["ckks_n64", "ckks_n128_flex"])
def test_ckks_golden(oracle, golden, name):
    g = golden(name)
    N, depth, sb, fb, dnum, st = (int(v) for v in g["args"])
    c = oracle.ckks(N, depth, sb, fb, dnum, st)
    try:
        assert [c.N, c.sizeQ, c.sizeP, c.dnum, c.alpha] == g["shape"].tolist()
        assert (c.moduli == g["moduli"]).all() and (c.roots == g["roots"]).all()
        for t in ("PInvModq", "PModq", "PHatInvModp", "PHatModq"):
            assert (getattr(c, t)() == g[t]).all(), t
        ea, eb = g["evk_a"], g["evk_b"]
        for sizeQl in (c.sizeQ, c.sizeQ - 1, 1):
            p = g[f"l{sizeQl}_c"]
            assert (c.modup(sizeQl, p) == g[f"l{sizeQl}_digits"]).all()
            k0, k1 = c.keyswitch_core(sizeQl, p, ea, eb)
            assert (np.stack([k0, k1]) == g[f"l{sizeQl}_ks"]).all()
            assert (c.evalmult(sizeQl, g[f"l{sizeQl}_ct1"], g[f"l{sizeQl}_ct2"], ea, eb) ==
                    g[f"l{sizeQl}_evalmult"]).all()
            if sizeQl > 1:
                ct1 = g[f"l{sizeQl}_ct1"]
                r = np.stack([c.rescale_poly(sizeQl, ct1[0]), c.rescale_poly(sizeQl, ct1[1])])
                assert (r == g[f"l{sizeQl}_rescale"]).all()
                A, B = c.rescale_tables(sizeQl)
                assert (A == g[f"l{sizeQl}_rescaleA"]).all() and (B == g[f"l{sizeQl}_rescaleB"]).all()
            for part in range(c.num_digits(sizeQl)):
                ds = c.digit_size(sizeQl, part)
                assert (c.PartQlHatInvModq(part, ds) == g[f"l{sizeQl}_p{part}_hatinv"]).all()
                m, cm = c.PartQlHatModp(sizeQl, part)
                assert (m == g[f"l{sizeQl}_p{part}_hatmodp"]).all() and (cm == g[f"l{sizeQl}_p{part}_compl"]).all()
        sizeQl = c.sizeQ - 1
        out = c.automorphism_ct(sizeQl, int(g["rot_k"][0]), g[f"l{sizeQl}_ct1"], g["rot_a"], g["rot_b"])
        assert (out == g["rot_out"]).all()
        ct, k, ra, rb = g[f"l{sizeQl}_ct1"], int(g["rot_k"][0]), g["rot_a"], g["rot_b"]
        digits = c.modup(sizeQl, ct[1])
        assert (c.fast_rotation(sizeQl, k, ct, digits, ra, rb) == g["rot_fast"]).all()
        assert (c.fast_rotation_ext(sizeQl, k, True, ct, digits, ra, rb) == g["rot_ext1"]).all()
        assert (c.fast_rotation_ext(sizeQl, k, False, ct, digits, ra, rb) == g["rot_ext0"]).all()
        assert (c.keyswitch_ext(sizeQl, True, ct) == g["ks_ext1"]).all()
        assert (c.keyswitch_ext(sizeQl, False, ct) == g["ks_ext0"]).all()
        assert (c.keyswitch_down(sizeQl, g["rot_ext1"]) == g["ks_down"]).all()
        # honest ciphertexts: the oracle's EvalMult reproduces the reference's product ciphertext, which the
        # reference decrypts to the slot-wise product (within CKKS noise)
        assert (c.evalmult(c.sizeQ, g["enc1"], g["enc2"], ea, eb) == g["enc_prod"]).all()
        assert np.allclose(g["dec_prod"], np.array([0.5, -0.25, 1.0, 0.125]) * np.array([2.0, 4.0, -1.0, 8.0]),
                           atol=1e-3)
    finally:
        c.close()


def test_scalar_arithmetic_edges(oracle):
    L = oracle.L
    rng = np.random.default_rng(5)
    for q in (113, (1 << 28) - 95 - 2 * 16 * 0, 1152921504606830593, 576460752303415297):
        if not L.orc_is_prime(q):
            continue
        mu = L.orc_barrett_mu(q)
        for a, b in [(0, 0), (q - 1, q - 1), (1, q - 1), (q // 2, q // 2 + 1)] + [
                (int(rng.integers(0, q)), int(rng.integers(0, q))) for _ in range(200)]:
            assert L.orc_mulmod_barrett(a, b, q, mu) == a * b % q
            wp = L.orc_shoup_precon(b, q)
            assert L.orc_mulmod_shoup(a, b, wp, q) == a * b % q
        for _ in range(200):
            hi, lo = int(rng.integers(0, 1 << 60)), int(rng.integers(0, 1 << 63)) * 2 + 1
            assert oracle.barrett128(hi, lo, q) == ((hi << 64) | lo) % q
