"""CPU tier: the arithmetic of the tensor-core base conversion (openfhe-development_b200/csrc/baseconv_tc.cu) restated
with numpy / Python integers, step by step as the kernel does it, against the oracle's ApproxSwitchCRTBasis
(dcrtpoly-impl.h:888-932).  This pins the FORMULATION -- byte digits, Montgomery-form matrix, the 16 s32 column sums per
output, the shift-add recombination and the single 64-bit Montgomery step -- independently of the hardware; the CUDA
kernel itself is checked in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle_lib import seeded_residues

MASK64 = (1 << 64) - 1


def _tables(moduli, src, dst):
    q = [int(moduli[i]) for i in src]
    Q = 1
    for v in q:
        Q *= v
    hatinv = np.array([pow((Q // qi) % qi, -1, qi) for qi in q], dtype=np.uint64)
    mat = np.array([[(Q // qi) % int(moduli[j]) for j in dst] for qi in q], dtype=np.uint64)
    return hatinv, mat


def tensor_core_formulation(q_src, hatinv, p_dst, mat, x):
    """out[j][r] exactly as bconv_tc_kernel computes it."""
    nq, N = x.shape
    # y_i = x_i * QHatInv_i mod q_i (Shoup multiply in the kernel; any exact modmul here)
    y = [(x[i].astype(object) * int(hatinv[i])) % int(q_src[i]) for i in range(nq)]
    # operand A: [coefficient][K = 16 limbs x 8 bytes] of u8 (zero padded limbs)
    A = np.zeros((N, 128), dtype=np.int64)
    for i in range(nq):
        yi = np.array(y[i], dtype=np.uint64)
        for a in range(8):
            A[:, 8 * i + a] = ((yi >> np.uint64(8 * a)) & np.uint64(255)).astype(np.int64)
    out = np.zeros((len(p_dst), N), dtype=np.uint64)
    for j, p in enumerate(int(v) for v in p_dst):
        # operand B: 16 columns (digit weights 2^(8c)) of the MONTGOMERY-form matrix column M'[.][j] = M[.][j] 2^64 mod p
        B = np.zeros((128, 16), dtype=np.int64)
        for i in range(nq):
            mp = (int(mat[i][j]) << 64) % p
            for a in range(8):
                for c in range(a, min(a + 8, 16)):
                    B[8 * i + a, c] = (mp >> (8 * (c - a))) & 255
        D = A @ B  # the s32 accumulators in TMEM
        assert D.max() < (1 << 23) and (D[:, 15] == 0).all()
        ninv = (-pow(p, -1, 1 << 64)) & MASK64
        for r in range(N):
            v = [int(t) for t in D[r]]
            e = [v[2 * k] + (v[2 * k + 1] << 8) for k in range(7)] + [v[14]]
            assert all(t < (1 << 32) for t in e)
            a128 = e[0] | (e[2] << 32) | (e[4] << 64) | (e[6] << 96)
            b128 = e[1] | (e[3] << 32) | (e[5] << 64) | (e[7] << 96)
            s = a128 + (b128 << 16)
            assert s < (1 << 124)
            lo, hi = s & MASK64, s >> 64
            u = (lo * ninv) & MASK64
            t = hi + ((u * p) >> 64) + (1 if lo else 0)
            assert t < 2 * p
            out[j, r] = t - p if t >= p else t
    return out


@pytest.mark.parametrize("N,ns,nd,bits", [(64, 15, 7, 59), (32, 16, 3, 59), (32, 1, 2, 59), (64, 5, 4, 30)])
def test_byte_digit_gemm_with_montgomery_step_equals_reference_conversion(oracle, N, ns, nd, bits):
    import ofhe_b200 as m

    moduli, _roots = m.dcrt_params(1024, ns + nd, bits)  # primes = 1 mod 2048, any of them work as CRT bases
    src, dst = list(range(ns)), list(range(ns, ns + nd))
    x = seeded_residues(7 + ns, moduli[src], N)
    x[:, 0] = 0
    x[:, 1] = moduli[src] - np.uint64(1)
    hatinv, mat = _tables(moduli, src, dst)
    want = oracle.approx_switch_crt_basis(moduli[src], hatinv, moduli[dst], mat, x)
    got = tensor_core_formulation(moduli[src], hatinv, moduli[dst], mat, x)
    assert (got == want).all()
