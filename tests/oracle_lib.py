"""ctypes bindings for the two CHECKERS used by the tests (never by the product):

* ``Oracle``    -- oracle/ofhe_oracle.c, the plain-C restatement of the reference's hot path;
* ``Reference`` -- oracle/_ref/libofhe_ref.so, the UNMODIFIED reference compiled from /root/reference
  (oracle/Makefile.ref) behind oracle/ref_shim.cpp.  Present in the build container and on the GPU box
  (the built .so travels with the snapshot); absent => tests that need it skip.

All polynomials are numpy uint64 arrays shaped [limbs, N] (EVALUATION = the reference's bit-reversed order).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libofhe_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libofhe_ref.so")

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build_oracle():
    """Compile the C restatement if needed (gcc only; seconds)."""
    src = os.path.join(ROOT, "oracle", "ofhe_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-f", os.path.join(ROOT, "oracle", "Makefile")], cwd=ROOT)
    return ORACLE_SO


def build_reference():
    """Compile the unmodified reference (only possible where /root/reference exists; minutes)."""
    if os.path.exists(REF_SO):
        return REF_SO
    if not os.path.isdir("/root/reference/src"):
        return None
    subprocess.check_call(["make", "-s", "-j8", "-f", os.path.join(ROOT, "oracle", "Makefile.ref")], cwd=ROOT)
    return REF_SO


def reference_available():
    return os.path.exists(REF_SO)


def _sig(lib, name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


def seeded_residues(seed, moduli, N):
    """Uniform residues in [0, q_l) per limb: the shared synthetic-input generator."""
    moduli = np.asarray(moduli, dtype=np.uint64)
    out = np.empty((len(moduli), N), dtype=np.uint64)
    for l, q in enumerate(moduli):
        rng = np.random.Generator(np.random.PCG64(int(seed) * 1000003 + l))
        out[l] = rng.integers(0, int(q), size=N, dtype=np.uint64)
    return out


class Oracle:
    """The C restatement (oracle/ofhe_oracle.c)."""

    def __init__(self):
        L = self.L = C.CDLL(build_oracle())
        _sig(L, "orc_shoup_precon", C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "orc_mulmod_shoup", C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "orc_barrett_mu", C.c_uint64, C.c_uint64)
        _sig(L, "orc_mulmod_barrett", C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "orc_barrett128_mu", None, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
        _sig(L, "orc_barrett128", C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "orc_is_prime", C.c_int, C.c_uint64)
        for n in ("orc_first_prime", "orc_last_prime"):
            _sig(L, n, C.c_uint64, C.c_uint32, C.c_uint64)
        for n in ("orc_next_prime", "orc_previous_prime", "orc_root_of_unity"):
            _sig(L, n, C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "orc_twiddles", None, C.c_uint64, C.c_uint64, C.c_uint32, u64p, u64p, u64p, u64p,
             C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
        _sig(L, "orc_ntt_fwd", None, C.c_uint64, C.c_uint64, C.c_uint32, u64p)
        _sig(L, "orc_ntt_inv", None, C.c_uint64, C.c_uint64, C.c_uint32, u64p)
        _sig(L, "orc_limb_binop", None, C.c_uint64, C.c_uint32, C.c_int, u64p, u64p, u64p)
        _sig(L, "orc_limb_mul_scalar", None, C.c_uint64, C.c_uint32, u64p, C.c_uint64, u64p)
        _sig(L, "orc_switch_modulus", None, C.c_uint32, C.c_uint64, C.c_uint64, u64p)
        _sig(L, "orc_automorphism_map", None, C.c_uint32, C.c_uint32, u32p)
        _sig(L, "orc_automorphism_eval", None, C.c_uint32, u32p, u64p, u64p)
        _sig(L, "orc_automorphism_coeff", None, C.c_uint32, C.c_uint32, C.c_uint64, u64p, u64p)
        _sig(L, "orc_approx_switch_crt_basis", None, C.c_uint32, C.c_uint32, u64p, u64p, C.c_uint32, u64p, u64p,
             u64p, u64p)
        _sig(L, "orc_switch_crt_basis", None, C.c_uint32, C.c_uint32, u64p, u64p, C.c_uint32, u64p, u64p, u64p, f64p,
             u64p, u64p)
        _sig(L, "orc_scale_and_round", None, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_int, u64p, f64p, u64p, u64p)
        _sig(L, "orc_ckks_create", C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int)
        _sig(L, "orc_hks_create_from_moduli", C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u64p)
        _sig(L, "orc_ckks_destroy", None, C.c_void_p)
        _sig(L, "orc_ckks_shape", None, C.c_void_p, u32p)
        _sig(L, "orc_ckks_moduli", None, C.c_void_p, u64p, u64p)
        for n in ("orc_ckks_PInvModq", "orc_ckks_PModq", "orc_ckks_PHatInvModp", "orc_ckks_PHatModq"):
            _sig(L, n, None, C.c_void_p, u64p)
        _sig(L, "orc_ckks_PartQlHatInvModq", None, C.c_void_p, C.c_uint32, C.c_uint32, u64p)
        _sig(L, "orc_ckks_PartQlHatModp", C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p)
        _sig(L, "orc_ckks_rescale_tables", None, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "orc_ckks_modup", C.c_int, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "orc_ckks_inner_product", None, C.c_void_p, C.c_uint32, C.c_uint32, u64p, u64p, u64p, u64p, u64p)
        _sig(L, "orc_ckks_approx_mod_down", None, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "orc_ckks_keyswitch_core", None, C.c_void_p, C.c_uint32, u64p, u64p, u64p, u64p, u64p)
        _sig(L, "orc_ckks_evalmult", None, C.c_void_p, C.c_uint32, u64p, u64p, u64p, u64p, u64p)
        _sig(L, "orc_ckks_rescale_poly", None, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "orc_ckks_automorphism_ct", None, C.c_void_p, C.c_uint32, C.c_uint32, u64p, u64p, u64p, u64p)
        _sig(L, "orc_ckks_fast_rotation", None, C.c_void_p, C.c_uint32, C.c_uint32, u64p, u64p, u64p, u64p, u64p)
        _sig(L, "orc_ckks_fast_rotation_ext", None, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, u64p, u64p, u64p, u64p,
             u64p)
        _sig(L, "orc_ckks_keyswitch_ext", None, C.c_void_p, C.c_uint32, C.c_int, u64p, u64p)
        _sig(L, "orc_ckks_keyswitch_down", None, C.c_void_p, C.c_uint32, u64p, u64p)

    # --- scalar / number theory -------------------------------------------------------------
    def barrett128(self, hi, lo, q):
        mh, ml = C.c_uint64(), C.c_uint64()
        self.L.orc_barrett128_mu(q, C.byref(mh), C.byref(ml))
        return self.L.orc_barrett128(hi, lo, q, mh.value, ml.value)

    def twiddles(self, q, root, N):
        T, Tp, Ti, Tip = (np.empty(N, dtype=np.uint64) for _ in range(4))
        ni, nip = C.c_uint64(), C.c_uint64()
        self.L.orc_twiddles(q, root, N, T, Tp, Ti, Tip, C.byref(ni), C.byref(nip))
        return T, Tp, Ti, Tip, ni.value, nip.value

    def ntt_fwd(self, q, root, x):
        x = np.ascontiguousarray(x, dtype=np.uint64).copy()
        self.L.orc_ntt_fwd(int(q), int(root), x.shape[-1], x)
        return x

    def ntt_inv(self, q, root, x):
        x = np.ascontiguousarray(x, dtype=np.uint64).copy()
        self.L.orc_ntt_inv(int(q), int(root), x.shape[-1], x)
        return x

    def dcrt_switch_format(self, moduli, roots, data, from_eval):
        out = np.empty_like(data)
        for l in range(data.shape[0]):
            f = self.ntt_inv if from_eval else self.ntt_fwd
            out[l] = f(moduli[l], roots[l], data[l])
        return out

    def dcrt_binop(self, moduli, op, a, b=None):
        out = np.empty_like(a)
        b = a if b is None else b
        for l in range(a.shape[0]):
            self.L.orc_limb_binop(int(moduli[l]), a.shape[1], op, np.ascontiguousarray(a[l]),
                                  np.ascontiguousarray(b[l]), out[l])
        return out

    def dcrt_times_scalars(self, moduli, a, scalars):
        out = np.empty_like(a)
        for l in range(a.shape[0]):
            self.L.orc_limb_mul_scalar(int(moduli[l]), a.shape[1], np.ascontiguousarray(a[l]), int(scalars[l]),
                                       out[l])
        return out

    def switch_modulus(self, qf, qt, x):
        x = np.ascontiguousarray(x, dtype=np.uint64).copy()
        self.L.orc_switch_modulus(x.shape[-1], int(qf), int(qt), x)
        return x

    def automorphism_map(self, N, k):
        m = np.empty(N, dtype=np.uint32)
        self.L.orc_automorphism_map(N, k, m)
        return m

    def dcrt_automorphism(self, k, a):
        m = self.automorphism_map(a.shape[1], k)
        return np.ascontiguousarray(a[:, m])

    def dcrt_automorphism_coeff(self, q, k, a):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        for i in range(a.shape[0]):
            self.L.orc_automorphism_coeff(a.shape[1], k, int(q[i]), a[i], out[i])
        return out

    def approx_switch_crt_basis(self, q, QHatInvModq, p, QHatModp, x):
        q, p = np.ascontiguousarray(q, dtype=np.uint64), np.ascontiguousarray(p, dtype=np.uint64)
        out = np.empty((len(p), x.shape[1]), dtype=np.uint64)
        self.L.orc_approx_switch_crt_basis(x.shape[1], len(q), q, np.ascontiguousarray(QHatInvModq), len(p), p,
                                           np.ascontiguousarray(QHatModp), np.ascontiguousarray(x), out)
        return out

    def switch_crt_basis(self, q, QHatInvModq, p, QHatModp_ji, alphaQModp, qInv, x):
        """a12; QHatModp_ji is [sizeP][sizeQ] (the reference's order for this function)."""
        q, p = np.ascontiguousarray(q, dtype=np.uint64), np.ascontiguousarray(p, dtype=np.uint64)
        out = np.empty((len(p), x.shape[1]), dtype=np.uint64)
        self.L.orc_switch_crt_basis(x.shape[1], len(q), q, np.ascontiguousarray(QHatInvModq), len(p), p,
                                    np.ascontiguousarray(QHatModp_ji), np.ascontiguousarray(alphaQModp),
                                    np.ascontiguousarray(qInv, dtype=np.float64), np.ascontiguousarray(x), out)
        return out

    def scale_and_round(self, o, out_limbs_first, tbl, frac, x):
        o = np.ascontiguousarray(o, dtype=np.uint64)
        sizeO, sizeI = len(o), x.shape[0] - len(o)
        out = np.empty((sizeO, x.shape[1]), dtype=np.uint64)
        self.L.orc_scale_and_round(x.shape[1], sizeI, sizeO, o, int(out_limbs_first), np.ascontiguousarray(tbl),
                                   np.ascontiguousarray(frac, dtype=np.float64), np.ascontiguousarray(x), out)
        return out

    def ckks(self, N, depth, scale_bits, first_bits, dnum=0, scal_tech=0):
        return OracleCkks(self, N, depth, scale_bits, first_bits, dnum, scal_tech)

    def hks_from_moduli(self, N, sizeQ, sizeP, dnum, moduli, roots):
        """Hybrid key-switch context (a16-a18, a20) over a caller-supplied chain, e.g. BFV's Q then P."""
        return OracleCkks(self, N, 0, 0, 0, dnum, 0, from_moduli=(sizeQ, sizeP, moduli, roots))


class _CkksCommon:
    """Shape + table accessors shared by OracleCkks and ReferenceCkks (same method names on purpose)."""

    def _init_shape(self, shape):
        self.N, self.sizeQ, self.sizeP, self.dnum, self.alpha = (int(v) for v in shape[:5])
        self.sizeQP = self.sizeQ + self.sizeP

    def num_digits(self, sizeQl):
        return min((sizeQl + self.alpha - 1) // self.alpha, self.dnum)

    def digit_size(self, sizeQl, part):
        return min(self.alpha, sizeQl - part * self.alpha)

    def moduli_QlP(self, sizeQl):
        return np.concatenate([self.moduli[:sizeQl], self.moduli[self.sizeQ:]])

    def roots_QlP(self, sizeQl):
        return np.concatenate([self.roots[:sizeQl], self.roots[self.sizeQ:]])


class OracleCkks(_CkksCommon):
    def __init__(self, orc, N, depth, scale_bits, first_bits, dnum, scal_tech, from_moduli=None):
        self.o = orc
        self.L = orc.L
        if from_moduli is not None:  # hybrid key-switch context over a given chain: (sizeQ, sizeP, moduli, roots)
            sq, sp, mod, rts = from_moduli
            self.h = self.L.orc_hks_create_from_moduli(N, sq, sp, dnum, np.ascontiguousarray(mod, dtype=np.uint64),
                                                       np.ascontiguousarray(rts, dtype=np.uint64))
        else:
            self.h = self.L.orc_ckks_create(N, depth, scale_bits, first_bits, dnum, scal_tech)
        s = np.zeros(6, dtype=np.uint32)
        self.L.orc_ckks_shape(self.h, s)
        self._init_shape(s)
        self.moduli = np.empty(self.sizeQP, dtype=np.uint64)
        self.roots = np.empty(self.sizeQP, dtype=np.uint64)
        self.L.orc_ckks_moduli(self.h, self.moduli, self.roots)

    def close(self):
        if self.h:
            self.L.orc_ckks_destroy(self.h)
            self.h = None

    def _vec(self, fn, n):
        out = np.empty(n, dtype=np.uint64)
        getattr(self.L, fn)(self.h, out)
        return out

    def PInvModq(self):
        return self._vec("orc_ckks_PInvModq", self.sizeQ)

    def PModq(self):
        return self._vec("orc_ckks_PModq", self.sizeQ)

    def PHatInvModp(self):
        return self._vec("orc_ckks_PHatInvModp", self.sizeP)

    def PHatModq(self):
        return self._vec("orc_ckks_PHatModq", self.sizeP * self.sizeQ).reshape(self.sizeP, self.sizeQ)

    def PartQlHatInvModq(self, part, digit_size):
        out = np.empty(digit_size, dtype=np.uint64)
        self.L.orc_ckks_PartQlHatInvModq(self.h, part, digit_size, out)
        return out

    def PartQlHatModp(self, sizeQl, part):
        sc = self.L.orc_ckks_PartQlHatModp(self.h, sizeQl, part, None, None)
        ds = self.digit_size(sizeQl, part)
        out = np.empty((ds, sc), dtype=np.uint64)
        cm = np.empty(sc, dtype=np.uint64)
        self.L.orc_ckks_PartQlHatModp(self.h, sizeQl, part, out.ctypes.data, cm.ctypes.data)
        return out, cm

    def rescale_tables(self, sizeQl):
        a = np.empty(sizeQl - 1, dtype=np.uint64)
        b = np.empty(sizeQl - 1, dtype=np.uint64)
        self.L.orc_ckks_rescale_tables(self.h, sizeQl, a, b)
        return a, b

    def modup(self, sizeQl, poly):
        nd = self.num_digits(sizeQl)
        out = np.empty((nd, sizeQl + self.sizeP, self.N), dtype=np.uint64)
        assert self.L.orc_ckks_modup(self.h, sizeQl, np.ascontiguousarray(poly), out) == nd
        return out

    def inner_product(self, sizeQl, digits, evk_a, evk_b):
        e0 = np.empty((sizeQl + self.sizeP, self.N), dtype=np.uint64)
        e1 = np.empty_like(e0)
        self.L.orc_ckks_inner_product(self.h, sizeQl, digits.shape[0], np.ascontiguousarray(digits), evk_a, evk_b,
                                      e0, e1)
        return e0, e1

    def approx_mod_down(self, sizeQl, x):
        out = np.empty((sizeQl, self.N), dtype=np.uint64)
        self.L.orc_ckks_approx_mod_down(self.h, sizeQl, np.ascontiguousarray(x), out)
        return out

    def keyswitch_core(self, sizeQl, poly, evk_a, evk_b):
        o0 = np.empty((sizeQl, self.N), dtype=np.uint64)
        o1 = np.empty_like(o0)
        self.L.orc_ckks_keyswitch_core(self.h, sizeQl, np.ascontiguousarray(poly), evk_a, evk_b, o0, o1)
        return o0, o1

    def evalmult(self, sizeQl, ct1, ct2, evk_a, evk_b):
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.L.orc_ckks_evalmult(self.h, sizeQl, np.ascontiguousarray(ct1), np.ascontiguousarray(ct2), evk_a, evk_b,
                                 out)
        return out

    def rescale_poly(self, sizeQl, x):
        out = np.empty((sizeQl - 1, self.N), dtype=np.uint64)
        self.L.orc_ckks_rescale_poly(self.h, sizeQl, np.ascontiguousarray(x), out)
        return out

    def automorphism_ct(self, sizeQl, k, ct, evk_a, evk_b):
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.L.orc_ckks_automorphism_ct(self.h, sizeQl, k, np.ascontiguousarray(ct), evk_a, evk_b, out)
        return out

    # hoisted rotations: digits = self.modup(sizeQl, ct[1])
    def fast_rotation(self, sizeQl, k, ct, digits, evk_a, evk_b):
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.L.orc_ckks_fast_rotation(self.h, sizeQl, k, np.ascontiguousarray(ct), np.ascontiguousarray(digits), evk_a,
                                      evk_b, out)
        return out

    def fast_rotation_ext(self, sizeQl, k, add_first, ct, digits, evk_a, evk_b):
        out = np.empty((2, sizeQl + self.sizeP, self.N), dtype=np.uint64)
        self.L.orc_ckks_fast_rotation_ext(self.h, sizeQl, k, int(add_first), np.ascontiguousarray(ct),
                                          np.ascontiguousarray(digits), evk_a, evk_b, out)
        return out

    def keyswitch_ext(self, sizeQl, add_first, ct):
        out = np.empty((2, sizeQl + self.sizeP, self.N), dtype=np.uint64)
        self.L.orc_ckks_keyswitch_ext(self.h, sizeQl, int(add_first), np.ascontiguousarray(ct), out)
        return out

    def keyswitch_down(self, sizeQl, ct_ext):
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.L.orc_ckks_keyswitch_down(self.h, sizeQl, np.ascontiguousarray(ct_ext), out)
        return out

    def bsgs_level(self, sizeQl, k_in, k_out, keys_in, keys_out, A, skip, ct, binop):
        """One level of the double-hoisted BSGS linear transform, composed from the oracle's primitives in the order
        of FHECKKSRNS::EvalCoeffsToSlots' loop body (ckksrns-fhe.cpp:1934-1987).  k_*: automorphism indices (0 = no
        rotation), keys_*: (evk_a, evk_b) per step, A: [b*g][sizeQl+sizeP][N], binop(q, op, a, b): limb-wise op."""
        g, b = len(k_in), len(k_out)
        qlp = np.concatenate([self.moduli[:sizeQl], self.moduli[self.sizeQ:]])
        ql = self.moduli[:sizeQl]
        digits = self.modup(sizeQl, ct[1])
        fr = [self.fast_rotation_ext(sizeQl, k_in[j], True, ct, digits, *keys_in[j]) if k_in[j] else
              self.keyswitch_ext(sizeQl, True, ct) for j in range(g)]
        first = outer = None
        for i in range(b):
            G = g * i
            inner = np.stack([binop(qlp, 2, fr[0][e], A[G]) for e in range(2)])
            for j in range(1, g):
                if G + j != skip:
                    inner = np.stack([binop(qlp, 0, inner[e], binop(qlp, 2, fr[j][e], A[G + j])) for e in range(2)])
            if i == 0:
                first = self.keyswitch_down(sizeQl, inner)[0]
                outer = inner.copy()
                outer[0] = 0
            elif k_out[i]:
                d = self.keyswitch_down(sizeQl, inner)
                m = np.empty(self.N, dtype=np.uint32)
                self.L.orc_automorphism_map(self.N, k_out[i], m)
                first = binop(ql, 0, first, np.ascontiguousarray(d[0][:, m]))
                rot = self.fast_rotation_ext(sizeQl, k_out[i], False, d, self.modup(sizeQl, d[1]), *keys_out[i])
                outer = np.stack([binop(qlp, 0, outer[e], rot[e]) for e in range(2)])
            else:
                first = binop(ql, 0, first, self.keyswitch_down(sizeQl, inner)[0])
                outer[1] = binop(qlp, 0, outer[1], inner[1])
        res = self.keyswitch_down(sizeQl, outer)
        res[0] = binop(ql, 0, res[0], first)
        return res


class Reference:
    """The unmodified reference behind oracle/ref_shim.cpp."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = self.L = C.CDLL(REF_SO)
        _sig(L, "ref_last_error", C.c_char_p)
        _sig(L, "ref_set_threads", None, C.c_int)
        _sig(L, "ref_get_threads", C.c_int)
        _sig(L, "ref_root_of_unity", C.c_uint64, C.c_uint32, C.c_uint64)
        for n in ("ref_first_prime", "ref_last_prime"):
            _sig(L, n, C.c_uint64, C.c_uint32, C.c_uint64)
        for n in ("ref_next_prime", "ref_previous_prime"):
            _sig(L, n, C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "ref_modmul_barrett", C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "ref_prep_modmul_const", C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "ref_modmul_const", C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "ref_barrett128", C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)
        _sig(L, "ref_ntt_fwd", C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, u64p)
        _sig(L, "ref_ntt_inv", C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, u64p)
        _sig(L, "ref_twiddles", C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, u64p, u64p, u64p, u64p,
             C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
        _sig(L, "ref_dcrt_binop", C.c_int, C.c_uint32, C.c_uint32, u64p, u64p, C.c_int, u64p, u64p, u64p)
        _sig(L, "ref_dcrt_times_scalars", C.c_int, C.c_uint32, C.c_uint32, u64p, u64p, u64p, u64p, u64p)
        _sig(L, "ref_dcrt_switch_format", C.c_int, C.c_uint32, C.c_uint32, u64p, u64p, C.c_int, u64p)
        _sig(L, "ref_dcrt_serialize", C.c_size_t, C.c_uint32, C.c_uint32, u64p, u64p, C.c_int, u64p, C.c_int,
             C.c_void_p, C.c_size_t)
        _sig(L, "ref_dcrt_automorphism", C.c_int, C.c_uint32, C.c_uint32, u64p, u64p, C.c_uint32, u64p, u64p)
        _sig(L, "ref_ckks_bsgs_level", C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
             np.ctypeslib.ndpointer(np.int32, flags="C"), np.ctypeslib.ndpointer(np.int32, flags="C"), u64p, C.c_uint32,
             u64p, u64p)
        _sig(L, "ref_dcrt_automorphism_coeff", C.c_int, C.c_uint32, C.c_uint32, u64p, u64p, C.c_uint32, u64p, u64p)
        _sig(L, "ref_automorphism_map", C.c_int, C.c_uint32, C.c_uint32, u32p)
        _sig(L, "ref_switch_modulus", C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, u64p)
        _sig(L, "ref_time_switch_format", C.c_double, C.c_uint32, C.c_uint32, u64p, u64p, u64p, C.c_int)
        _sig(L, "ref_ckks_create", C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int)
        _sig(L, "ref_ckks_destroy", None, C.c_void_p)
        _sig(L, "ref_ckks_shape", None, C.c_void_p, u32p)
        _sig(L, "ref_ckks_moduli", None, C.c_void_p, u64p, u64p)
        for n in ("ref_ckks_PInvModq", "ref_ckks_PModq", "ref_ckks_PHatInvModp", "ref_ckks_PHatModq"):
            _sig(L, n, None, C.c_void_p, u64p)
        _sig(L, "ref_ckks_PartQlHatInvModq", None, C.c_void_p, C.c_uint32, C.c_uint32, u64p)
        _sig(L, "ref_ckks_PartQlHatModp", C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p)
        _sig(L, "ref_ckks_rescale_tables", None, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "ref_ckks_evalmult_key", C.c_int, C.c_void_p, u64p, u64p)
        _sig(L, "ref_ckks_rotate_key", C.c_uint32, C.c_void_p, C.c_int32, u64p, u64p)
        _sig(L, "ref_ckks_approx_switch_crt_basis_digit", C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, u64p, u64p)
        _sig(L, "ref_ckks_modup", C.c_int, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "ref_ckks_approx_mod_down", C.c_int, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "ref_ckks_keyswitch_core", C.c_int, C.c_void_p, C.c_uint32, u64p, u64p, u64p)
        _sig(L, "ref_ckks_rescale_poly", C.c_int, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "ref_ckks_evalmult", C.c_int, C.c_void_p, C.c_uint32, u64p, u64p, u64p)
        _sig(L, "ref_ckks_time_evalmult", C.c_double, C.c_void_p, C.c_uint32, u64p, u64p, C.c_int, C.c_int)
        _sig(L, "ref_ckks_rescale", C.c_int, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "ref_ckks_rotate", C.c_int, C.c_void_p, C.c_uint32, C.c_int32, u64p, u64p)
        _sig(L, "ref_ckks_fast_rotate", C.c_int, C.c_void_p, C.c_uint32, C.c_int32, u64p, u64p)
        _sig(L, "ref_ckks_fast_rotate_ext", C.c_int, C.c_void_p, C.c_uint32, C.c_int32, C.c_int, u64p, u64p)
        _sig(L, "ref_ckks_keyswitch_ext", C.c_int, C.c_void_p, C.c_uint32, C.c_int, u64p, u64p)
        _sig(L, "ref_ckks_keyswitch_down", C.c_int, C.c_void_p, C.c_uint32, u64p, u64p)
        _sig(L, "ref_bfv_create", C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32)
        _sig(L, "ref_bfv_destroy", None, C.c_void_p)
        _sig(L, "ref_bfv_shape", None, C.c_void_p, u32p, C.c_void_p, C.c_void_p)
        _sig(L, "ref_bfv_table", C.c_int, C.c_void_p, C.c_int, C.c_void_p)
        _sig(L, "ref_bfv_table_dims", None, C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32))
        for n in ("ref_bfv_expand_crt_basis", "ref_bfv_switch_crt_basis", "ref_bfv_fast_expand_p_over_q",
                  "ref_bfv_scale_and_round"):
            _sig(L, n, C.c_int, C.c_void_p, u64p, u64p)
        _sig(L, "ref_bfv_evalmult_norelin", C.c_int, C.c_void_p, u64p, u64p, u64p)
        _sig(L, "ref_bfv_ks_shape", None, C.c_void_p, u32p, C.c_void_p, C.c_void_p)
        _sig(L, "ref_bfv_evalmult_key", C.c_int, C.c_void_p, u64p, u64p)
        _sig(L, "ref_bfv_evalmult", C.c_int, C.c_void_p, u64p, u64p, u64p)
        _sig(L, "ref_ckks_encrypt", C.c_int, C.c_void_p, f64p, C.c_uint32, u64p)
        _sig(L, "ref_ckks_decrypt", C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, u64p, f64p, C.c_uint32)

    def _ok(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.ref_last_error().decode())

    def twiddles(self, q, root, N):
        T, Tp, Ti, Tip = (np.empty(N, dtype=np.uint64) for _ in range(4))
        ni, nip = C.c_uint64(), C.c_uint64()
        self._ok(self.L.ref_twiddles(q, root, N, T, Tp, Ti, Tip, C.byref(ni), C.byref(nip)))
        return T, Tp, Ti, Tip, ni.value, nip.value

    def ntt_fwd(self, q, root, x):
        x = np.ascontiguousarray(x, dtype=np.uint64).copy()
        self._ok(self.L.ref_ntt_fwd(int(q), int(root), x.shape[-1], x))
        return x

    def ntt_inv(self, q, root, x):
        x = np.ascontiguousarray(x, dtype=np.uint64).copy()
        self._ok(self.L.ref_ntt_inv(int(q), int(root), x.shape[-1], x))
        return x

    def dcrt_switch_format(self, moduli, roots, data, from_eval):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        self._ok(self.L.ref_dcrt_switch_format(d.shape[1], d.shape[0], np.ascontiguousarray(moduli),
                                               np.ascontiguousarray(roots), int(from_eval), d))
        return d

    def dcrt_serialize(self, moduli, roots, evaluation, data, json=False):
        """Serial::Serialize(DCRTPoly, BINARY | JSON) of a polynomial with these towers -> bytes."""
        data = np.ascontiguousarray(data, dtype=np.uint64)
        L, N = data.shape
        buf = np.zeros(L * (N * (24 if json else 8) + 1024) + 8192, dtype=np.uint8)
        n = self.L.ref_dcrt_serialize(N, L, np.ascontiguousarray(moduli, dtype=np.uint64),
                                      np.ascontiguousarray(roots, dtype=np.uint64), int(evaluation), data, int(json),
                                      buf.ctypes.data, buf.size)
        if n == 0:
            raise RuntimeError(self.L.ref_last_error().decode() or "serialization failed")
        return buf[:n].tobytes()

    def dcrt_binop(self, moduli, roots, op, a, b=None):
        out = np.empty_like(a)
        b = a if b is None else b
        self._ok(self.L.ref_dcrt_binop(a.shape[1], a.shape[0], np.ascontiguousarray(moduli),
                                       np.ascontiguousarray(roots), op, np.ascontiguousarray(a),
                                       np.ascontiguousarray(b), out))
        return out

    def dcrt_times_scalars(self, moduli, roots, a, scalars):
        out = np.empty_like(a)
        self._ok(self.L.ref_dcrt_times_scalars(a.shape[1], a.shape[0], np.ascontiguousarray(moduli),
                                               np.ascontiguousarray(roots), np.ascontiguousarray(a),
                                               np.ascontiguousarray(scalars, dtype=np.uint64), out))
        return out

    def dcrt_automorphism(self, moduli, roots, k, a):
        out = np.empty_like(a)
        self._ok(self.L.ref_dcrt_automorphism(a.shape[1], a.shape[0], np.ascontiguousarray(moduli),
                                              np.ascontiguousarray(roots), k, np.ascontiguousarray(a), out))
        return out

    def dcrt_automorphism_coeff(self, moduli, roots, k, a):
        out = np.empty_like(a)
        self._ok(self.L.ref_dcrt_automorphism_coeff(a.shape[1], a.shape[0], np.ascontiguousarray(moduli),
                                                    np.ascontiguousarray(roots), k, np.ascontiguousarray(a), out))
        return out

    def automorphism_map(self, N, k):
        m = np.empty(N, dtype=np.uint32)
        self._ok(self.L.ref_automorphism_map(N, k, m))
        return m

    def switch_modulus(self, qf, qt, x):
        x = np.ascontiguousarray(x, dtype=np.uint64).copy()
        self._ok(self.L.ref_switch_modulus(x.shape[-1], int(qf), int(qt), x))
        return x

    def ckks(self, N, depth, scale_bits, first_bits, dnum=0, scal_tech=0):
        return ReferenceCkks(self, N, depth, scale_bits, first_bits, dnum, scal_tech)

    def bfv(self, N, depth, scale_bits=60, ptm=65537, dnum=0):
        return ReferenceBfv(self, N, depth, scale_bits, ptm, dnum)


class ReferenceCkks(_CkksCommon):
    def __init__(self, ref, N, depth, scale_bits, first_bits, dnum, scal_tech):
        self.r = ref
        self.L = ref.L
        self.h = self.L.ref_ckks_create(N, depth, scale_bits, first_bits, dnum, scal_tech)
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        s = np.zeros(6, dtype=np.uint32)
        self.L.ref_ckks_shape(self.h, s)
        self._init_shape(s)
        self.moduli = np.empty(self.sizeQP, dtype=np.uint64)
        self.roots = np.empty(self.sizeQP, dtype=np.uint64)
        self.L.ref_ckks_moduli(self.h, self.moduli, self.roots)

    def close(self):
        if self.h:
            self.L.ref_ckks_destroy(self.h)
            self.h = None

    def _vec(self, fn, n):
        out = np.empty(n, dtype=np.uint64)
        getattr(self.L, fn)(self.h, out)
        return out

    def PInvModq(self):
        return self._vec("ref_ckks_PInvModq", self.sizeQ)

    def PModq(self):
        return self._vec("ref_ckks_PModq", self.sizeQ)

    def PHatInvModp(self):
        return self._vec("ref_ckks_PHatInvModp", self.sizeP)

    def PHatModq(self):
        return self._vec("ref_ckks_PHatModq", self.sizeP * self.sizeQ).reshape(self.sizeP, self.sizeQ)

    def PartQlHatInvModq(self, part, digit_size):
        out = np.empty(digit_size, dtype=np.uint64)
        self.L.ref_ckks_PartQlHatInvModq(self.h, part, digit_size, out)
        return out

    def PartQlHatModp(self, sizeQl, part):
        sc = self.L.ref_ckks_PartQlHatModp(self.h, sizeQl, part, None, None)
        ds = self.digit_size(sizeQl, part)
        out = np.empty((ds, sc), dtype=np.uint64)
        cm = np.empty(sc, dtype=np.uint64)
        self.L.ref_ckks_PartQlHatModp(self.h, sizeQl, part, out.ctypes.data, cm.ctypes.data)
        return out, cm

    def rescale_tables(self, sizeQl):
        a = np.empty(sizeQl - 1, dtype=np.uint64)
        b = np.empty(sizeQl - 1, dtype=np.uint64)
        self.L.ref_ckks_rescale_tables(self.h, sizeQl, a, b)
        return a, b

    def evalmult_key(self):
        a = np.empty((self.dnum, self.sizeQP, self.N), dtype=np.uint64)
        b = np.empty_like(a)
        self.r._ok(self.L.ref_ckks_evalmult_key(self.h, a, b))
        return a, b

    def rotate_key(self, index):
        a = np.empty((self.dnum, self.sizeQP, self.N), dtype=np.uint64)
        b = np.empty_like(a)
        k = self.L.ref_ckks_rotate_key(self.h, index, a, b)
        if k == 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        return k, a, b

    def approx_switch_crt_basis_digit(self, sizeQl, part, x):
        sc = self.L.ref_ckks_PartQlHatModp(self.h, sizeQl, part, None, None)
        out = np.empty((sc, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_approx_switch_crt_basis_digit(self.h, sizeQl, part, np.ascontiguousarray(x), out))
        return out

    def modup(self, sizeQl, poly):
        nd = self.num_digits(sizeQl)
        out = np.empty((nd, sizeQl + self.sizeP, self.N), dtype=np.uint64)
        n = self.L.ref_ckks_modup(self.h, sizeQl, np.ascontiguousarray(poly), out)
        if n != nd:
            raise RuntimeError(self.L.ref_last_error().decode())
        return out

    def approx_mod_down(self, sizeQl, x):
        out = np.empty((sizeQl, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_approx_mod_down(self.h, sizeQl, np.ascontiguousarray(x), out))
        return out

    def keyswitch_core(self, sizeQl, poly):
        o0 = np.empty((sizeQl, self.N), dtype=np.uint64)
        o1 = np.empty_like(o0)
        self.r._ok(self.L.ref_ckks_keyswitch_core(self.h, sizeQl, np.ascontiguousarray(poly), o0, o1))
        return o0, o1

    def rescale_poly(self, sizeQl, x):
        out = np.empty((sizeQl - 1, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_rescale_poly(self.h, sizeQl, np.ascontiguousarray(x), out))
        return out

    def evalmult(self, sizeQl, ct1, ct2):
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_evalmult(self.h, sizeQl, np.ascontiguousarray(ct1), np.ascontiguousarray(ct2),
                                            out))
        return out

    def time_evalmult(self, sizeQl, ct1, ct2, warmup, reps):
        s = self.L.ref_ckks_time_evalmult(self.h, sizeQl, np.ascontiguousarray(ct1), np.ascontiguousarray(ct2),
                                          warmup, reps)
        if s < 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        return s

    def rescale(self, sizeQl, ct):
        out = np.empty((2, sizeQl - 1, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_rescale(self.h, sizeQl, np.ascontiguousarray(ct), out))
        return out

    def rotate(self, sizeQl, index, ct):
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_rotate(self.h, sizeQl, index, np.ascontiguousarray(ct), out))
        return out

    def fast_rotate(self, sizeQl, index, ct):
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_fast_rotate(self.h, sizeQl, index, np.ascontiguousarray(ct), out))
        return out

    def fast_rotate_ext(self, sizeQl, index, add_first, ct):
        out = np.empty((2, sizeQl + self.sizeP, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_fast_rotate_ext(self.h, sizeQl, index, int(add_first), np.ascontiguousarray(ct), out))
        return out

    def keyswitch_ext(self, sizeQl, add_first, ct):
        out = np.empty((2, sizeQl + self.sizeP, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_keyswitch_ext(self.h, sizeQl, int(add_first), np.ascontiguousarray(ct), out))
        return out

    def keyswitch_down(self, sizeQl, ct_ext):
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_keyswitch_down(self.h, sizeQl, np.ascontiguousarray(ct_ext), out))
        return out

    def bsgs_level(self, sizeQl, rot_in, rot_out, A, skip, ct):
        """rot_*: ROTATION indices (keys generated beforehand with rotate_key)."""
        out = np.empty((2, sizeQl, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_bsgs_level(self.h, sizeQl, len(rot_in), len(rot_out),
                                              np.ascontiguousarray(rot_in, dtype=np.int32),
                                              np.ascontiguousarray(rot_out, dtype=np.int32), np.ascontiguousarray(A),
                                              skip, np.ascontiguousarray(ct), out))
        return out

    def encrypt(self, vals):
        v = np.ascontiguousarray(vals, dtype=np.float64)
        out = np.empty((2, self.sizeQ, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_ckks_encrypt(self.h, v, len(v), out))
        return out

    def decrypt(self, sizeQl, deg, ct, n):
        out = np.empty(n, dtype=np.float64)
        self.r._ok(self.L.ref_ckks_decrypt(self.h, sizeQl, deg, np.ascontiguousarray(ct), out, n))
        return out


BFV_TABLES = {"QlHatInvModq": 0, "QlHatModr": 1, "alphaQlModr": 2, "qInv": 3, "mNegRlQlHatInvModq": 4, "qInvModr": 5,
              "RlHatInvModr": 6, "RlHatModq": 7, "alphaRlModq": 8, "rInv": 9, "tQlSlHatInvModsDivsModq": 10,
              "tQlSlHatInvModsDivsFrac": 11}
_BFV_DOUBLE = {3, 9, 11}


class ReferenceBfv:
    """BFV (HPSPOVERQ) context of the unmodified reference: tables and DCRTPoly calls of rows a12 / a13."""

    def __init__(self, ref, N, depth, scale_bits, ptm, dnum):
        self.r, self.L = ref, ref.L
        self.h = self.L.ref_bfv_create(N, depth, scale_bits, ptm, dnum)
        if not self.h:
            raise RuntimeError(self.L.ref_last_error().decode())
        s = np.zeros(3, dtype=np.uint32)
        self.L.ref_bfv_shape(self.h, s, None, None)
        self.N, self.sizeQ, self.sizeR = (int(v) for v in s)
        self.moduli = np.empty(self.sizeQ + self.sizeR, dtype=np.uint64)
        self.roots = np.empty_like(self.moduli)
        self.L.ref_bfv_shape(self.h, s, self.moduli.ctypes.data, self.roots.ctypes.data)

    def close(self):
        if self.h:
            self.L.ref_bfv_destroy(self.h)
            self.h = None

    def table(self, name):
        w = BFV_TABLES[name]
        rows, cols = C.c_uint32(), C.c_uint32()
        self.L.ref_bfv_table_dims(self.h, w, C.byref(rows), C.byref(cols))
        shape = (rows.value,) if cols.value == 0 else (rows.value, cols.value)
        out = np.empty(shape, dtype=np.float64 if w in _BFV_DOUBLE else np.uint64)
        self.r._ok(self.L.ref_bfv_table(self.h, w, out.ctypes.data))
        return out

    def tables(self):
        return {k: self.table(k) for k in BFV_TABLES}

    def _call(self, fn, x, n_out):
        out = np.empty((n_out, self.N), dtype=np.uint64)
        self.r._ok(getattr(self.L, fn)(self.h, np.ascontiguousarray(x), out))
        return out

    def expand_crt_basis(self, x):
        return self._call("ref_bfv_expand_crt_basis", x, self.sizeQ + self.sizeR)

    def switch_crt_basis(self, x):
        return self._call("ref_bfv_switch_crt_basis", x, self.sizeR)

    def fast_expand_p_over_q(self, x):
        return self._call("ref_bfv_fast_expand_p_over_q", x, self.sizeQ + self.sizeR)

    def scale_and_round(self, x):
        return self._call("ref_bfv_scale_and_round", x, self.sizeQ)

    def evalmult_norelin(self, ct1, ct2):
        out = np.empty((3, self.sizeQ, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_bfv_evalmult_norelin(self.h, np.ascontiguousarray(ct1), np.ascontiguousarray(ct2), out))
        return out

    def ks_shape(self):
        """HYBRID key-switching side: dict(sizeP, dnum, alpha, moduliP, rootsP)."""
        s = np.zeros(3, dtype=np.uint32)
        self.L.ref_bfv_ks_shape(self.h, s, None, None)
        mp, rp = np.empty(int(s[0]), dtype=np.uint64), np.empty(int(s[0]), dtype=np.uint64)
        self.L.ref_bfv_ks_shape(self.h, s, mp.ctypes.data, rp.ctypes.data)
        return dict(sizeP=int(s[0]), dnum=int(s[1]), alpha=int(s[2]), moduliP=mp, rootsP=rp)

    def evalmult_key(self):
        ks = self.ks_shape()
        a = np.empty((ks["dnum"], self.sizeQ + ks["sizeP"], self.N), dtype=np.uint64)
        b = np.empty_like(a)
        self.r._ok(self.L.ref_bfv_evalmult_key(self.h, a, b))
        return a, b

    def evalmult(self, ct1, ct2):
        out = np.empty((2, self.sizeQ, self.N), dtype=np.uint64)
        self.r._ok(self.L.ref_bfv_evalmult(self.h, np.ascontiguousarray(ct1), np.ascontiguousarray(ct2), out))
        return out
