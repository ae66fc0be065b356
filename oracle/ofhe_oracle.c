/* oracle/ofhe_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's RNS/NTT + hybrid key-switch hot
 * path (openfhe-development v1.5.1), one function per SURVEY.md section-8 row,
 * each citing the reference file:line it follows (paths relative to
 * /root/reference).  It exists to CHECK the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may compile, load or call it.  The product (openfhe-development_b200/) never
 * links it and has no CPU fallback.
 *
 * PARITY PIN: this restatement is pinned against (i) the reference's own
 * known-answer test UnitTestTransform.cpp:52-90 (q=113) and the round-trip
 * vectors of UnitTestNTT.cpp:53-119 (tests/golden/), and (ii) the UNMODIFIED
 * reference compiled here into oracle/_ref (oracle/Makefile.ref), residue for
 * residue, by tests/test_oracle_vs_reference.py; golden vectors generated from
 * that build are committed under tests/golden/ so the pin also holds on boxes
 * where oracle/_ref is absent.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp oracle/ofhe_oracle.c -o oracle/_build/libofhe_oracle.so
 */
#include "ofhe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef orc_u128 u128;

/* ------------------------------------------------------------------ a1: Shoup constant multiplication
 * ubintnat.h:1437-1486: PrepModMulConst  w' = floor(w * 2^64 / q);
 * ModMulFastConst  qhat = hi64(x*w') + 1;  y = (int64)(x*w - qhat*q);  y < 0 ? y + q : y            */
uint64_t orc_shoup_precon(uint64_t w, uint64_t q) {
    return (uint64_t)(((u128)w << 64) / q);
}
uint64_t orc_mulmod_shoup(uint64_t x, uint64_t w, uint64_t wp, uint64_t q) {
    uint64_t qhat = (uint64_t)(((u128)x * wp) >> 64) + 1;
    int64_t y     = (int64_t)(x * w - qhat * q);
    return (uint64_t)(y >= 0 ? y : y + (int64_t)q);
}

/* ------------------------------------------------------------------ a2: Barrett multiplication
 * ubintnat.h:642-647 ComputeMu: mu = floor(2^(2k+3) / q), k = bit length of q;
 * ubintnat.h:1389-1401 ModMulFastEq: n = k-2; p = a*b; t = ((p >> n) * mu) >> (n+7); r = lo(p) - lo(t*q)  */
static uint32_t msb64(uint64_t x) { /* nbtheory.h:169-181 GetMSB = bit length */
    return x ? 64 - (uint32_t)__builtin_clzll(x) : 0;
}
uint64_t orc_barrett_mu(uint64_t q) {
    return (uint64_t)(((u128)1 << (2 * msb64(q) + 3)) / q);
}
uint64_t orc_mulmod_barrett(uint64_t a, uint64_t b, uint64_t q, uint64_t mu) {
    u128 p     = (u128)a * b;
    int64_t n  = (int64_t)msb64(q) - 2;
    u128 t     = (u128)(uint64_t)(p >> n) * mu; /* RShiftD keeps the low word of the shifted product */
    uint64_t r = (uint64_t)p - q * (uint64_t)(t >> (n + 7));
    if (r >= q)
        r -= q;
    return r;
}

/* ------------------------------------------------------------------ a3: 128 -> 64 Barrett
 * utilities-int.h:60-100 BarrettUint128ModUint64 with mu = floor(2^128 / q)
 * (rns-cryptoparameters.cpp:283-288 builds mu with BigInteger; q is never a power of two so
 * floor(2^128/q) == floor((2^128-1)/q)).                                                             */
void orc_barrett128_mu(uint64_t q, uint64_t* mu_hi, uint64_t* mu_lo) {
    u128 mu = (~(u128)0) / q;
    *mu_hi  = (uint64_t)(mu >> 64);
    *mu_lo  = (uint64_t)mu;
}
uint64_t orc_barrett128(uint64_t a_hi, uint64_t a_lo, uint64_t q, uint64_t mu_hi, uint64_t mu_lo) {
    uint64_t left_hi = (uint64_t)(((u128)a_lo * mu_lo) >> 64);
    u128 middle      = (u128)a_lo * mu_hi;
    uint64_t mid_lo = (uint64_t)middle, mid_hi = (uint64_t)(middle >> 64);
    uint64_t tmp1  = mid_lo + left_hi;
    uint64_t carry = tmp1 < mid_lo;
    uint64_t tmp2  = mid_hi + carry;
    middle         = (u128)a_hi * mu_lo;
    mid_lo         = (uint64_t)middle;
    mid_hi         = (uint64_t)(middle >> 64);
    carry          = (uint64_t)(mid_lo + tmp1) < mid_lo;
    left_hi        = mid_hi + carry;
    tmp1           = a_hi * mu_hi + tmp2 + left_hi;
    uint64_t r     = a_lo - tmp1 * q;
    while (r >= q)
        r -= q;
    return r;
}

uint64_t orc_powmod(uint64_t b, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    b %= q;
    while (e) {
        if (e & 1)
            r = (uint64_t)((u128)r * b % q);
        b = (uint64_t)((u128)b * b % q);
        e >>= 1;
    }
    return r;
}
uint64_t orc_invmod(uint64_t a, uint64_t q) { /* q prime */
    return orc_powmod(a, q - 2, q);
}
static uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) {
    return (uint64_t)((u128)a * b % q);
}

/* ------------------------------------------------------------------ a8 (host side): primes and the minimal root
 * nbtheory-impl.h:329-393 FirstPrime/LastPrime/NextPrime/PreviousPrime step by m from 2^bits;
 * the reference's Miller-Rabin is randomised, a deterministic base set decides identically below 2^64. */
int orc_is_prime(uint64_t n) {
    static const uint64_t bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2)
        return 0;
    for (int i = 0; i < 12; ++i) {
        if (n == bases[i])
            return 1;
        if (n % bases[i] == 0)
            return 0;
    }
    uint64_t d = n - 1;
    int s      = 0;
    while (!(d & 1)) {
        d >>= 1;
        ++s;
    }
    for (int i = 0; i < 12; ++i) {
        uint64_t x = orc_powmod(bases[i], d, n);
        if (x == 1 || x == n - 1)
            continue;
        int comp = 1;
        for (int r = 1; r < s; ++r) {
            x = mulmod(x, x, n);
            if (x == n - 1) {
                comp = 0;
                break;
            }
        }
        if (comp)
            return 0;
    }
    return 1;
}
uint64_t orc_first_prime(uint32_t bits, uint64_t m) { /* nbtheory-impl.h:329-347 */
    uint64_t q = (uint64_t)1 << bits, r = q % m, qn = q + 1 - r;
    if (r > 0)
        qn += m;
    while (!orc_is_prime(qn))
        qn += m;
    return qn;
}
uint64_t orc_last_prime(uint32_t bits, uint64_t m) { /* nbtheory-impl.h:349-373 */
    uint64_t q = (uint64_t)1 << bits, r = q % m, qn = q + 1 - r;
    if (r < 2)
        qn -= m;
    while (!orc_is_prime(qn))
        qn -= m;
    return qn;
}
uint64_t orc_next_prime(uint64_t q, uint64_t m) { /* :375-383 */
    uint64_t qn = q + m;
    while (!orc_is_prime(qn))
        qn += m;
    return qn;
}
uint64_t orc_previous_prime(uint64_t q, uint64_t m) { /* :385-393 */
    uint64_t qn = q - m;
    while (!orc_is_prime(qn))
        qn -= m;
    return qn;
}
/* nbtheory-impl.h:183-231 RootOfUnity: take any primitive m-th root, then return the numerically
 * smallest of all its odd powers (m a power of two), so the answer does not depend on the start. */
uint64_t orc_root_of_unity(uint64_t m, uint64_t q) {
    uint64_t r = 0;
    for (uint64_t g = 2;; ++g) {
        r = orc_powmod(g, (q - 1) / m, q);
        if (orc_powmod(r, m / 2, q) == q - 1)
            break;
    }
    uint64_t r2 = mulmod(r, r, q), x = r, best = r;
    for (uint64_t i = 1; i < m / 2; ++i) {
        x = mulmod(x, r2, q);
        if (x < best)
            best = x;
    }
    return best;
}

/* ------------------------------------------------------------------ a5: twiddle tables
 * transformnat-impl.h:714-756 PreCompute: T[brev(i)] = psi^i, Tinv[brev(i)] = psi^-i, Shoup constants of
 * both, and n^-1 (entry msb = log2 N of the cyclotomic-order-inverse table, use sites :685,707).        */
static uint32_t brev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; ++i)
        r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}
static uint32_t ilog2(uint32_t n) {
    uint32_t l = 0;
    while ((1u << l) < n)
        ++l;
    return l;
}
void orc_twiddles(uint64_t q, uint64_t root, uint32_t N, uint64_t* T, uint64_t* Tp, uint64_t* Ti, uint64_t* Tip,
                  uint64_t* nInv, uint64_t* nInvPrecon) {
    uint32_t lg      = ilog2(N);
    uint64_t rootInv = orc_invmod(root, q), x = 1, xi = 1;
    for (uint32_t i = 0; i < N; ++i) {
        uint32_t j = brev(i, lg);
        T[j]       = x;
        Tp[j]      = orc_shoup_precon(x, q);
        Ti[j]      = xi;
        Tip[j]     = orc_shoup_precon(xi, q);
        x          = mulmod(x, root, q);
        xi         = mulmod(xi, rootInv, q);
    }
    *nInv       = orc_invmod(N % q, q);
    *nInvPrecon = orc_shoup_precon(*nInv, q);
}

/* ------------------------------------------------------------------ a6: forward NTT
 * transformnat-impl.h:303-374 ForwardTransformToBitReverseInPlace (Cooley-Tukey, natural in,
 * bit-reversed out, canonical after every stage; the reference peels the last stage, same arithmetic). */
static void ntt_fwd_tables(uint64_t q, uint32_t N, const uint64_t* T, const uint64_t* Tp, uint64_t* x) {
    for (uint32_t m = 1, t = N >> 1; m < N; m <<= 1, t >>= 1) {
        for (uint32_t i = 0; i < m; ++i) {
            uint64_t w = T[m + i], wp = Tp[m + i];
            uint32_t j1 = 2 * i * t, j2 = j1 + t;
            for (uint32_t j = j1; j < j2; ++j) {
                uint64_t v  = orc_mulmod_shoup(x[j + t], w, wp, q);
                uint64_t lo = x[j];
                uint64_t hi = lo + v;
                if (hi >= q)
                    hi -= q;
                if (lo < v)
                    lo += q;
                lo -= v;
                x[j]     = hi;
                x[j + t] = lo;
            }
        }
    }
}
/* ------------------------------------------------------------------ a7: inverse NTT
 * transformnat-impl.h:512-625 InverseTransformFromBitReverseInPlace (Gentleman-Sande; n^-1 folded into
 * the last stage: low half * nInv, high half * (Tinv[1] * nInv), :538-539,603-624).                     */
static void ntt_inv_tables(uint64_t q, uint32_t N, const uint64_t* Ti, const uint64_t* Tip, uint64_t nInv,
                           uint64_t nInvP, uint64_t* x) {
    uint64_t w1  = orc_mulmod_shoup(Ti[1], nInv, nInvP, q);
    uint64_t w1p = orc_shoup_precon(w1, q);
    for (uint32_t m = N >> 1, t = 1; m > 1; m >>= 1, t <<= 1) {
        for (uint32_t i = 0; i < m; ++i) {
            uint64_t w = Ti[m + i], wp = Tip[m + i];
            uint32_t j1 = 2 * i * t, j2 = j1 + t;
            for (uint32_t j = j1; j < j2; ++j) {
                uint64_t lo = x[j], hi = x[j + t];
                uint64_t d = lo;
                if (d < hi)
                    d += q;
                d -= hi;
                lo += hi;
                if (lo >= q)
                    lo -= q;
                x[j]     = lo;
                x[j + t] = orc_mulmod_shoup(d, w, wp, q);
            }
        }
    }
    uint32_t h = N >> 1;
    for (uint32_t j = 0; j < h; ++j) {
        uint64_t lo = x[j], hi = x[j + h];
        uint64_t d = lo;
        if (d < hi)
            d += q;
        d -= hi;
        lo += hi;
        if (lo >= q)
            lo -= q;
        x[j]     = orc_mulmod_shoup(lo, nInv, nInvP, q);
        x[j + h] = orc_mulmod_shoup(d, w1, w1p, q);
    }
}
void orc_ntt_fwd(uint64_t q, uint64_t root, uint32_t N, uint64_t* x) {
    uint64_t* tb = (uint64_t*)malloc(sizeof(uint64_t) * 4 * N);
    uint64_t ni, nip;
    orc_twiddles(q, root, N, tb, tb + N, tb + 2 * N, tb + 3 * N, &ni, &nip);
    ntt_fwd_tables(q, N, tb, tb + N, x);
    free(tb);
}
void orc_ntt_inv(uint64_t q, uint64_t root, uint32_t N, uint64_t* x) {
    uint64_t* tb = (uint64_t*)malloc(sizeof(uint64_t) * 4 * N);
    uint64_t ni, nip;
    orc_twiddles(q, root, N, tb, tb + N, tb + 2 * N, tb + 3 * N, &ni, &nip);
    ntt_inv_tables(q, N, tb + 2 * N, tb + 3 * N, ni, nip, x);
    free(tb);
}

/* ------------------------------------------------------------------ a4 / a9: limb-wise ops
 * mubintvecnat.h:417-423 ModAddNoCheckEq, :492-503 ModMulNoCheckEq (Barrett), mubintvecnat.cpp:273 ModSubEq,
 * dcrtpoly-impl.h:347-621 apply them limb by limb.                                                       */
void orc_limb_binop(uint64_t q, uint32_t N, int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    uint64_t mu = orc_barrett_mu(q);
    for (uint32_t i = 0; i < N; ++i) {
        uint64_t r;
        switch (op) {
            case 0:
                r = a[i] + b[i];
                if (r >= q)
                    r -= q;
                break;
            case 1:
                r = a[i] >= b[i] ? a[i] - b[i] : a[i] + q - b[i];
                break;
            case 2:
                r = orc_mulmod_barrett(a[i], b[i], q, mu);
                break;
            default:
                r = a[i] ? q - a[i] : 0;
                break;
        }
        out[i] = r;
    }
}
/* mubintvecnat.cpp:295-304 ModMulEq(scalar): Shoup with the scalar's precon */
void orc_limb_mul_scalar(uint64_t q, uint32_t N, const uint64_t* a, uint64_t s, uint64_t* out) {
    uint64_t sp = orc_shoup_precon(s, q);
    for (uint32_t i = 0; i < N; ++i)
        out[i] = orc_mulmod_shoup(a[i], s, sp, q);
}
/* mubintvecnat.cpp:109-122 SwitchModulus (centred lift); the shrinking branch reduces both operands
 * of the subtraction modulo the new modulus first (ubintnat.h:891-903 ModSubEq).                        */
void orc_switch_modulus(uint32_t N, uint64_t qf, uint64_t qt, uint64_t* x) {
    uint64_t half = qf >> 1, diff = qf > qt ? qf - qt : qt - qf;
    for (uint32_t i = 0; i < N; ++i) {
        uint64_t v = x[i];
        if (qt > qf) {
            if (v > half)
                v += diff;
        }
        else {
            uint64_t a = v % qt, s = (v > half ? diff : 0) % qt;
            v = a >= s ? a - s : a + qt - s;
        }
        x[i] = v;
    }
}
/* ------------------------------------------------------------------ a15: automorphism
 * nbtheory2.cpp:264-275 PrecomputeAutoMap; poly-impl.h:366-376 out[j] = in[map[j]] (EVALUATION).        */
void orc_automorphism_map(uint32_t N, uint32_t k, uint32_t* map) {
    uint32_t m = N << 1, logm = ilog2(m), logn = ilog2(N);
    for (uint32_t j = 0; j < N; ++j) {
        uint32_t jTmp   = (j << 1) + 1;
        uint32_t idx    = (uint32_t)(((uint64_t)jTmp * k) - ((((uint64_t)jTmp * k) >> logm) << logm)) >> 1;
        uint32_t jrev   = brev(j, logn);
        uint32_t idxrev = brev(idx, logn);
        map[jrev]       = idxrev;
    }
}
void orc_automorphism_eval(uint32_t N, const uint32_t* map, const uint64_t* in, uint64_t* out) {
    for (uint32_t j = 0; j < N; ++j)
        out[j] = in[map[j]];
}
/* COEFFICIENT format, poly-impl.h:359-362: out[jk mod N] = ((jk div N) odd) ? q - in[j] : in[j]; the subtraction is a
 * plain integer one, so a zero coefficient with an odd wrap count comes out as q (not reduced), as in the reference. */
void orc_automorphism_coeff(uint32_t N, uint32_t k, uint64_t q, const uint64_t* in, uint64_t* out) {
    uint32_t logn = ilog2(N), mask = N - 1;
    for (uint32_t j = 0, jk = 0; j < N; ++j, jk += k)
        out[jk & mask] = ((jk >> logn) & 1u) ? q - in[j] : in[j];
}

/* ------------------------------------------------------------------ a10: ApproxSwitchCRTBasis
 * dcrtpoly-impl.h:888-918 (int128 branch): y_i = x_i * QHatInv_i mod q_i (Shoup);
 * out_j = Barrett128( sum_i y_i * QHatModp[i][j] ) mod p_j.                                             */
void orc_approx_switch_crt_basis(uint32_t N, uint32_t sizeQ, const uint64_t* q, const uint64_t* QHatInvModq,
                                 uint32_t sizeP, const uint64_t* p, const uint64_t* QHatModp, const uint64_t* in,
                                 uint64_t* out) {
    uint64_t* pre = (uint64_t*)malloc(sizeof(uint64_t) * (sizeQ + 2 * sizeP));
    uint64_t *muh = pre + sizeQ, *mul = muh + sizeP;
    for (uint32_t i = 0; i < sizeQ; ++i)
        pre[i] = orc_shoup_precon(QHatInvModq[i], q[i]);
    for (uint32_t j = 0; j < sizeP; ++j)
        orc_barrett128_mu(p[j], &muh[j], &mul[j]);
#pragma omp parallel
    {
        u128* sum = (u128*)malloc(sizeof(u128) * sizeP);
#pragma omp for
        for (uint32_t r = 0; r < N; ++r) {
            for (uint32_t j = 0; j < sizeP; ++j)
                sum[j] = 0;
            for (uint32_t i = 0; i < sizeQ; ++i) {
                uint64_t y = orc_mulmod_shoup(in[(size_t)i * N + r], QHatInvModq[i], pre[i], q[i]);
                for (uint32_t j = 0; j < sizeP; ++j)
                    sum[j] += (u128)y * QHatModp[(size_t)i * sizeP + j];
            }
            for (uint32_t j = 0; j < sizeP; ++j)
                out[(size_t)j * N + r] =
                    orc_barrett128((uint64_t)(sum[j] >> 64), (uint64_t)sum[j], p[j], muh[j], mul[j]);
        }
        free(sum);
    }
    free(pre);
}

/* ------------------------------------------------------------------ a20: CKKS moduli chain + HYBRID tables */
struct orc_ckks {
    uint32_t N, sizeQ, sizeP, dnum, alpha;
    uint64_t *q, *rq; /* sizeQ + sizeP moduli / roots (Q then P) */
    uint64_t* tw;     /* per modulus: T,Tp,Ti,Tip (4N) */
    uint64_t *nInv, *nInvP;
};
static const uint64_t* T_(const orc_ckks* c, uint32_t mi) { return c->tw + (size_t)mi * 4 * c->N; }

/* product of a set of moduli, reduced mod m (replaces the reference's BigInteger products) */
static uint64_t prod_mod(const uint64_t* v, uint32_t n, int skip, uint64_t m) {
    uint64_t r = 1 % m;
    for (uint32_t i = 0; i < n; ++i)
        if ((int)i != skip)
            r = mulmod(r, v[i] % m, m);
    return r;
}

/* ckksrns-parametergeneration.cpp:409-518 SinglePrimeModuliGen */
static void ckks_moduli_q(uint32_t N, uint32_t numPrimes, uint32_t scaleBits, uint32_t firstBits, int scalTech,
                          uint64_t* mq) {
    /* scalTech 3 = FLEXIBLEAUTOEXT: the FLEXIBLEAUTO chain, then one extra modulus written to mq[numPrimes] */
    const int ext = scalTech == 3;
    if (ext)
        scalTech = 2;
    uint64_t m = 2 * (uint64_t)N;
    uint64_t q = orc_first_prime(scaleBits, m);
    mq[numPrimes - 1] = q;
    uint64_t maxPrime = q, minPrime = q;
    if (numPrimes > 1) {
        if (scalTech != 2) {
            uint64_t qPrev = q, qNext = q;
            uint32_t cnt = 0;
            for (uint32_t i = numPrimes - 2; i >= 1; --i, ++cnt) {
                if ((cnt % 2) == 0) {
                    qPrev = orc_previous_prime(qPrev, m);
                    mq[i] = qPrev;
                }
                else {
                    qNext = orc_next_prime(qNext, m);
                    mq[i] = qNext;
                }
                if (mq[i] > maxPrime)
                    maxPrime = mq[i];
                else if (mq[i] < minPrime)
                    minPrime = mq[i];
            }
        }
        else { /* FLEXIBLEAUTO :443-487 */
            double sf    = (double)mq[numPrimes - 1];
            uint32_t cnt = 0;
            for (uint32_t i = numPrimes - 2; i >= 1; --i, ++cnt) {
                sf             = pow(sf, 2) / (double)mq[i + 1];
                uint64_t sfInt = (uint64_t)llround(sf);
                uint64_t sfRem = sfInt % m;
                int same       = 1;
                if ((cnt % 2) == 0) {
                    uint64_t qPrev = sfInt - m - sfRem + 1;
                    while (same) {
                        same  = 0;
                        qPrev = orc_previous_prime(qPrev, m);
                        for (uint32_t j = i + 1; j < numPrimes; ++j)
                            if (qPrev == mq[j]) {
                                same = 1;
                                break;
                            }
                    }
                    mq[i] = qPrev;
                }
                else {
                    uint64_t qNext = sfInt + m - sfRem + 1;
                    while (same) {
                        same  = 0;
                        qNext = orc_next_prime(qNext, m);
                        for (uint32_t j = i + 1; j < numPrimes; ++j)
                            if (qNext == mq[j]) {
                                same = 1;
                                break;
                            }
                    }
                    mq[i] = qNext;
                }
                if (mq[i] > maxPrime)
                    maxPrime = mq[i];
                else if (mq[i] < minPrime)
                    minPrime = mq[i];
            }
        }
    }
    if (firstBits == scaleBits)
        mq[0] = orc_next_prime(maxPrime, m);
    else {
        mq[0] = orc_last_prime(firstBits, m);
        for (uint32_t i = 1; i < numPrimes; ++i)
            if (mq[i] == mq[0]) {
                mq[0] = orc_next_prime(maxPrime, m);
                break;
            }
    }
    if (ext) { /* :518-529, DEFAULT_EXTRA_MOD_SIZE = 20 (rns-modulus-limits.h:42) */
        if (mq[0] > maxPrime)
            maxPrime = mq[0];
        uint64_t t = orc_first_prime(20 - 1, m);
        for (uint32_t i = 0; i < numPrimes; ++i)
            if (mq[i] == t) {
                t = orc_next_prime(maxPrime, m);
                break;
            }
        mq[numPrimes] = t;
    }
}

static uint32_t bitlen_prod(const uint64_t* v, uint32_t n) { /* exact bit length of a product of u64 */
    /* schoolbook big-number product, little-endian u64 words */
    uint64_t w[80];
    uint32_t len = 1;
    memset(w, 0, sizeof(w));
    w[0] = 1;
    for (uint32_t i = 0; i < n; ++i) {
        u128 carry = 0;
        for (uint32_t k = 0; k < len; ++k) {
            u128 t = (u128)w[k] * v[i] + carry;
            w[k]   = (uint64_t)t;
            carry  = t >> 64;
        }
        if (carry)
            w[len++] = (uint64_t)carry;
    }
    return (len - 1) * 64 + msb64(w[len - 1]);
}

orc_ckks* orc_ckks_create(uint32_t N, uint32_t depth, uint32_t scaleBits, uint32_t firstBits, uint32_t dnum,
                          int scalTech) {
    orc_ckks* c = (orc_ckks*)calloc(1, sizeof(orc_ckks));
    uint32_t sizeQ = depth + 1;
    /* scheme-utils.h:35-43 ComputeNumLargeDigits */
    if (dnum == 0)
        dnum = depth > 3 ? 3 : (depth > 0 ? 2 : 1);
    uint64_t mq[256];
    ckks_moduli_q(N, sizeQ, scaleBits, firstBits, scalTech, mq);
    if (scalTech == 3) /* ckksrns-parametergeneration.cpp:169: the vector holds numPrimes + 1 moduli */
        sizeQ += 1;
    /* rns-cryptoparameters.cpp:86-97 numPerPartQ; :130-148 sizeP from the widest digit, auxBits = 60 */
    uint32_t alpha = (sizeQ + dnum - 1) / dnum;
    uint32_t maxBits = 0;
    for (uint32_t j = 0; j < dnum; ++j) {
        uint32_t s = j * alpha, e = (j + 1) * alpha < sizeQ ? (j + 1) * alpha : sizeQ;
        uint32_t b = bitlen_prod(mq + s, e - s);
        if (b > maxBits)
            maxBits = b;
    }
    uint32_t sizeP = (maxBits + 59) / 60;
    /* :151-180 P primes: PreviousPrime from FirstPrime(60, 2N) (ckksrns-cryptoparameters.cpp:168-171 step 2N),
     * skipping any value already in Q */
    uint64_t m = 2 * (uint64_t)N, pPrev = orc_first_prime(60, m);
    for (uint32_t i = 0; i < sizeP; ++i) {
        int found;
        do {
            mq[sizeQ + i] = orc_previous_prime(pPrev, m);
            found         = 0;
            for (uint32_t j = 0; j < sizeQ; ++j)
                if (mq[sizeQ + i] == mq[j])
                    found = 1;
            pPrev = mq[sizeQ + i];
        } while (found);
    }
    c->N = N, c->sizeQ = sizeQ, c->sizeP = sizeP, c->dnum = dnum, c->alpha = alpha;
    uint32_t tot = sizeQ + sizeP;
    c->q     = (uint64_t*)malloc(sizeof(uint64_t) * tot);
    c->rq    = (uint64_t*)malloc(sizeof(uint64_t) * tot);
    c->nInv  = (uint64_t*)malloc(sizeof(uint64_t) * tot);
    c->nInvP = (uint64_t*)malloc(sizeof(uint64_t) * tot);
    c->tw    = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)tot * 4 * N);
    memcpy(c->q, mq, sizeof(uint64_t) * tot);
#pragma omp parallel for schedule(dynamic)
    for (uint32_t i = 0; i < tot; ++i) {
        c->rq[i]    = orc_root_of_unity(m, c->q[i]);
        uint64_t* t = c->tw + (size_t)i * 4 * N;
        orc_twiddles(c->q[i], c->rq[i], N, t, t + N, t + 2 * N, t + 3 * N, &c->nInv[i], &c->nInvP[i]);
    }
    return c;
}
/* the same hybrid key-switch context over a caller-supplied chain (BFV's Q and P, rns-cryptoparameters.cpp:48-401
 * is scheme independent): moduli / roots = Q (sizeQ) then P (sizeP); numPerPartQ = ceil(sizeQ / dnum) (:86-97) */
orc_ckks* orc_hks_create_from_moduli(uint32_t N, uint32_t sizeQ, uint32_t sizeP, uint32_t dnum, const uint64_t* moduli,
                                     const uint64_t* roots) {
    orc_ckks* c = (orc_ckks*)calloc(1, sizeof(orc_ckks));
    uint32_t tot = sizeQ + sizeP;
    c->N = N, c->sizeQ = sizeQ, c->sizeP = sizeP, c->dnum = dnum, c->alpha = (sizeQ + dnum - 1) / dnum;
    c->q     = (uint64_t*)malloc(sizeof(uint64_t) * tot);
    c->rq    = (uint64_t*)malloc(sizeof(uint64_t) * tot);
    c->nInv  = (uint64_t*)malloc(sizeof(uint64_t) * tot);
    c->nInvP = (uint64_t*)malloc(sizeof(uint64_t) * tot);
    c->tw    = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)tot * 4 * N);
    memcpy(c->q, moduli, sizeof(uint64_t) * tot);
    memcpy(c->rq, roots, sizeof(uint64_t) * tot);
#pragma omp parallel for schedule(dynamic)
    for (uint32_t i = 0; i < tot; ++i) {
        uint64_t* t = c->tw + (size_t)i * 4 * N;
        orc_twiddles(c->q[i], c->rq[i], N, t, t + N, t + 2 * N, t + 3 * N, &c->nInv[i], &c->nInvP[i]);
    }
    return c;
}
void orc_ckks_destroy(orc_ckks* c) {
    if (!c)
        return;
    free(c->q), free(c->rq), free(c->nInv), free(c->nInvP), free(c->tw), free(c);
}
void orc_ckks_shape(const orc_ckks* c, uint32_t* s) {
    s[0] = c->N, s[1] = c->sizeQ, s[2] = c->sizeP, s[3] = c->dnum, s[4] = c->alpha;
}
void orc_ckks_moduli(const orc_ckks* c, uint64_t* moduli, uint64_t* roots) {
    memcpy(moduli, c->q, sizeof(uint64_t) * (c->sizeQ + c->sizeP));
    memcpy(roots, c->rq, sizeof(uint64_t) * (c->sizeQ + c->sizeP));
}
/* rns-cryptoparameters.cpp:199-229 */
void orc_ckks_PModq(const orc_ckks* c, uint64_t* out) {
    for (uint32_t i = 0; i < c->sizeQ; ++i)
        out[i] = prod_mod(c->q + c->sizeQ, c->sizeP, -1, c->q[i]);
}
void orc_ckks_PInvModq(const orc_ckks* c, uint64_t* out) {
    for (uint32_t i = 0; i < c->sizeQ; ++i)
        out[i] = orc_invmod(prod_mod(c->q + c->sizeQ, c->sizeP, -1, c->q[i]), c->q[i]);
}
void orc_ckks_PHatInvModp(const orc_ckks* c, uint64_t* out) {
    const uint64_t* p = c->q + c->sizeQ;
    for (uint32_t j = 0; j < c->sizeP; ++j)
        out[j] = orc_invmod(prod_mod(p, c->sizeP, (int)j, p[j]), p[j]);
}
void orc_ckks_PHatModq(const orc_ckks* c, uint64_t* out) {
    const uint64_t* p = c->q + c->sizeQ;
    for (uint32_t j = 0; j < c->sizeP; ++j)
        for (uint32_t i = 0; i < c->sizeQ; ++i)
            out[(size_t)j * c->sizeQ + i] = prod_mod(p, c->sizeP, (int)j, c->q[i]);
}
/* rns-cryptoparameters.cpp:296-319: digit `part` truncated to digit_size limbs */
void orc_ckks_PartQlHatInvModq(const orc_ckks* c, uint32_t part, uint32_t digit_size, uint64_t* out) {
    const uint64_t* d = c->q + part * c->alpha;
    for (uint32_t i = 0; i < digit_size; ++i)
        out[i] = orc_invmod(prod_mod(d, digit_size, (int)i, d[i]), d[i]);
}
/* complementary basis of digit `part` at sizeQl limbs: rns-cryptoparameters.cpp:252-294 */
static uint32_t compl_basis(const orc_ckks* c, uint32_t sizeQl, uint32_t part, uint32_t* idx) {
    uint32_t s = part * c->alpha, e = s + c->alpha < sizeQl ? s + c->alpha : sizeQl, n = 0;
    for (uint32_t i = 0; i < sizeQl; ++i)
        if (i < s || i >= e)
            idx[n++] = i;
    for (uint32_t j = 0; j < c->sizeP; ++j)
        idx[n++] = c->sizeQ + j;
    return n;
}
/* rns-cryptoparameters.cpp:321-349 */
uint32_t orc_ckks_PartQlHatModp(const orc_ckks* c, uint32_t sizeQl, uint32_t part, uint64_t* out,
                                uint64_t* compl_moduli) {
    uint32_t idx[256];
    uint32_t sc = compl_basis(c, sizeQl, part, idx);
    uint32_t s = part * c->alpha, e = s + c->alpha < sizeQl ? s + c->alpha : sizeQl, ds = e - s;
    if (compl_moduli)
        for (uint32_t j = 0; j < sc; ++j)
            compl_moduli[j] = c->q[idx[j]];
    if (out)
        for (uint32_t i = 0; i < ds; ++i)
            for (uint32_t j = 0; j < sc; ++j)
                out[(size_t)i * sc + j] = prod_mod(c->q + s, ds, (int)i, c->q[idx[j]]);
    return sc;
}
/* ckksrns-cryptoparameters.cpp:60-83: Ql = q_0..q_{l-1}, l = sizeQl-1 the dropped limb:
 * QlQlInvModqlDivqlModq[i] = ((Ql^-1 mod q_l) * Ql) / q_l  mod q_i ;  qlInvModq[i] = q_l^-1 mod q_i.
 * The division is exact-floor of a BigInteger; mod q_i it equals (QlInv*Ql - r) * q_l^-1 with
 * r = (QlInv*Ql) mod q_l = 1, i.e. (QlInv * (Ql mod q_i) - 1) * q_l^-1 mod q_i.                          */
void orc_ckks_rescale_tables(const orc_ckks* c, uint32_t sizeQl, uint64_t* A, uint64_t* B) {
    uint32_t l      = sizeQl - 1;
    uint64_t ql     = c->q[l];
    uint64_t QlInv  = orc_invmod(prod_mod(c->q, l, -1, ql), ql);
    for (uint32_t i = 0; i < l; ++i) {
        uint64_t qi  = c->q[i];
        uint64_t Qli = prod_mod(c->q, l, -1, qi);
        uint64_t t   = mulmod(QlInv % qi, Qli, qi);
        t            = t >= 1 ? t - 1 : t + qi - 1;
        B[i]         = orc_invmod(ql % qi, qi);
        A[i]         = mulmod(t, B[i], qi);
    }
}

/* helpers on limbs of a context modulus */
static void limb_ntt(const orc_ckks* c, uint32_t mi, uint64_t* x) {
    const uint64_t* t = T_(c, mi);
    ntt_fwd_tables(c->q[mi], c->N, t, t + c->N, x);
}
static void limb_intt(const orc_ckks* c, uint32_t mi, uint64_t* x) {
    const uint64_t* t = T_(c, mi);
    ntt_inv_tables(c->q[mi], c->N, t + 2 * c->N, t + 3 * c->N, c->nInv[mi], c->nInvP[mi], x);
}

/* ------------------------------------------------------------------ a16: EvalKeySwitchPrecomputeCore (ModUp)
 * keyswitch-hybrid.cpp:314-379.  poly [sizeQl][N] EVALUATION -> digits [numPartQl][sizeQl+sizeP][N].     */
int orc_ckks_modup(const orc_ckks* c, uint32_t sizeQl, const uint64_t* poly, uint64_t* digits) {
    uint32_t N = c->N, sizeQlP = sizeQl + c->sizeP;
    uint32_t numPart = (sizeQl + c->alpha - 1) / c->alpha;
    if (numPart > c->dnum)
        numPart = c->dnum;
    for (uint32_t part = 0; part < numPart; ++part) {
        uint32_t s = part * c->alpha, ds = sizeQl - s < c->alpha ? sizeQl - s : c->alpha;
        uint32_t idx[256];
        uint32_t sc = compl_basis(c, sizeQl, part, idx);
        uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(ds + sc) * N);
        uint64_t* cv  = tmp + (size_t)ds * N;
        memcpy(tmp, poly + (size_t)s * N, sizeof(uint64_t) * (size_t)ds * N);
#pragma omp parallel for
        for (uint32_t i = 0; i < ds; ++i)
            limb_intt(c, s + i, tmp + (size_t)i * N); /* :359 SetFormat(COEFFICIENT) */
        uint64_t *hinv = (uint64_t*)malloc(sizeof(uint64_t) * ds), *hmod = (uint64_t*)malloc(sizeof(uint64_t) * ds * sc),
                 *pm = (uint64_t*)malloc(sizeof(uint64_t) * sc);
        orc_ckks_PartQlHatInvModq(c, part, ds, hinv);
        orc_ckks_PartQlHatModp(c, sizeQl, part, hmod, pm);
        orc_approx_switch_crt_basis(N, ds, c->q + s, hinv, sc, pm, hmod, tmp, cv); /* :360-365 */
#pragma omp parallel for
        for (uint32_t j = 0; j < sc; ++j)
            limb_ntt(c, idx[j], cv + (size_t)j * N); /* :366 SetFormat(EVALUATION) */
        uint64_t* D = digits + (size_t)part * sizeQlP * N;
        /* :370-377 assemble: complementary limbs around the original (EVALUATION) digit limbs */
        for (uint32_t i = 0; i < s; ++i)
            memcpy(D + (size_t)i * N, cv + (size_t)i * N, sizeof(uint64_t) * N);
        for (uint32_t i = s; i < s + ds; ++i)
            memcpy(D + (size_t)i * N, poly + (size_t)i * N, sizeof(uint64_t) * N);
        for (uint32_t i = s + ds; i < sizeQlP; ++i)
            memcpy(D + (size_t)i * N, cv + (size_t)(i - ds) * N, sizeof(uint64_t) * N);
        free(tmp), free(hinv), free(hmod), free(pm);
    }
    return (int)numPart;
}

/* ------------------------------------------------------------------ a17: EvalFastKeySwitchCoreExt
 * keyswitch-hybrid.cpp:402-435: e0[i] = sum_j D_j[i]*b_j[idx], e1[i] = sum_j D_j[i]*a_j[idx],
 * idx = i < sizeQl ? i : i + (sizeQ - sizeQl).                                                           */
void orc_ckks_inner_product(const orc_ckks* c, uint32_t sizeQl, uint32_t nd, const uint64_t* digits,
                            const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* e0, uint64_t* e1) {
    uint32_t N = c->N, sizeQlP = sizeQl + c->sizeP, sizeQP = c->sizeQ + c->sizeP, delta = c->sizeQ - sizeQl;
#pragma omp parallel for
    for (uint32_t i = 0; i < sizeQlP; ++i) {
        uint32_t idx = i >= sizeQl ? i + delta : i;
        uint64_t q = c->q[idx], mu = orc_barrett_mu(q);
        uint64_t *o0 = e0 + (size_t)i * N, *o1 = e1 + (size_t)i * N;
        memset(o0, 0, sizeof(uint64_t) * N);
        memset(o1, 0, sizeof(uint64_t) * N);
        for (uint32_t j = 0; j < nd; ++j) {
            const uint64_t* d = digits + ((size_t)j * sizeQlP + i) * N;
            const uint64_t* a = evk_a + ((size_t)j * sizeQP + idx) * N;
            const uint64_t* b = evk_b + ((size_t)j * sizeQP + idx) * N;
            for (uint32_t r = 0; r < N; ++r) {
                uint64_t t0 = o0[r] + orc_mulmod_barrett(d[r], b[r], q, mu);
                uint64_t t1 = o1[r] + orc_mulmod_barrett(d[r], a[r], q, mu);
                o0[r]       = t0 >= q ? t0 - q : t0;
                o1[r]       = t1 >= q ? t1 - q : t1;
            }
        }
    }
}

/* ------------------------------------------------------------------ a11: ApproxModDown (CKKS: t = 0)
 * dcrtpoly-impl.h:966-1005: P limbs -> INTT -> a10 (P -> Ql) -> NTT -> (x_i - conv_i) * PInvModq_i.      */
void orc_ckks_approx_mod_down(const orc_ckks* c, uint32_t sizeQl, const uint64_t* in, uint64_t* out) {
    uint32_t N = c->N, sizeP = c->sizeP;
    uint64_t* partP = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(sizeP + sizeQl) * N);
    uint64_t* conv  = partP + (size_t)sizeP * N;
    memcpy(partP, in + (size_t)sizeQl * N, sizeof(uint64_t) * (size_t)sizeP * N);
#pragma omp parallel for
    for (uint32_t j = 0; j < sizeP; ++j)
        limb_intt(c, c->sizeQ + j, partP + (size_t)j * N);
    uint64_t* PHatInv = (uint64_t*)malloc(sizeof(uint64_t) * sizeP);
    uint64_t* PHat    = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)sizeP * c->sizeQ);
    uint64_t* PHatl   = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)sizeP * sizeQl);
    uint64_t* PInv    = (uint64_t*)malloc(sizeof(uint64_t) * c->sizeQ);
    orc_ckks_PHatInvModp(c, PHatInv);
    orc_ckks_PHatModq(c, PHat);
    orc_ckks_PInvModq(c, PInv);
    for (uint32_t j = 0; j < sizeP; ++j)
        memcpy(PHatl + (size_t)j * sizeQl, PHat + (size_t)j * c->sizeQ, sizeof(uint64_t) * sizeQl);
    orc_approx_switch_crt_basis(N, sizeP, c->q + c->sizeQ, PHatInv, sizeQl, c->q, PHatl, partP, conv);
#pragma omp parallel for
    for (uint32_t i = 0; i < sizeQl; ++i) {
        uint64_t q = c->q[i], w = PInv[i], wp = orc_shoup_precon(w, q);
        uint64_t* cv = conv + (size_t)i * N;
        limb_ntt(c, i, cv);
        for (uint32_t r = 0; r < N; ++r) {
            uint64_t x = in[(size_t)i * N + r], y = cv[r];
            uint64_t d = x >= y ? x - y : x + q - y;
            out[(size_t)i * N + r] = orc_mulmod_shoup(d, w, wp, q); /* Times(scalar) -> ModMulEq, Shoup */
        }
    }
    free(partP), free(PHatInv), free(PHat), free(PHatl), free(PInv);
}

/* ------------------------------------------------------------------ a18: KeySwitchCore
 * keyswitch-hybrid.cpp:308-312,381-400 */
void orc_ckks_keyswitch_core(const orc_ckks* c, uint32_t sizeQl, const uint64_t* poly, const uint64_t* evk_a,
                             const uint64_t* evk_b, uint64_t* out0, uint64_t* out1) {
    uint32_t N = c->N, sizeQlP = sizeQl + c->sizeP;
    uint64_t* digits = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)c->dnum * sizeQlP * N);
    uint64_t* e0     = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)2 * sizeQlP * N);
    uint64_t* e1     = e0 + (size_t)sizeQlP * N;
    int nd           = orc_ckks_modup(c, sizeQl, poly, digits);
    orc_ckks_inner_product(c, sizeQl, (uint32_t)nd, digits, evk_a, evk_b, e0, e1);
    orc_ckks_approx_mod_down(c, sizeQl, e0, out0);
    orc_ckks_approx_mod_down(c, sizeQl, e1, out1);
    free(digits), free(e0);
}

/* ------------------------------------------------------------------ a19: EvalMult = EvalMultCore + relinearise
 * base-leveledshe.cpp:608-646 (tensor) and :202-214 (KeySwitchCore on c2; c0 += ab0; c1 += ab1).          */
void orc_ckks_evalmult(const orc_ckks* c, uint32_t sizeQl, const uint64_t* ct1, const uint64_t* ct2,
                       const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* out) {
    uint32_t N = c->N;
    size_t P   = (size_t)sizeQl * N;
    uint64_t* c2  = (uint64_t*)malloc(sizeof(uint64_t) * 4 * P);
    uint64_t *t = c2 + P, *ab0 = c2 + 2 * P, *ab1 = c2 + 3 * P;
    const uint64_t *a0 = ct1, *a1 = ct1 + P, *b0 = ct2, *b1 = ct2 + P;
    for (uint32_t i = 0; i < sizeQl; ++i) {
        size_t o = (size_t)i * N;
        orc_limb_binop(c->q[i], N, 2, a0 + o, b0 + o, out + o);     /* c0 = a0*b0 */
        orc_limb_binop(c->q[i], N, 2, a0 + o, b1 + o, out + P + o); /* c1 = a0*b1 + a1*b0 */
        orc_limb_binop(c->q[i], N, 2, a1 + o, b0 + o, t + o);
        orc_limb_binop(c->q[i], N, 0, out + P + o, t + o, out + P + o);
        orc_limb_binop(c->q[i], N, 2, a1 + o, b1 + o, c2 + o); /* c2 = a1*b1 */
    }
    orc_ckks_keyswitch_core(c, sizeQl, c2, evk_a, evk_b, ab0, ab1);
    for (uint32_t i = 0; i < sizeQl; ++i) {
        size_t o = (size_t)i * N;
        orc_limb_binop(c->q[i], N, 0, out + o, ab0 + o, out + o);
        orc_limb_binop(c->q[i], N, 0, out + P + o, ab1 + o, out + P + o);
    }
    free(c2);
}

/* ------------------------------------------------------------------ a14: DropLastElementAndScale (CKKS rescale)
 * dcrtpoly-impl.h:693-712: z = INTT(x_l); per i<l: t = NTT(SwitchModulus(z) * A_i); x_i = x_i*B_i + t.     */
void orc_ckks_rescale_poly(const orc_ckks* c, uint32_t sizeQl, const uint64_t* x, uint64_t* out) {
    uint32_t N = c->N, l = sizeQl - 1;
    uint64_t* A = (uint64_t*)malloc(sizeof(uint64_t) * 2 * l);
    uint64_t* B = A + l;
    orc_ckks_rescale_tables(c, sizeQl, A, B);
    uint64_t* z = (uint64_t*)malloc(sizeof(uint64_t) * N);
    memcpy(z, x + (size_t)l * N, sizeof(uint64_t) * N);
    limb_intt(c, l, z);
#pragma omp parallel for
    for (uint32_t i = 0; i < l; ++i) {
        uint64_t q   = c->q[i];
        uint64_t* t  = (uint64_t*)malloc(sizeof(uint64_t) * N);
        uint64_t* o  = out + (size_t)i * N;
        memcpy(t, z, sizeof(uint64_t) * N);
        orc_switch_modulus(N, c->q[l], q, t);
        orc_limb_mul_scalar(q, N, t, A[i], t);
        limb_ntt(c, i, t);
        orc_limb_mul_scalar(q, N, x + (size_t)i * N, B[i], o);
        orc_limb_binop(q, N, 0, o, t, o);
        free(t);
    }
    free(A), free(z);
}

/* ------------------------------------------------------------------ EvalAutomorphism
 * base-leveledshe.cpp:382-422 + keyswitch-hybrid.cpp:199-214: (c0 + ab0, ab1) with ab = KeySwitchCore(c1),
 * then both polys through the automorphism gather.                                                        */
void orc_ckks_automorphism_ct(const orc_ckks* c, uint32_t sizeQl, uint32_t k, const uint64_t* ct,
                              const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* out) {
    uint32_t N = c->N;
    size_t P   = (size_t)sizeQl * N;
    uint64_t* ab  = (uint64_t*)malloc(sizeof(uint64_t) * 2 * P);
    uint32_t* map = (uint32_t*)malloc(sizeof(uint32_t) * N);
    orc_ckks_keyswitch_core(c, sizeQl, ct + P, evk_a, evk_b, ab, ab + P);
    orc_automorphism_map(N, k, map);
    for (uint32_t i = 0; i < sizeQl; ++i) {
        size_t o = (size_t)i * N;
        orc_limb_binop(c->q[i], N, 0, ct + o, ab + o, ab + o);
        orc_automorphism_eval(N, map, ab + o, out + o);
        orc_automorphism_eval(N, map, ab + P + o, out + P + o);
    }
    free(ab), free(map);
}


/* ------------------------------------------------------------------ hoisted rotations
 * EvalFastRotation (base-leveledshe.cpp:433-463): with digits = EvalKeySwitchPrecomputeCore(c1) computed ONCE for
 * several rotations, ba = EvalFastKeySwitchCore(digits, key_k) (inner product + ModDown), out = (auto_k(ba0 + c0),
 * auto_k(ba1)).  digits: [numPartQl][sizeQl+sizeP][N] from orc_ckks_modup.                                       */
void orc_ckks_fast_rotation(const orc_ckks* c, uint32_t sizeQl, uint32_t k, const uint64_t* ct, const uint64_t* digits,
                            const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* out) {
    uint32_t N = c->N, sizeQlP = sizeQl + c->sizeP;
    uint32_t numPart = (sizeQl + c->alpha - 1) / c->alpha;
    if (numPart > c->dnum)
        numPart = c->dnum;
    size_t P = (size_t)sizeQl * N, E = (size_t)sizeQlP * N;
    uint64_t* e   = (uint64_t*)malloc(sizeof(uint64_t) * 2 * E);
    uint64_t* ba  = (uint64_t*)malloc(sizeof(uint64_t) * 2 * P);
    uint32_t* map = (uint32_t*)malloc(sizeof(uint32_t) * N);
    orc_ckks_inner_product(c, sizeQl, numPart, digits, evk_a, evk_b, e, e + E);
    orc_ckks_approx_mod_down(c, sizeQl, e, ba);
    orc_ckks_approx_mod_down(c, sizeQl, e + E, ba + P);
    orc_automorphism_map(N, k, map);
    for (uint32_t i = 0; i < sizeQl; ++i) {
        size_t o = (size_t)i * N;
        orc_limb_binop(c->q[i], N, 0, ba + o, ct + o, ba + o);
        orc_automorphism_eval(N, map, ba + o, out + o);
        orc_automorphism_eval(N, map, ba + P + o, out + P + o);
    }
    free(e), free(ba), free(map);
}
/* EvalFastRotationExt (ckksrns-leveledshe.cpp:534-582): the result STAYS in the extended basis QlP:
 * cTilda = EvalFastKeySwitchCoreExt(digits, key_k); if addFirst, cTilda0 += [c0 * (P mod q_i) | 0]; both polys through
 * auto_k on all sizeQl + sizeP limbs.  out: [2][sizeQl+sizeP][N].                                                */
void orc_ckks_fast_rotation_ext(const orc_ckks* c, uint32_t sizeQl, uint32_t k, int add_first, const uint64_t* ct,
                                const uint64_t* digits, const uint64_t* evk_a, const uint64_t* evk_b, uint64_t* out) {
    uint32_t N = c->N, sizeQlP = sizeQl + c->sizeP;
    uint32_t numPart = (sizeQl + c->alpha - 1) / c->alpha;
    if (numPart > c->dnum)
        numPart = c->dnum;
    size_t E = (size_t)sizeQlP * N;
    uint64_t* e    = (uint64_t*)malloc(sizeof(uint64_t) * 2 * E);
    uint64_t* tmp  = (uint64_t*)malloc(sizeof(uint64_t) * N);
    uint64_t* pmod = (uint64_t*)malloc(sizeof(uint64_t) * c->sizeQ);
    uint32_t* map  = (uint32_t*)malloc(sizeof(uint32_t) * N);
    orc_ckks_inner_product(c, sizeQl, numPart, digits, evk_a, evk_b, e, e + E);
    orc_ckks_PModq(c, pmod);
    if (add_first)
        for (uint32_t i = 0; i < sizeQl; ++i) {
            size_t o = (size_t)i * N;
            orc_limb_mul_scalar(c->q[i], N, ct + o, pmod[i], tmp);
            orc_limb_binop(c->q[i], N, 0, e + o, tmp, e + o);
        }
    orc_automorphism_map(N, k, map);
    for (uint32_t b = 0; b < 2; ++b)
        for (uint32_t i = 0; i < sizeQlP; ++i)
            orc_automorphism_eval(N, map, e + b * E + (size_t)i * N, out + b * E + (size_t)i * N);
    free(e), free(tmp), free(pmod), free(map);
}
/* KeySwitchExt (keyswitch-hybrid.cpp:217-243): lift a ciphertext to QlP: [c_k * (P mod q_i) | 0], c_0 only if addFirst */
void orc_ckks_keyswitch_ext(const orc_ckks* c, uint32_t sizeQl, int add_first, const uint64_t* ct, uint64_t* out) {
    uint32_t N = c->N, sizeQlP = sizeQl + c->sizeP;
    size_t E = (size_t)sizeQlP * N;
    uint64_t* pmod = (uint64_t*)malloc(sizeof(uint64_t) * c->sizeQ);
    orc_ckks_PModq(c, pmod);
    memset(out, 0, sizeof(uint64_t) * 2 * E);
    for (uint32_t b = 0; b < 2; ++b)
        if (add_first || b > 0)
            for (uint32_t i = 0; i < sizeQl; ++i)
                orc_limb_mul_scalar(c->q[i], N, ct + ((size_t)b * sizeQl + i) * N, pmod[i], out + b * E + (size_t)i * N);
    free(pmod);
}
/* KeySwitchDown (keyswitch-hybrid.cpp:246-277): ApproxModDown of both polynomials of an extended ciphertext */
void orc_ckks_keyswitch_down(const orc_ckks* c, uint32_t sizeQl, const uint64_t* ct_ext, uint64_t* out) {
    size_t E = (size_t)(sizeQl + c->sizeP) * c->N, P = (size_t)sizeQl * c->N;
    orc_ckks_approx_mod_down(c, sizeQl, ct_ext, out);
    orc_ckks_approx_mod_down(c, sizeQl, ct_ext + E, out + P);
}


/* ------------------------------------------------------------------ a12: exact SwitchCRTBasis (BFV, HPS)
 * dcrtpoly-impl.h:1046-1084: as a10 plus the overflow count  nu = 0.5 + sum_i double(y_i) * qInv[i]  (IEEE double,
 * one multiply then one add per term, in index order, NO fused multiply-add), alpha = (size_t)nu, and
 * out_j = ( Barrett128(sum_i y_i * QHatModp[j][i]) - alphaQModp[alpha][j] ) mod p_j.   NOTE the [j][i] table order. */
void orc_switch_crt_basis(uint32_t N, uint32_t sizeQ, const uint64_t* q, const uint64_t* QHatInvModq, uint32_t sizeP,
                          const uint64_t* p, const uint64_t* QHatModp /*[sizeP][sizeQ]*/,
                          const uint64_t* alphaQModp /*[sizeQ+1][sizeP]*/, const double* qInv, const uint64_t* in,
                          uint64_t* out) {
    uint64_t* pre = (uint64_t*)malloc(sizeof(uint64_t) * (sizeQ + 2 * sizeP));
    uint64_t *muh = pre + sizeQ, *mul = muh + sizeP;
    for (uint32_t i = 0; i < sizeQ; ++i)
        pre[i] = orc_shoup_precon(QHatInvModq[i], q[i]);
    for (uint32_t j = 0; j < sizeP; ++j)
        orc_barrett128_mu(p[j], &muh[j], &mul[j]);
#pragma omp parallel
    {
        uint64_t* y = (uint64_t*)malloc(sizeof(uint64_t) * sizeQ);
#pragma omp for
        for (uint32_t r = 0; r < N; ++r) {
            volatile double nu = 0.5; /* volatile: keep every intermediate rounded to double */
            for (uint32_t i = 0; i < sizeQ; ++i) {
                y[i]              = orc_mulmod_shoup(in[(size_t)i * N + r], QHatInvModq[i], pre[i], q[i]);
                volatile double t = (double)y[i] * qInv[i];
                nu                = nu + t;
            }
            size_t alpha = (size_t)nu;
            for (uint32_t j = 0; j < sizeP; ++j) {
                u128 sum = 0;
                for (uint32_t i = 0; i < sizeQ; ++i)
                    sum += (u128)y[i] * QHatModp[(size_t)j * sizeQ + i];
                uint64_t v = orc_barrett128((uint64_t)(sum >> 64), (uint64_t)sum, p[j], muh[j], mul[j]);
                uint64_t a = alphaQModp[alpha * sizeP + j];
                out[(size_t)j * N + r] = v >= a ? v - a : v + p[j] - a; /* ModSubFast */
            }
        }
        free(y);
    }
    free(pre);
}

/* ------------------------------------------------------------------ a13: ScaleAndRound DCRT -> DCRT (BFV)
 * dcrtpoly-impl.h:1513-1586.  in: [sizeI + sizeO][N]; the sizeO output-basis limbs are first when in_first == 0
 * (output modulus == first input modulus: inputIndex = sizeO), else last (outputIndex = sizeI).
 * nu = 0.5 + sum_i frac[i] * double(x_i);  if nu <= 2^63 - 1024 (isConvertableToNativeInt, utils/utilities.h:100-126)
 * alpha = (uint64)nu reduced mod o_j, else alpha = (uint128)nu reduced by Barrett128;
 * out_j = Barrett128( sum_i x_i * tbl[j][i] + x_{o_j} * tbl[j][sizeI] ) + alpha  (mod o_j).                         */
void orc_scale_and_round(uint32_t N, uint32_t sizeI, uint32_t sizeO, const uint64_t* o, int out_limbs_first,
                         const uint64_t* tbl /*[sizeO][sizeI+1]*/, const double* frac, const uint64_t* in,
                         uint64_t* out) {
    uint32_t inputIndex = out_limbs_first ? sizeO : 0, outputIndex = out_limbs_first ? 0 : sizeI;
    uint64_t* muh = (uint64_t*)malloc(sizeof(uint64_t) * 2 * sizeO);
    uint64_t* mul = muh + sizeO;
    for (uint32_t j = 0; j < sizeO; ++j)
        orc_barrett128_mu(o[j], &muh[j], &mul[j]);
    const double limit = 9223372036854774784.0; /* (double)(2^63 - 2^9 - 1) */
#pragma omp parallel for
    for (uint32_t r = 0; r < N; ++r) {
        volatile double nu = 0.5;
        for (uint32_t i = 0; i < sizeI; ++i) {
            volatile double t = frac[i] * (double)in[(size_t)(i + inputIndex) * N + r];
            nu                = nu + t;
        }
        int small        = fabs(nu) <= limit;
        uint64_t alpha64 = small ? (uint64_t)nu : 0;
        u128 alpha128    = small ? 0 : (u128)nu;
        for (uint32_t j = 0; j < sizeO; ++j) {
            const uint64_t* tj = tbl + (size_t)j * (sizeI + 1);
            u128 sum           = 0;
            for (uint32_t i = 0; i < sizeI; ++i)
                sum += (u128)in[(size_t)(i + inputIndex) * N + r] * tj[i];
            sum += (u128)in[(size_t)(outputIndex + j) * N + r] * tj[sizeI];
            uint64_t v = orc_barrett128((uint64_t)(sum >> 64), (uint64_t)sum, o[j], muh[j], mul[j]);
            uint64_t a = small ? (alpha64 >= o[j] ? alpha64 % o[j] : alpha64)
                               : orc_barrett128((uint64_t)(alpha128 >> 64), (uint64_t)alpha128, o[j], muh[j], mul[j]);
            uint64_t s = v + a;
            out[(size_t)j * N + r] = s >= o[j] ? s - o[j] : s;
        }
    }
    free(muh);
}
