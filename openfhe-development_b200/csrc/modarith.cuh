// modarith.cuh -- 64-bit modular arithmetic primitives for sm_100a (device side).
//
// Replaces the scalar layer of the reference's math HAL
//   src/core/include/math/hal/intnat/ubintnat.h:1437-1486  (Shoup / Harvey constant multiplication)
//   src/core/include/math/hal/intnat/ubintnat.h:1348-1418  (Barrett multiplication)
//   src/core/include/utils/utilities-int.h:47-100          (128 -> 64 Barrett reduction)
// Every routine returns the same canonical residue in [0,q) as the reference does; the internals are
// chosen for the GPU integer pipe (IMAD.WIDE chains, lazy [0,2q)/[0,4q) ranges) and are not a port.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ofhe {

// Per-modulus constants, one entry per registered modulus, resident in global memory.
struct ModConst {
    uint64_t q;         // the prime, q < 2^60
    uint64_t mu;        // floor(2^(63+k) / q), k = bit length of q: single-word Barrett constant
    uint64_t mu128_hi;  // floor(2^128 / q) high / low words (utilities-int.h mu)
    uint64_t mu128_lo;
    uint64_t nInv;      // N^-1 mod q and its Shoup constant (transformnat-impl.h:743-752)
    uint64_t nInvS;
    uint64_t w1nInv;    // Tinv[1] * N^-1 mod q and its Shoup constant (transformnat-impl.h:538-539)
    uint64_t w1nInvS;
    uint64_t r30;       // 2^30 mod q, 2^60 mod q with Shoup constants, and floor(2^64 / q) (the Shoup constant of 1):
    uint64_t r30s;      // column-wise lazy reduction of the base-conversion accumulators (baseconv.cu)
    uint64_t r60;
    uint64_t r60s;
    uint64_t one_s;
    uint32_t k;         // bit length of q
    uint32_t pad;
};

struct __align__(16) Twiddle {  // one table entry: w and its Shoup constant floor(w * 2^64 / q)
    uint64_t w;
    uint64_t ws;
};

__device__ __forceinline__ uint64_t mulhi(uint64_t a, uint64_t b) {
    return __umul64hi(a, b);
}

// x * w mod q, result in [0, 2q); x may be ANY 64-bit value (Harvey's lazy form of ModMulFastConst).
__device__ __forceinline__ uint64_t shoup_lazy(uint64_t x, uint64_t w, uint64_t ws, uint64_t q) {
    return x * w - mulhi(x, ws) * q;
}
// same with the negated modulus nq = 2^64 - q: one multiply-add chain, no 64-bit negate on the (half-rate) ALU pipe
__device__ __forceinline__ uint64_t shoup_lazy_n(uint64_t x, uint64_t w, uint64_t ws, uint64_t nq) {
    return x * w + mulhi(x, ws) * nq;
}
__device__ __forceinline__ uint64_t lazy_reduce_n(uint64_t x, uint64_t one_s, uint64_t nq) {
    return x + mulhi(x, one_s) * nq;
}
// canonical variant, identical result to NativeIntegerT::ModMulFastConst
__device__ __forceinline__ uint64_t shoup(uint64_t x, uint64_t w, uint64_t ws, uint64_t q) {
    uint64_t r = shoup_lazy(x, w, ws, q);
    return r >= q ? r - q : r;
}
// x mod q into [0, 2q) for ANY 64-bit x: Shoup multiplication by 1 (one_s = floor(2^64 / q))
__device__ __forceinline__ uint64_t lazy_reduce(uint64_t x, uint64_t one_s, uint64_t q) {
    return x - mulhi(x, one_s) * q;
}
__device__ __forceinline__ uint64_t csub(uint64_t x, uint64_t q) {  // x in [0,2q) -> [0,q)
    return x >= q ? x - q : x;
}
__device__ __forceinline__ uint64_t addmod(uint64_t a, uint64_t b, uint64_t q) {
    uint64_t r = a + b;
    return r >= q ? r - q : r;
}
__device__ __forceinline__ uint64_t submod(uint64_t a, uint64_t b, uint64_t q) {
    return a >= b ? a - b : a + q - b;
}

// 128-bit product helper
__device__ __forceinline__ void mul128(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
    lo = a * b;
    hi = mulhi(a, b);
}

// (hi:lo) mod q for hi:lo < q * 2^64 style products a*b with a,b < q (single-word Barrett, error <= 2).
__device__ __forceinline__ uint64_t barrett_reduce_prod(uint64_t hi, uint64_t lo, uint64_t q, uint64_t mu,
                                                        uint32_t k) {
    // x1 = floor(p / 2^(k-1)) < 2^(k+1);  qhat = floor(x1 * mu / 2^64) in {Q-2,Q-1,Q}
    uint64_t x1   = (hi << (65 - k)) | (lo >> (k - 1));
    uint64_t qhat = mulhi(x1, mu);
    uint64_t r    = lo - qhat * q;
    r             = r >= 2 * q ? r - 2 * q : r;
    return r >= q ? r - q : r;
}
// a * b mod q for canonical a, b: same value as NativeIntegerT::ModMulFast (ubintnat.h:1389-1401)
__device__ __forceinline__ uint64_t mulmod(uint64_t a, uint64_t b, const ModConst& m) {
    uint64_t hi, lo;
    mul128(a, b, hi, lo);
    return barrett_reduce_prod(hi, lo, m.q, m.mu, m.k);
}

// (hi:lo) mod q for an arbitrary 128-bit value: same value as BarrettUint128ModUint64
// (utilities-int.h:60-100).  qhat = low word of floor(a * mu128 / 2^128) with the lowest partial product
// dropped (under-estimates by at most one more), so r < 4q before the final conditional subtractions.
__device__ __forceinline__ uint64_t barrett128(uint64_t hi, uint64_t lo, uint64_t q, uint64_t mu_hi,
                                               uint64_t mu_lo) {
    uint64_t m1h, m1l, m2h, m2l;
    mul128(lo, mu_hi, m1h, m1l);
    mul128(hi, mu_lo, m2h, m2l);
    uint64_t s     = m1l + m2l;
    uint64_t carry = s < m1l;
    uint64_t qhat  = hi * mu_hi + m1h + m2h + carry;
    uint64_t r     = lo - qhat * q;  // true remainder + {0,1,2,3} * q  (mod 2^64 arithmetic is exact: r < 4q < 2^62)
    r              = r >= 2 * q ? r - 2 * q : r;
    return r >= q ? r - q : r;
}

}  // namespace ofhe
