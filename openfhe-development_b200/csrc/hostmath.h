// hostmath.h -- host-side number theory used at setup time only (table construction, parameter chains).
//
// The reference does this with BigInteger at context creation (src/core/include/math/nbtheory-impl.h,
// src/pke/lib/schemerns/rns-cryptoparameters.cpp:48-401).  Every quantity the device needs is a residue
// modulo one 64-bit prime, so here products of moduli are only ever formed modulo the target prime and no
// multiprecision type exists.
#pragma once
#include <cstdint>
#include <vector>

namespace ofhe {
namespace hm {

using u128 = unsigned __int128;

inline uint64_t mul(uint64_t a, uint64_t b, uint64_t q) {
    return (uint64_t)((u128)a * b % q);
}
inline uint64_t pow(uint64_t b, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    b %= q;
    for (; e; e >>= 1) {
        if (e & 1)
            r = mul(r, b, q);
        b = mul(b, b, q);
    }
    return r;
}
inline uint64_t inv(uint64_t a, uint64_t q) {  // q prime
    return pow(a % q, q - 2, q);
}
inline uint64_t shoup_const(uint64_t w, uint64_t q) {  // floor(w * 2^64 / q)   (ubintnat.h:1437-1444)
    return (uint64_t)(((u128)w << 64) / q);
}
inline uint32_t bit_length(uint64_t x) {
    return x ? 64 - (uint32_t)__builtin_clzll(x) : 0;
}
// product of v[0..n) skipping index `skip` (or none if skip < 0), modulo m
inline uint64_t prod_mod(const uint64_t* v, size_t n, long skip, uint64_t m) {
    uint64_t r = 1 % m;
    for (size_t i = 0; i < n; ++i)
        if ((long)i != skip)
            r = mul(r, v[i] % m, m);
    return r;
}

bool is_prime(uint64_t n);
uint64_t first_prime(uint32_t bits, uint64_t m);     // nbtheory-impl.h:329-347
uint64_t last_prime(uint32_t bits, uint64_t m);      // :349-373
uint64_t next_prime(uint64_t q, uint64_t m);         // :375-383
uint64_t previous_prime(uint64_t q, uint64_t m);     // :385-393
uint64_t minimal_root_of_unity(uint64_t m, uint64_t q);  // :183-231 (numerically smallest primitive m-th root)

// CKKS modulus chain + HYBRID auxiliary primes, same selection rules as
// ParameterGenerationCKKSRNS::SinglePrimeModuliGen (pke/lib/scheme/ckksrns/ckksrns-parametergeneration.cpp:409-518)
// and CryptoParametersRNS::PrecomputeCRTTables (pke/lib/schemerns/rns-cryptoparameters.cpp:86-180).
struct CkksChain {
    uint32_t N = 0, sizeQ = 0, sizeP = 0, dnum = 0, alpha = 0;
    std::vector<uint64_t> moduli;  // Q then P
    std::vector<uint64_t> roots;
};
// scal_tech: 0 FIXEDMANUAL, 1 FIXEDAUTO, 2 FLEXIBLEAUTO
CkksChain ckks_chain(uint32_t N, uint32_t depth, uint32_t scale_bits, uint32_t first_bits, uint32_t dnum,
                     int scal_tech);

// BFV (HPSPOVERQ) multiplication tables at the full level, from (Q, R, t) alone -- the quantities
// CryptoParametersBFVRNS::PrecomputeCRTTables forms with BigInteger (pke/lib/scheme/bfvrns/bfvrns-cryptoparameters.cpp:140-600),
// here with word-size modular arithmetic only.  `which` and the layouts follow the getters:
//   0 QlHatInvModq[sizeQ]            1 QlHatModr[sizeR][sizeQ]       2 alphaQlModr[sizeQ+1][sizeR]   3 qInv[sizeQ] (double)
//   4 negRlQlHatInvModq[sizeQ]       5 qInvModr[sizeQ][sizeR]        6 RlHatInvModr[sizeR]           7 RlHatModq[sizeQ][sizeR]
//   8 alphaRlModq[sizeR+1][sizeQ]    9 rInv[sizeR] (double)         10 tQlSlHatInvModsDivsModq[sizeQ][sizeR+1]
//  11 tQlSlHatInvModsDivsFrac[sizeR] (double)
// Returns the number of elements written (uint64_t or double, by table), 0 for an unknown table.
size_t bfv_table(const uint64_t* Q, uint32_t sizeQ, const uint64_t* R, uint32_t sizeR, uint64_t t, int which, void* out);

}  // namespace hm
}  // namespace ofhe
