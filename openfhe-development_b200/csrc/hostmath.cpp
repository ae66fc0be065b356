// hostmath.cpp -- see hostmath.h.  Setup-time only; nothing here is on the per-operation path.
#include "hostmath.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace ofhe {
namespace hm {

bool is_prime(uint64_t n) {
    // deterministic Miller-Rabin: these witnesses decide every n < 2^64 (the reference's randomised test,
    // nbtheory-impl.h, reaches the same verdict)
    static const uint64_t W[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2)
        return false;
    for (uint64_t w : W) {
        if (n == w)
            return true;
        if (n % w == 0)
            return false;
    }
    uint64_t odd = n - 1;
    int twos     = __builtin_ctzll(odd);
    odd >>= twos;
    for (uint64_t w : W) {
        uint64_t y = pow(w, odd, n);
        if (y == 1 || y == n - 1)
            continue;
        bool witness = true;
        for (int i = 1; i < twos && witness; ++i) {
            y = mul(y, y, n);
            if (y == n - 1)
                witness = false;
        }
        if (witness)
            return false;
    }
    return true;
}

uint64_t first_prime(uint32_t bits, uint64_t m) {
    const uint64_t two_k = uint64_t(1) << bits;
    uint64_t cand        = two_k - (two_k % m) + 1;  // == 1 (mod m), <= 2^bits + 1
    if (two_k % m != 0)
        cand += m;
    while (!is_prime(cand))
        cand += m;
    return cand;
}
uint64_t last_prime(uint32_t bits, uint64_t m) {
    const uint64_t two_k = uint64_t(1) << bits;
    uint64_t cand        = two_k - (two_k % m) + 1;
    if (two_k % m < 2)
        cand -= m;
    while (!is_prime(cand))
        cand -= m;
    if (bit_length(cand) != bits)
        throw std::runtime_error("last_prime: no prime of the requested bit length");
    return cand;
}
uint64_t next_prime(uint64_t q, uint64_t m) {
    do
        q += m;
    while (!is_prime(q));
    return q;
}
uint64_t previous_prime(uint64_t q, uint64_t m) {
    do
        q -= m;
    while (!is_prime(q));
    return q;
}

uint64_t minimal_root_of_unity(uint64_t m, uint64_t q) {
    if ((q - 1) % m != 0)
        throw std::runtime_error("root of unity: (q-1) is not divisible by the cyclotomic order");
    // any element of order exactly m (m a power of two): g^((q-1)/m) with (.)^(m/2) == -1
    uint64_t zeta = 0;
    for (uint64_t g = 2; zeta == 0; ++g) {
        uint64_t c = pow(g, (q - 1) / m, q);
        if (pow(c, m / 2, q) == q - 1)
            zeta = c;
    }
    // the m/2 primitive roots are the odd powers of zeta; keep the smallest
    const uint64_t step = mul(zeta, zeta, q);
    uint64_t cur = zeta, smallest = zeta;
    for (uint64_t i = 1; i < m / 2; ++i) {
        cur      = mul(cur, step, q);
        smallest = std::min(smallest, cur);
    }
    return smallest;
}

namespace {
// bit length of the product of a few 64-bit numbers (little-endian word vector)
uint32_t product_bits(const uint64_t* v, size_t n) {
    std::vector<uint64_t> acc{1};
    for (size_t i = 0; i < n; ++i) {
        u128 carry = 0;
        for (auto& w : acc) {
            u128 t = (u128)w * v[i] + carry;
            w      = (uint64_t)t;
            carry  = t >> 64;
        }
        if (carry)
            acc.push_back((uint64_t)carry);
    }
    return (uint32_t)(acc.size() - 1) * 64 + bit_length(acc.back());
}
}  // namespace

CkksChain ckks_chain(uint32_t N, uint32_t depth, uint32_t scale_bits, uint32_t first_bits, uint32_t dnum,
                     int scal_tech) {
    if (scal_tech < 0 || scal_tech > 3)
        throw std::runtime_error(
            "ckks_chain: scaling technique must be FIXEDMANUAL(0), FIXEDAUTO(1), FLEXIBLEAUTO(2) or FLEXIBLEAUTOEXT(3)");
    const bool ext = scal_tech == 3;  // FLEXIBLEAUTOEXT = the FLEXIBLEAUTO chain plus one extra 20-bit-class modulus
    if (ext)
        scal_tech = 2;
    CkksChain c;
    c.N     = N;
    c.sizeQ = depth + 1;
    if (dnum == 0)  // scheme-utils.h:35-43
        dnum = depth > 3 ? 3 : (depth > 0 ? 2 : 1);
    const uint64_t m = 2 * (uint64_t)N;
    uint32_t L = c.sizeQ;
    std::vector<uint64_t> Q(L);
    Q[L - 1]    = first_prime(scale_bits, m);
    uint64_t hi = Q[L - 1];
    if (L > 1) {
        if (scal_tech != 2) {
            // alternate below / above the scaling prime so the product stays near 2^(scale_bits * k)
            uint64_t below = Q[L - 1], above = Q[L - 1];
            for (uint32_t i = L - 2, n = 0; i >= 1; --i, ++n) {
                Q[i] = (n % 2 == 0) ? (below = previous_prime(below, m)) : (above = next_prime(above, m));
                hi   = std::max(hi, Q[i]);
            }
        }
        else {
            // FLEXIBLEAUTO: track the running scaling factor sf_{k+1} = sf_k^2 / q and pick the nearest prime
            // alternately below / above it (ckksrns-parametergeneration.cpp:443-487)
            double sf = (double)Q[L - 1];
            for (uint32_t i = L - 2, n = 0; i >= 1; --i, ++n) {
                sf             = std::pow(sf, 2) / (double)Q[i + 1];
                uint64_t sfInt = (uint64_t)std::llround(sf);
                uint64_t anchor = sfInt - (sfInt % m) + 1;
                uint64_t cand   = (n % 2 == 0) ? anchor - m : anchor + m;
                for (;;) {
                    cand = (n % 2 == 0) ? previous_prime(cand, m) : next_prime(cand, m);
                    if (std::find(Q.begin() + i + 1, Q.end(), cand) == Q.end())
                        break;
                }
                Q[i] = cand;
                hi   = std::max(hi, cand);
            }
        }
    }
    if (first_bits == scale_bits)
        Q[0] = next_prime(hi, m);
    else {
        Q[0] = last_prime(first_bits, m);
        if (std::find(Q.begin() + 1, Q.end(), Q[0]) != Q.end())
            Q[0] = next_prime(hi, m);
    }
    if (ext) {  // ckksrns-parametergeneration.cpp:518-529, DEFAULT_EXTRA_MOD_SIZE = 20 (rns-modulus-limits.h:42)
        hi            = std::max(hi, Q[0]);
        uint64_t cand = first_prime(20 - 1, m);
        if (std::find(Q.begin(), Q.end(), cand) != Q.end())
            cand = next_prime(hi, m);
        Q.push_back(cand);
        c.sizeQ += 1;
    }
    // digits and auxiliary basis
    L       = c.sizeQ;
    c.dnum  = dnum;
    c.alpha = (L + dnum - 1) / dnum;
    if (L <= c.alpha * (dnum - 1))
        throw std::runtime_error("ckks_chain: cannot distribute the towers into the requested number of digits");
    uint32_t widest = 0;
    for (uint32_t j = 0; j < dnum; ++j) {
        uint32_t lo = j * c.alpha, hi_ = std::min(L, (j + 1) * c.alpha);
        widest      = std::max(widest, product_bits(Q.data() + lo, hi_ - lo));
    }
    c.sizeP = (widest + 59) / 60;  // auxBits = 60 (AUXMODSIZE for 64-bit native integers)
    std::vector<uint64_t> P(c.sizeP);
    uint64_t walk = first_prime(60, m);
    for (uint32_t i = 0; i < c.sizeP; ++i) {
        do
            walk = previous_prime(walk, m);
        while (std::find(Q.begin(), Q.end(), walk) != Q.end());
        P[i] = walk;
    }
    c.moduli = Q;
    c.moduli.insert(c.moduli.end(), P.begin(), P.end());
    c.roots.resize(c.moduli.size());
#pragma omp parallel for schedule(dynamic)
    for (long i = 0; i < (long)c.moduli.size(); ++i)
        c.roots[i] = minimal_root_of_unity(m, c.moduli[i]);
    return c;
}

size_t bfv_table(const uint64_t* Q, uint32_t nQ, const uint64_t* R, uint32_t nR, uint64_t t, int which, void* out) {
    uint64_t* u = static_cast<uint64_t*>(out);
    double* d   = static_cast<double*>(out);
    // (QR / m)^-1 mod m for a modulus m that is Q[skipQ] or R[skipR]
    auto qr_hat_inv = [&](long skipQ, long skipR, uint64_t m) {
        return inv(mul(prod_mod(Q, nQ, skipQ, m), prod_mod(R, nR, skipR, m), m), m);
    };
    switch (which) {
        case 0:  // (Q/q_i)^-1 mod q_i                                                         :181-191
            for (uint32_t i = 0; i < nQ; ++i)
                u[i] = inv(prod_mod(Q, nQ, (long)i, Q[i]), Q[i]);
            return nQ;
        case 1:  // (Q/q_i) mod r_j, [j][i]                                                     :194-202
            for (uint32_t j = 0; j < nR; ++j)
                for (uint32_t i = 0; i < nQ; ++i)
                    u[(size_t)j * nQ + i] = prod_mod(Q, nQ, (long)i, R[j]);
            return (size_t)nQ * nR;
        case 2:  // k * Q mod r_j, [k][j], k = 0..sizeQ                                         :336-355
            for (uint32_t k = 0; k <= nQ; ++k)
                for (uint32_t j = 0; j < nR; ++j)
                    u[(size_t)k * nR + j] = mul(prod_mod(Q, nQ, -1, R[j]), k, R[j]);
            return (size_t)(nQ + 1) * nR;
        case 3:  // 1 / q_i in double                                                           :262-265
            for (uint32_t i = 0; i < nQ; ++i)
                d[i] = 1. / static_cast<double>(Q[i]);
            return nQ;
        case 4:  // q_i - R * (Q/q_i)^-1 mod q_i                                                :506-521
            for (uint32_t i = 0; i < nQ; ++i) {
                const uint64_t q = Q[i];
                u[i] = q - mul(prod_mod(R, nR, -1, q), inv(prod_mod(Q, nQ, (long)i, q), q), q);
            }
            return nQ;
        case 5:  // q_i^-1 mod r_j, [i][j]                                                      :524-530
            for (uint32_t i = 0; i < nQ; ++i)
                for (uint32_t j = 0; j < nR; ++j)
                    u[(size_t)i * nR + j] = inv(Q[i] % R[j], R[j]);
            return (size_t)nQ * nR;
        case 6:  // (R/r_j)^-1 mod r_j                                                          :361-370
            for (uint32_t j = 0; j < nR; ++j)
                u[j] = inv(prod_mod(R, nR, (long)j, R[j]), R[j]);
            return nR;
        case 7:  // (R/r_j) mod q_i, [i][j]                                                     :372-400
            for (uint32_t i = 0; i < nQ; ++i)
                for (uint32_t j = 0; j < nR; ++j)
                    u[(size_t)i * nR + j] = prod_mod(R, nR, (long)j, Q[i]);
            return (size_t)nQ * nR;
        case 8:  // k * R mod q_i, [k][i], k = 0..sizeR                                         :402-425
            for (uint32_t k = 0; k <= nR; ++k)
                for (uint32_t i = 0; i < nQ; ++i)
                    u[(size_t)k * nQ + i] = mul(prod_mod(R, nR, -1, Q[i]), k, Q[i]);
            return (size_t)(nR + 1) * nQ;
        case 9:  // 1 / r_j in double                                                           :429-432
            for (uint32_t j = 0; j < nR; ++j)
                d[j] = 1. / static_cast<double>(R[j]);
            return nR;
        case 10:  // floor(t Q (QR/s)^-1 / s) mod q_i, s = r_j (j < sizeR) or q_i (last column), [i][sizeR+1]   :565-593
            for (uint32_t i = 0; i < nQ; ++i) {
                const uint64_t q = Q[i];
                for (uint32_t j = 0; j < nR; ++j) {
                    // X = t Q c with c = (QR/r_j)^-1 mod r_j; X = r_j D + rem and q_i | X, so D = -rem / r_j (mod q_i)
                    const uint64_t r   = R[j];
                    const uint64_t rem = mul(mul(t % r, prod_mod(Q, nQ, -1, r), r), qr_hat_inv(-1, (long)j, r), r);
                    u[(size_t)i * (nR + 1) + j] = mul((q - rem % q) % q, inv(r % q, q), q);
                }
                // s = q_i divides Q: the quotient is exact, t (Q/q_i) c mod q_i
                u[(size_t)i * (nR + 1) + nR] = mul(mul(t % q, prod_mod(Q, nQ, (long)i, q), q), qr_hat_inv((long)i, -1, q), q);
            }
            return (size_t)nQ * (nR + 1);
        case 11:  // ((t Q (QR/r_j)^-1) mod r_j) / r_j, both converted to double first             :569-577
            for (uint32_t j = 0; j < nR; ++j) {
                const uint64_t r   = R[j];
                const uint64_t rem = mul(mul(t % r, prod_mod(Q, nQ, -1, r), r), qr_hat_inv(-1, (long)j, r), r);
                d[j]               = static_cast<double>(rem) / static_cast<double>(r);
            }
            return nR;
        default: return 0;
    }
}

}  // namespace hm
}  // namespace ofhe
