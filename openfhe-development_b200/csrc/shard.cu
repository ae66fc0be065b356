// shard.cu -- CKKS EvalMult + hybrid key switch with the DCRTPoly sharded BY RNS LIMB across G ranks (one GPU each).
//
// everything limb-local (tensor product, NTTs, evk inner product, ModDown tail) runs on the
// owner of the limb; the two fast base conversions need every source limb of a coefficient, so there are exactly two
// exchange steps per key switch, both all-gathers of INVERSE-NTT'd (and QHatInv-premultiplied) limbs:
//     gatherQ  [G][slotsQ][N]      the sizeQl source limbs of ModUp        (keyswitch-hybrid.cpp:359-365)
//     gatherP  [G][2][slotsP][N]   the sizeP limbs of both ModDown inputs  (dcrtpoly-impl.h:976-988)
// This file is the per-rank compute between the exchanges; the collective itself is issued by the caller (NCCL through
// torch.distributed, in place on the gather buffers) so the library stays free of a communication dependency.
//
// Ownership: Q limb i -> rank i mod G (local row i div G); P limb j -> rank (j + sizeQ) mod G (local row j div G).  Owned
// sets at a lower level are prefixes of the full-level ones, so evaluation-key shards never move.
// Local slabs: "Ql" = [nOwnQ][N], "QlP" = [nOwnQ + nOwnP][N] (Q rows first).  Key shard = [dnum][nKeyRows][N] with
// nKeyRows = nOwnQ(full) + nOwnP.
#include <memory>

#include "hks.h"
#include "hostmath.h"
#include "shard.h"

namespace ofhe {

std::unique_ptr<BConv> make_bconv(const Ring& r, const std::vector<uint32_t>& src_idx, const std::vector<uint32_t>& src_pos,
                                  const std::vector<uint32_t>& dst_idx, const std::vector<uint32_t>& dst_pos);

namespace {
void put(LimbMap& m, uint32_t idx, uint32_t pos) {
    OFHE_REQUIRE(m.n < (uint32_t)kMaxLimbs && pos < 65536, "shard: per-launch limb capacity exceeded");
    m.idx[m.n] = (uint16_t)idx;
    m.pos[m.n] = (uint16_t)pos;
    ++m.n;
}
}  // namespace

Shard::Shard(Ring* r, uint32_t sizeQ_, uint32_t sizeP_, uint32_t dnum_, uint32_t rank_, uint32_t world_)
    : ring(r), sizeQ(sizeQ_), sizeP(sizeP_), dnum(dnum_), rank(rank_), world(world_) {
    OFHE_REQUIRE(world >= 1 && rank < world, "shard: bad rank / world size");
    OFHE_REQUIRE(sizeQ >= 1 && sizeP >= 1 && dnum >= 1 && sizeQ + sizeP <= r->n_moduli, "shard: Q and P must be registered in the ring");
    alpha = (sizeQ + dnum - 1) / dnum;
    OFHE_REQUIRE(sizeQ > alpha * (dnum - 1), "shard: cannot distribute the towers into the requested number of digits");
    OFHE_REQUIRE(alpha <= 64, "shard: digits above 64 limbs are not supported");
    slotsP = (sizeP + world - 1) / world;
    for (uint32_t j = 0; j < sizeP; ++j)
        if (owner_p(j) == rank)
            ownP.push_back(j);
    nOwnQFull = 0;
    for (uint32_t i = 0; i < sizeQ; ++i)
        nOwnQFull += owner_q(i) == rank;
}
Shard::~Shard() {
    if (ws_d)
        cudaFree(ws_d);
}
uint64_t* Shard::scratch(uint32_t sizeQl) {
    const size_t need = workspace_words(sizeQl);
    if (need > ws_words) {
        if (ws_d) {
            OFHE_CUDA(cudaDeviceSynchronize());
            OFHE_CUDA(cudaFree(ws_d));
            ws_d = nullptr;
        }
        OFHE_CUDA(cudaMalloc(&ws_d, sizeof(uint64_t) * need));
        ws_words = need;
    }
    return ws_d;
}
Shard::Level::~Level() {
    if (fold_q_d)
        cudaFree(fold_q_d);
    if (fold_p_d)
        cudaFree(fold_p_d);
}

Shard::Level& Shard::level(uint32_t sizeQl) {
    OFHE_REQUIRE(sizeQl >= 1 && sizeQl <= sizeQ, "shard: level out of range");
    std::lock_guard<std::mutex> g(mu);
    auto it = levels.find(sizeQl);
    if (it != levels.end())
        return *it->second;
    OFHE_CUDA(cudaSetDevice(ring->ctx->device));
    const Ring& r = *ring;
    auto lv       = std::make_unique<Level>();
    lv->slotsQ    = (sizeQl + world - 1) / world;
    lv->n_digits  = std::min((sizeQl + alpha - 1) / alpha, dnum);
    for (uint32_t i = 0; i < sizeQl; ++i)
        if (owner_q(i) == rank)
            lv->ownQ.push_back(i);
    const uint32_t nq = (uint32_t)lv->ownQ.size(), np = (uint32_t)ownP.size(), nown = nq + np;
    lv->nOwnQ = nq;
    // limb maps of the local slabs
    lv->map_q.n = lv->map_qp.n = lv->map_q2.n = lv->map_p2_src.n = 0;
    for (uint32_t t = 0; t < nq; ++t) {
        put(lv->map_q, lv->ownQ[t], t);
        put(lv->map_qp, lv->ownQ[t], t);
    }
    for (uint32_t t = 0; t < np; ++t)
        put(lv->map_qp, sizeQ + ownP[t], nq + t);
    // phase A: INTT of the owned rows of c2 into this rank's segment of gatherQ, QHatInv of the limb's digit folded in
    {
        std::vector<uint64_t> fold(4 * (size_t)std::max(nq, 1u));
        for (uint32_t t = 0; t < nq; ++t) {
            const uint32_t i = lv->ownQ[t], j = i / alpha, s = j * alpha, ds = std::min(alpha, sizeQl - s);
            const uint64_t q = r.moduli[i];
            const uint64_t w = hm::inv(hm::prod_mod(r.moduli.data() + s, ds, (long)(i - s), q), q);
            make_inv_fold(r, i, w, &fold[4 * (size_t)t]);
            lv->gatherq_pos.pos[t] = (uint16_t)(rank * lv->slotsQ + t);
        }
        OFHE_CUDA(cudaMalloc(&lv->fold_q_d, sizeof(uint64_t) * fold.size()));
        h2d_blocking(lv->fold_q_d, fold.data(), sizeof(uint64_t) * fold.size());
    }
    // phase B: per digit, convert the gathered digit limbs into the OWNED complementary rows of that digit's local slab
    lv->map_modup_ntt.n = 0;
    for (uint32_t j = 0; j < lv->n_digits; ++j) {
        const uint32_t s = j * alpha, ds = std::min(alpha, sizeQl - s);
        std::vector<uint32_t> si, sp, di, dp;
        for (uint32_t i = s; i < s + ds; ++i) {
            si.push_back(i);
            sp.push_back(owner_q(i) * lv->slotsQ + i / world);
        }
        uint32_t first_own = UINT32_MAX, n_own_in_digit = 0;
        for (uint32_t t = 0; t < nq; ++t) {
            const uint32_t i = lv->ownQ[t];
            if (i >= s && i < s + ds) {
                first_own = std::min(first_own, t);
                ++n_own_in_digit;
            }
            else {
                di.push_back(i);
                dp.push_back(t);
                put(lv->map_modup_ntt, i, j * nown + t);
            }
        }
        for (uint32_t t = 0; t < np; ++t) {
            di.push_back(sizeQ + ownP[t]);
            dp.push_back(nq + t);
            put(lv->map_modup_ntt, sizeQ + ownP[t], j * nown + nq + t);
        }
        lv->own_first.push_back(n_own_in_digit ? first_own : 0);
        lv->own_count.push_back(n_own_in_digit);
        lv->modup.push_back(di.empty() ? nullptr : make_bconv(r, si, sp, di, dp));
    }
    // phase B tail: INTT of the owned P rows of (e0, e1) into this rank's segment of gatherP, PHatInv folded in
    {
        const uint64_t* P = r.moduli.data() + sizeQ;
        std::vector<uint64_t> fold(4 * (size_t)std::max(2 * np, 1u));
        uint32_t n = 0;
        for (uint32_t b = 0; b < 2; ++b)
            for (uint32_t t = 0; t < np; ++t, ++n) {
                const uint32_t jj = ownP[t];
                const uint64_t w  = hm::inv(hm::prod_mod(P, sizeP, (long)jj, P[jj]), P[jj]);
                make_inv_fold(r, sizeQ + jj, w, &fold[4 * (size_t)n]);
                put(lv->map_p2_src, sizeQ + jj, b * nown + nq + t);
                lv->gatherp_pos.pos[n] = (uint16_t)((rank * 2 + b) * slotsP + t);
            }
        OFHE_CUDA(cudaMalloc(&lv->fold_p_d, sizeof(uint64_t) * fold.size()));
        h2d_blocking(lv->fold_p_d, fold.data(), sizeof(uint64_t) * fold.size());
    }
    // phase C: all P limbs (gathered) -> the owned Q rows, both polynomials in one launch
    if (nq) {
        std::vector<uint32_t> si, sp, di, dp;
        for (uint32_t j = 0; j < sizeP; ++j) {
            si.push_back(sizeQ + j);
            sp.push_back(owner_p(j) * 2 * slotsP + j / world);
        }
        for (uint32_t t = 0; t < nq; ++t) {
            di.push_back(lv->ownQ[t]);
            dp.push_back(t);
        }
        lv->moddown = make_bconv(r, si, sp, di, dp);
        const uint64_t* P = r.moduli.data() + sizeQ;
        for (uint32_t b = 0; b < 2; ++b)
            for (uint32_t t = 0; t < nq; ++t) {
                put(lv->map_q2, lv->ownQ[t], b * nq + t);
                if (b == 0) {
                    const uint64_t q = r.moduli[lv->ownQ[t]];
                    lv->pinv.s[t]    = hm::inv(hm::prod_mod(P, sizeP, -1, q), q);
                    lv->pinv.ss[t]   = hm::shoup_const(lv->pinv.s[t], q);
                }
            }
    }
    // rescale: owned limbs below the last one keep their local rows; constants as in Hks::level
    lv->map_low.n = 0;
    if (sizeQl >= 2) {
        const uint32_t l = sizeQl - 1;
        for (uint32_t t = 0; t < nq; ++t) {
            const uint32_t i = lv->ownQ[t];
            if (i == l) {
                lv->last_row = (int)t;
                continue;
            }
            const uint64_t qi = r.moduli[i], B = hm::inv(r.moduli[l] % qi, qi);
            const uint32_t n  = lv->map_low.n;
            lv->rescaleA.s[n]  = qi - B;
            lv->rescaleA.ss[n] = hm::shoup_const(qi - B, qi);
            lv->rescaleB.s[n]  = B;
            lv->rescaleB.ss[n] = hm::shoup_const(B, qi);
            put(lv->map_low, i, t);
        }
    }
    auto& ref      = *lv;
    levels[sizeQl] = std::move(lv);
    return ref;
}

size_t Shard::workspace_words(uint32_t sizeQl) {
    Level& lv       = level(sizeQl);
    const size_t N  = ring->N, nq = lv.nOwnQ, nown = nq + ownP.size();
    return N * (nq + (size_t)dnum * nown + 2 * nown + 2 * nq);  // c2 | digits | e0,e1 | conv
}

void Shard::evalmult_a(uint32_t sizeQl, const uint64_t* ct1, const uint64_t* ct2, uint64_t* out, uint64_t* gatherQ,
                       cudaStream_t st, const PeerOut* peers) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, P = (size_t)lv.nOwnQ * N;
    uint64_t* ws   = scratch(sizeQl);
    uint64_t* c2   = ws;
    if (lv.nOwnQ == 0)
        return;
    launch_tensor(r, ct1, ct1 + P, ct2, ct2 + P, out, out + P, c2, lv.map_q, st);  // base-leveledshe.cpp:620-624
    launch_ntt_inv(r, c2, gatherQ, lv.map_q, st, &lv.gatherq_pos, lv.fold_q_d, peers);  // keyswitch-hybrid.cpp:359
}

void Shard::evalmult_b(uint32_t sizeQl, const uint64_t* gatherQ, const uint64_t* evk_a, const uint64_t* evk_b,
                       uint64_t* gatherP, cudaStream_t st, const PeerOut* peers) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, nq = lv.nOwnQ, np = ownP.size(), nown = nq + np;
    uint64_t* ws     = scratch(sizeQl);
    uint64_t* c2     = ws;
    uint64_t* digits = c2 + nq * N;
    uint64_t* e01    = digits + (size_t)dnum * nown * N;
    if (nown == 0)
        return;
    // :360-365 -- every digit's conversion in ONE tensor-core launch when they all qualify
    bool all_tc = lv.n_digits <= 4;
    BConvJob jobs[4];
    uint32_t njobs = 0;
    for (uint32_t j = 0; j < lv.n_digits && all_tc; ++j)
        if (lv.modup[j]) {
            all_tc        = bconv_tc_eligible(r, *lv.modup[j]);
            jobs[njobs++] = BConvJob{lv.modup[j].get(), gatherQ, digits + (size_t)j * nown * N, 1, 0, 0};
        }
    if (all_tc && njobs)
        launch_bconv_tc(r, jobs, njobs, /*premultiplied=*/true, st);
    for (uint32_t j = 0; j < lv.n_digits; ++j) {
        uint64_t* D = digits + (size_t)j * nown * N;
        if (lv.modup[j] && !all_tc)
            launch_bconv(r, *lv.modup[j], gatherQ, D, 1, 0, 0, /*premultiplied=*/true, st);
        if (lv.own_count[j])                                                                   // :372-373
            OFHE_CUDA(cudaMemcpyAsync(D + (size_t)lv.own_first[j] * N, c2 + (size_t)lv.own_first[j] * N,
                                      sizeof(uint64_t) * lv.own_count[j] * N, cudaMemcpyDeviceToDevice, st));
    }
    launch_ntt_fwd(r, digits, digits, lv.map_modup_ntt, st);  // :366
    // :402-435 on the owned rows; key rows of the P part sit after ALL owned Q rows of the full-level key
    launch_ks_inner_product(r, lv.n_digits, digits, nown * N, evk_a, evk_b, (size_t)(nOwnQFull + np) * N, (uint32_t)nq,
                            (uint32_t)(nOwnQFull - nq), e01, e01 + nown * N, lv.map_qp, st);
    if (np)
        launch_ntt_inv(r, e01, gatherP, lv.map_p2_src, st, &lv.gatherp_pos, lv.fold_p_d, peers);  // dcrtpoly-impl.h:976-984
}

void Shard::evalmult_c(uint32_t sizeQl, const uint64_t* gatherP, uint64_t* out, cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, nq = lv.nOwnQ, np = ownP.size(), nown = nq + np;
    if (nq == 0)
        return;
    uint64_t* ws     = scratch(sizeQl);
    uint64_t* digits = ws + nq * N;
    uint64_t* e01    = digits + (size_t)dnum * nown * N;
    uint64_t* conv   = e01 + 2 * nown * N;
    launch_bconv(r, *lv.moddown, gatherP, conv, 2, (size_t)slotsP * N, nq * N, /*premultiplied=*/true, st);  // :987-988
    launch_ntt_fwd(r, conv, conv, lv.map_q2, st);                                                              // :1001
    for (uint32_t b = 0; b < 2; ++b)  // (e - conv) * P^-1 + (c0 | c1)            :1002, base-leveledshe.cpp:209-210
        launch_sub_mul_scalars(r, out + b * nq * N, e01 + b * nown * N, conv + b * nq * N, lv.pinv, out + b * nq * N,
                               lv.map_q, st);
}

void Shard::rotate_a(uint32_t sizeQl, const uint64_t* ct, uint64_t* tmp, uint64_t* gatherQ, cudaStream_t st,
                     const PeerOut* peers) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, P = (size_t)lv.nOwnQ * N;
    uint64_t* c2   = scratch(sizeQl);  // the polynomial to switch sits where evalmult_a leaves c2 (evalmult_b reads it)
    if (lv.nOwnQ == 0)
        return;
    // KeySwitchInPlace on (c0, c1): (c0 + ab0, ab1) with ab = KeySwitchCore(c1)           (keyswitch-hybrid.cpp:199-214)
    OFHE_CUDA(cudaMemcpyAsync(c2, ct + P, sizeof(uint64_t) * P, cudaMemcpyDeviceToDevice, st));
    OFHE_CUDA(cudaMemcpyAsync(tmp, ct, sizeof(uint64_t) * P, cudaMemcpyDeviceToDevice, st));
    OFHE_CUDA(cudaMemsetAsync(tmp + P, 0, sizeof(uint64_t) * P, st));
    launch_ntt_inv(r, c2, gatherQ, lv.map_q, st, &lv.gatherq_pos, lv.fold_q_d, peers);
}

void Shard::rotate_c(uint32_t sizeQl, uint32_t k, const uint64_t* gatherP, uint64_t* tmp, uint64_t* out, cudaStream_t st) {
    Level& lv = level(sizeQl);
    if (lv.nOwnQ == 0)
        return;
    evalmult_c(sizeQl, gatherP, tmp, st);
    // both polynomials through the automorphism gather, limb by limb                         (base-leveledshe.cpp:414-420)
    launch_automorphism(*ring, out, tmp, ring_automap(ring, k, st), 2 * lv.nOwnQ, st);
}

void Shard::rescale_a(uint32_t sizeQl, const uint64_t* ct, uint64_t* gatherL, cudaStream_t st, const PeerOut* peers) {
    OFHE_REQUIRE(sizeQl >= 2, "rescale needs at least two limbs");
    Level& lv = level(sizeQl);
    if (lv.last_row < 0)
        return;  // another rank owns the dropped limb
    LimbMap two;
    PosMap dst;
    two.n = 0;
    for (uint32_t b = 0; b < 2; ++b) {
        dst.pos[two.n] = (uint16_t)b;
        put(two, sizeQl - 1, b * lv.nOwnQ + (uint32_t)lv.last_row);
    }
    launch_ntt_inv(*ring, ct, gatherL, two, st, &dst, nullptr, peers);  // dcrtpoly-impl.h:696-697, both polynomials
}

void Shard::rescale_b(uint32_t sizeQl, const uint64_t* ct, const uint64_t* gatherL, uint64_t* out, cudaStream_t st) {
    OFHE_REQUIRE(sizeQl >= 2, "rescale needs at least two limbs");
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, nq = lv.nOwnQ, nlow = lv.map_low.n;
    if (nlow == 0)
        return;
    uint64_t* t = scratch(sizeQl);
    for (uint32_t b = 0; b < 2; ++b) {
        launch_switch_modulus_scale(r, t, gatherL + b * N, sizeQl - 1, lv.rescaleA, lv.map_low, st);  // :703-704
        launch_ntt_fwd(r, t, t, lv.map_low, st);                                                      // :705-706
        launch_mul_scalars_add(r, out + b * nlow * N, ct + b * nq * N, lv.rescaleB, t, lv.map_low, st);  // :707-708
    }
}

}  // namespace ofhe
