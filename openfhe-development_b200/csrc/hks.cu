// hks.cu -- hybrid key switching (ModUp -> evk inner product -> ModDown), CKKS EvalMult / rescale / rotation,
// composed from the kernels of ntt.cu / baseconv.cu / elementwise.cu on device-resident slabs.
//
// Device-side replacement for
//   KeySwitchHYBRID::EvalKeySwitchPrecomputeCore   pke/lib/keyswitch/keyswitch-hybrid.cpp:314-379   (ModUp)
//   KeySwitchHYBRID::EvalFastKeySwitchCoreExt      :402-435                                          (inner product)
//   KeySwitchHYBRID::EvalFastKeySwitchCore / KeySwitchCore / KeySwitchInPlace   :381-400,308-312,199-214
//   DCRTPolyImpl::ApproxModDown                    core/include/lattice/hal/default/dcrtpoly-impl.h:966-1005
//   DCRTPolyImpl::DropLastElementAndScale          dcrtpoly-impl.h:693-712
//   LeveledSHEBase::EvalMult / EvalMultCore / EvalAutomorphism   pke/lib/schemebase/base-leveledshe.cpp:202-214,
//                                                                608-646,382-422
// and for the table construction of CryptoParametersRNS::PrecomputeCRTTables
// (pke/lib/schemerns/rns-cryptoparameters.cpp:199-349) and CryptoParametersCKKSRNS::PrecomputeCRTTables
// (pke/lib/scheme/ckksrns/ckksrns-cryptoparameters.cpp:60-83): every table is a function of the moduli alone and
// is rebuilt here with word-size modular arithmetic, then kept on the device.
#include <memory>

#include "hks.h"
#include "hostmath.h"

namespace ofhe {

std::unique_ptr<BConv> make_bconv(const Ring& r, const std::vector<uint32_t>& src_idx,
                                  const std::vector<uint32_t>& src_pos, const std::vector<uint32_t>& dst_idx,
                                  const std::vector<uint32_t>& dst_pos) {
    // source basis {q_i}: QHatInv_i = (prod_{k != i} q_k)^-1 mod q_i;  matrix M[i][j] = prod_{k != i} q_k mod p_j
    const size_t nq = src_idx.size(), np = dst_idx.size();
    OFHE_REQUIRE(nq >= 1 && nq <= 64 && np <= (size_t)kMaxLimbs, "base conversion: unsupported basis size");
    std::vector<uint64_t> q(nq), p(np);
    for (size_t i = 0; i < nq; ++i)
        q[i] = r.moduli[src_idx[i]];
    for (size_t j = 0; j < np; ++j)
        p[j] = r.moduli[dst_idx[j]];
    std::vector<uint64_t> hinv(2 * nq);
    std::vector<uint32_t> mat(2 * nq * np);
    for (size_t i = 0; i < nq; ++i) {
        uint64_t w      = hm::inv(hm::prod_mod(q.data(), nq, (long)i, q[i]), q[i]);
        hinv[2 * i]     = w;
        hinv[2 * i + 1] = hm::shoup_const(w, q[i]);
        for (size_t j = 0; j < np; ++j) {
            uint64_t m               = hm::prod_mod(q.data(), nq, (long)i, p[j]);
            mat[2 * (i * np + j)]     = (uint32_t)(m & ((1u << 30) - 1));
            mat[2 * (i * np + j) + 1] = (uint32_t)(m >> 30);
        }
    }
    auto t   = std::make_unique<BConv>();
    t->sizeQ = (uint32_t)nq;
    t->sizeP = (uint32_t)np;
    t->qmap.n = (uint32_t)nq;
    t->pmap.n = (uint32_t)np;
    for (size_t i = 0; i < nq; ++i) {
        t->qmap.idx[i] = (uint16_t)src_idx[i];
        t->qmap.pos[i] = (uint16_t)src_pos[i];
    }
    for (size_t j = 0; j < np; ++j) {
        t->pmap.idx[j] = (uint16_t)dst_idx[j];
        t->pmap.pos[j] = (uint16_t)dst_pos[j];
    }
    OFHE_CUDA(cudaMalloc(&t->qhatinv_d, sizeof(uint64_t) * hinv.size()));
    OFHE_CUDA(cudaMalloc(&t->mat_d, sizeof(uint32_t) * mat.size()));
    h2d_blocking(t->qhatinv_d, hinv.data(), sizeof(uint64_t) * hinv.size());
    h2d_blocking(t->mat_d, mat.data(), sizeof(uint32_t) * mat.size());
    bconv_finalize(r, *t);
    return t;
}

namespace {

void fill_scalar(ScalarVec& sv, uint32_t i, uint64_t s, uint64_t q) {
    sv.s[i]  = s;
    sv.ss[i] = hm::shoup_const(s, q);
}

}  // namespace

Hks::Hks(Ring* r, uint32_t sizeQ_, uint32_t sizeP_, uint32_t dnum_) : ring(r), sizeQ(sizeQ_), sizeP(sizeP_), dnum(dnum_) {
    OFHE_REQUIRE(sizeQ >= 1 && dnum >= 1 && sizeQ + sizeP <= r->n_moduli, "hks: Q and P must be registered in the ring");
    OFHE_REQUIRE(sizeQ + sizeP <= 128, "hks: at most 128 limbs in QP");
    alpha = (sizeQ + dnum - 1) / dnum;  // rns-cryptoparameters.cpp:86-95
    OFHE_REQUIRE(sizeQ > alpha * (dnum - 1), "hks: cannot distribute the towers into the requested number of digits");
    OFHE_REQUIRE(alpha <= 64, "hks: digits above 64 limbs are not supported");
    // [P]_{q_i}, [P^-1]_{q_i}  (rns-cryptoparameters.cpp:199-212)
    const uint64_t* P = r->moduli.data() + sizeQ;
    PModq.resize(sizeQ);
    PInvModq.resize(sizeQ);
    for (uint32_t i = 0; i < sizeQ; ++i) {
        uint64_t q  = r->moduli[i];
        PModq[i]    = hm::prod_mod(P, sizeP, -1, q);
        PInvModq[i] = sizeP ? hm::inv(PModq[i], q) : 1;
        pmodq_sv.s[i]  = PModq[i];
        pmodq_sv.ss[i] = hm::shoup_const(PModq[i], q);
    }
}
Hks::~Hks() = default;
Hks::Level::~Level() {
    if (fold_modup_d)
        cudaFree(fold_modup_d);
    if (fold_moddown_d)
        cudaFree(fold_moddown_d);
    if (pinv2_d)
        cudaFree(pinv2_d);
}

Hks::Level& Hks::level(uint32_t sizeQl) {
    OFHE_REQUIRE(sizeQl >= 1 && sizeQl <= sizeQ, "hks: level out of range");
    std::lock_guard<std::mutex> g(mu);
    auto it = levels.find(sizeQl);
    if (it != levels.end())
        return *it->second;
    OFHE_CUDA(cudaSetDevice(ring->ctx->device));
    auto lv        = std::make_unique<Level>();
    const Ring& r  = *ring;
    const uint32_t sizeQlP = sizeQl + sizeP;
    lv->n_digits   = std::min((sizeQl + alpha - 1) / alpha, dnum);  // keyswitch-hybrid.cpp:327-329
    // modulus index of row i of a Ql.P slab
    std::vector<uint32_t> qlp(sizeQlP);
    for (uint32_t i = 0; i < sizeQlP; ++i)
        qlp[i] = i < sizeQl ? i : sizeQ + (i - sizeQl);
    lv->map_QlP = make_map(qlp.data(), sizeQlP);
    lv->map_Ql  = make_map(qlp.data(), sizeQl);
    // ---- ModUp: digit j converts rows [j*alpha, j*alpha+ds) of the INTT'd input into all the other rows of its own
    //      Ql.P slab (complementary partition, rns-cryptoparameters.cpp:252-294, written in place: no assembly copy)
    // The one forward-NTT launch over every converted row of every digit is ordered MODULUS-major (rows of different
    // digits that share a modulus are adjacent in the grid): their twiddle rows are fetched from HBM once per launch.
    lv->map_modup_ntt.n = 0;
    for (uint32_t j = 0; j < lv->n_digits; ++j) {
        const uint32_t s = j * alpha, ds = std::min(alpha, sizeQl - s);
        std::vector<uint32_t> si, sp, di, dp;
        for (uint32_t i = 0; i < ds; ++i) {
            si.push_back(s + i);
            sp.push_back(s + i);  // row in the [sizeQl][N] coefficient-format copy of c
        }
        for (uint32_t i = 0; i < sizeQlP; ++i)
            if (i < s || i >= s + ds) {
                di.push_back(qlp[i]);
                dp.push_back(i);  // row in digit j's own [sizeQlP][N] slab
            }
        lv->modup.push_back(make_bconv(r, si, sp, di, dp));
    }
    for (uint32_t i = 0; i < sizeQlP; ++i)
        for (uint32_t j = 0; j < lv->n_digits; ++j) {
            const uint32_t s = j * alpha, ds = std::min(alpha, sizeQl - s);
            if (i >= s && i < s + ds)
                continue;  // the digit's own rows keep the original evaluation-format limbs
            const uint32_t n = lv->map_modup_ntt.n++;
            OFHE_REQUIRE(n < (uint32_t)kMaxLimbs && j * sizeQlP + i < 65536, "hks: ModUp exceeds the per-launch limb capacity");
            lv->map_modup_ntt.idx[n] = (uint16_t)qlp[i];
            lv->map_modup_ntt.pos[n] = (uint16_t)(j * sizeQlP + i);
        }
    {   // QHatInv of every source limb folded into the constants of the inverse NTT that precedes the conversion
        std::vector<uint64_t> fold(4 * (size_t)sizeQl);
        for (uint32_t j = 0; j < lv->n_digits; ++j) {
            const uint32_t s = j * alpha, ds = std::min(alpha, sizeQl - s);
            for (uint32_t i = 0; i < ds; ++i) {
                uint64_t q = r.moduli[s + i];
                uint64_t w = hm::inv(hm::prod_mod(r.moduli.data() + s, ds, (long)i, q), q);
                make_inv_fold(r, s + i, w, &fold[4 * (size_t)(s + i)]);
            }
        }
        OFHE_CUDA(cudaMalloc(&lv->fold_modup_d, sizeof(uint64_t) * fold.size()));
        h2d_blocking(lv->fold_modup_d, fold.data(), sizeof(uint64_t) * fold.size());
    }
    if (sizeP) {
        const uint64_t* P = r.moduli.data() + sizeQ;
        std::vector<uint64_t> fold(8 * (size_t)sizeP);
        for (uint32_t j = 0; j < sizeP; ++j)  // launch order of map_P2: modulus-major (j, then polynomial b)
            for (uint32_t b = 0; b < 2; ++b) {
                uint64_t w = hm::inv(hm::prod_mod(P, sizeP, (long)j, P[j]), P[j]);
                make_inv_fold(r, sizeQ + j, w, &fold[4 * (size_t)(2 * j + b)]);
            }
        OFHE_CUDA(cudaMalloc(&lv->fold_moddown_d, sizeof(uint64_t) * fold.size()));
        h2d_blocking(lv->fold_moddown_d, fold.data(), sizeof(uint64_t) * fold.size());
    }
    // ---- ModDown: P rows -> Ql rows  (PHatInvModp / PHatModq, rns-cryptoparameters.cpp:214-229)
    if (sizeP) {
        std::vector<uint32_t> si, sp, di, dp;
        for (uint32_t j = 0; j < sizeP; ++j) {
            si.push_back(sizeQ + j);
            sp.push_back(j);
        }
        for (uint32_t i = 0; i < sizeQl; ++i) {
            di.push_back(i);
            dp.push_back(i);
        }
        lv->moddown = make_bconv(r, si, sp, di, dp);
    }
    // two-polynomial maps used by the batched ModDown
    // (both in modulus-major launch order: the two polynomials' rows of one modulus are adjacent)
    lv->map_P2.n = 0;
    lv->map_Ql2.n = 0;
    for (uint32_t j = 0; j < sizeP; ++j)
        for (uint32_t b = 0; b < 2; ++b) {
            uint32_t n           = lv->map_P2.n++;
            lv->map_P2.idx[n]    = (uint16_t)(sizeQ + j);
            lv->map_P2.pos[n]    = (uint16_t)(b * sizeQlP + sizeQl + j);  // P rows of e_b inside [2][sizeQlP][N]
            lv->gather_P2.pos[n] = (uint16_t)(b * sizeP + j);             // -> row of p_tmp [2][sizeP][N]
        }
    for (uint32_t i = 0; i < sizeQl; ++i)
        for (uint32_t b = 0; b < 2; ++b) {
            uint32_t n         = lv->map_Ql2.n++;
            lv->map_Ql2.idx[n] = (uint16_t)i;
            lv->map_Ql2.pos[n] = (uint16_t)(b * sizeQl + i);
            fill_scalar(lv->pinv2, b * sizeQl + i, PInvModq[i], r.moduli[i]);
            lv->tail_erow.pos[n] = (uint16_t)(b * sizeQlP + i);
            lv->tail_drow.pos[n] = (uint16_t)(b * sizeQl + i);
        }
    {
        std::vector<uint64_t> pv(4 * (size_t)sizeQl);
        for (uint32_t n = 0; n < 2 * sizeQl; ++n) {
            const uint32_t i = lv->map_Ql2.idx[n];
            pv[2 * n]     = lv->pinv2.s[i];
            pv[2 * n + 1] = lv->pinv2.ss[i];
        }
        OFHE_CUDA(cudaMalloc(&lv->pinv2_d, sizeof(uint64_t) * pv.size()));
        h2d_blocking(lv->pinv2_d, pv.data(), sizeof(uint64_t) * pv.size());
    }
    // ---- CKKS rescale constants for dropping limb l = sizeQl-1 (ckksrns-cryptoparameters.cpp:60-83):
    //      A_i = floor(QlInv * Ql / q_l) mod q_i = -q_l^-1 mod q_i (q_i | Ql),  B_i = q_l^-1 mod q_i
    if (sizeQl >= 2) {
        const uint32_t l = sizeQl - 1;
        for (uint32_t i = 0; i < l; ++i) {
            uint64_t qi = r.moduli[i], B = hm::inv(r.moduli[l] % qi, qi);
            fill_scalar(lv->rescaleA, i, qi - B, qi);
            fill_scalar(lv->rescaleB, i, B, qi);
        }
    }
    auto& ref = *lv;
    levels[sizeQl] = std::move(lv);
    return ref;
}

size_t Hks::workspace_words(uint32_t sizeQl) const {
    const size_t N = ring->N, sizeQlP = sizeQl + sizeP;
    //  coeff copy of c | digits | e0,e1 | P rows (coefficient format) of e0,e1 | converted Ql rows of both | c2
    return N * ((size_t)sizeQl + (size_t)dnum * sizeQlP + 2 * sizeQlP + 2 * sizeP + 2 * (size_t)sizeQl + sizeQl);
}

size_t Hks::max_workspace_words(uint32_t sizeQl) const {
    const size_t N = ring->N, P = (size_t)sizeQl * N, E = (size_t)(sizeQl + sizeP) * N;
    return std::max(workspace_words(sizeQl) + 2 * P, 2 * E + 2 * sizeP * N + 4 * P);
}

void Hks::modup(uint32_t sizeQl, const uint64_t* c, uint64_t* digits, uint64_t* coeff_tmp, cudaStream_t st,
                bool own_rows) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, sizeQlP = sizeQl + sizeP;
    // 1. all digits at once: coefficient-format copy of c                            (keyswitch-hybrid.cpp:359)
    launch_ntt_inv(r, c, coeff_tmp, lv.map_Ql, st, nullptr, lv.fold_modup_d);  // y_i = INTT(c)_i * QHatInv_i
    // 2. fast base conversion straight into the complementary rows of D_j               (:360-365): every digit in
    //    ONE tensor-core launch when the digits qualify (<= 16 limbs each), else one CUDA-core launch per digit
    bool all_tc = lv.n_digits <= 4;
    for (uint32_t j = 0; j < lv.n_digits && all_tc; ++j)
        all_tc = bconv_tc_eligible(r, *lv.modup[j]);
    if (all_tc) {
        BConvJob jobs[4];
        for (uint32_t j = 0; j < lv.n_digits; ++j)
            jobs[j] = BConvJob{lv.modup[j].get(), coeff_tmp, digits + (size_t)j * sizeQlP * N, 1, 0, 0};
        launch_bconv_tc(r, jobs, lv.n_digits, /*premultiplied=*/true, st);
    }
    for (uint32_t j = 0; j < lv.n_digits; ++j) {
        uint64_t* D      = digits + (size_t)j * sizeQlP * N;
        const uint32_t s = j * alpha, ds = std::min(alpha, sizeQl - s);
        if (!all_tc)
            launch_bconv(r, *lv.modup[j], coeff_tmp, D, 1, 0, 0, /*premultiplied=*/true, st);
        // 3. the digit's own rows keep the ORIGINAL evaluation-format limbs of c          (:372-373); inside the fused
        //    key switch the inner product reads them from c directly and the copies are skipped
        if (own_rows)
            OFHE_CUDA(cudaMemcpyAsync(D + (size_t)s * N, c + (size_t)s * N, sizeof(uint64_t) * ds * N,
                                      cudaMemcpyDeviceToDevice, st));
    }
    // 4. one launch transforms every converted row of every digit back to EVALUATION      (:366)
    launch_ntt_fwd(r, digits, digits, lv.map_modup_ntt, st);
}

void Hks::inner_product(uint32_t sizeQl, uint32_t n_digits, const uint64_t* digits, const EvalKey& evk, uint64_t* e01,
                        cudaStream_t st, const uint64_t* c_own) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, sizeQlP = sizeQl + sizeP, sizeQP = sizeQ + sizeP;
    OFHE_REQUIRE(evk.hks == this, "evaluation key belongs to a different key-switch context");
    OFHE_REQUIRE(n_digits <= evk.dnum, "more digits than the evaluation key holds");
    launch_ks_inner_product(r, n_digits, digits, sizeQlP * N, evk.a_d, evk.b_d, sizeQP * N, sizeQl, sizeQ - sizeQl, e01,
                            e01 + sizeQlP * N, lv.map_QlP, st, c_own, alpha);
}

void Hks::mod_down2(uint32_t sizeQl, const uint64_t* e01, uint64_t* out01, const uint64_t* addend01, uint64_t* p_tmp,
                    uint64_t* conv_tmp, cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, sizeQlP = sizeQl + sizeP;
    OFHE_REQUIRE(sizeP >= 1, "ModDown needs a non-empty auxiliary basis");
    // P rows of both polynomials -> coefficient format, gathered into p_tmp [2][sizeP][N]   (dcrtpoly-impl.h:976-984)
    launch_ntt_inv(r, e01, p_tmp, lv.map_P2, st, &lv.gather_P2, lv.fold_moddown_d);  // times PHatInvModp
    // P -> Ql conversion of both polynomials in one launch                                     (:987-988)
    launch_bconv(r, *lv.moddown, p_tmp, conv_tmp, 2, (size_t)sizeP * N, (size_t)sizeQl * N, /*premultiplied=*/true, st);
    // back to EVALUATION (:1001) and out = (e - conv) * P^-1 (+ addend) (:1002): the tail is the transform's epilogue
    if (ntt_fwd_tail_fused(r)) {
        NttTail tl{e01, addend01, out01, lv.pinv2_d, lv.tail_erow, lv.tail_drow};
        launch_ntt_fwd(r, conv_tmp, conv_tmp, lv.map_Ql2, st, nullptr, &tl);
        return;
    }
    launch_ntt_fwd(r, conv_tmp, conv_tmp, lv.map_Ql2, st);
    for (uint32_t b = 0; b < 2; ++b) {
        ScalarVec sv;
        for (uint32_t i = 0; i < sizeQl; ++i) {
            sv.s[i]  = lv.pinv2.s[i];
            sv.ss[i] = lv.pinv2.ss[i];
        }
        launch_sub_mul_scalars(r, out01 + (size_t)b * sizeQl * N, e01 + (size_t)b * sizeQlP * N,
                               conv_tmp + (size_t)b * sizeQl * N, sv,
                               addend01 ? addend01 + (size_t)b * sizeQl * N : nullptr, lv.map_Ql, st);
    }
}

void Hks::keyswitch_core(uint32_t sizeQl, const uint64_t* c, const EvalKey& evk, uint64_t* out01,
                         const uint64_t* addend01, uint64_t* ws, cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const size_t N = ring->N, sizeQlP = sizeQl + sizeP;
    uint64_t* coeff  = ws;
    uint64_t* digits = coeff + (size_t)sizeQl * N;
    uint64_t* e01    = digits + (size_t)dnum * sizeQlP * N;
    uint64_t* p_tmp  = e01 + 2 * sizeQlP * N;
    uint64_t* conv   = p_tmp + 2 * (size_t)sizeP * N;
    modup(sizeQl, c, digits, coeff, st, /*own_rows=*/false);
    inner_product(sizeQl, lv.n_digits, digits, evk, e01, st, /*c_own=*/c);
    mod_down2(sizeQl, e01, out01, addend01, p_tmp, conv, st);
}

void Hks::evalmult(uint32_t sizeQl, const uint64_t* ct1, const uint64_t* ct2, const EvalKey& evk, uint64_t* out,
                   cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, P = (size_t)sizeQl * N;
    uint64_t* ws = static_cast<uint64_t*>(r.ctx->get_scratch(sizeof(uint64_t) * workspace_words(sizeQl), st));
    uint64_t* c2 = ws + (workspace_words(sizeQl) - P);
    // tensor product: out = (c0, c1), c2 to scratch                                     (base-leveledshe.cpp:620-624)
    launch_tensor(r, ct1, ct1 + P, ct2, ct2 + P, out, out + P, c2, lv.map_Ql, st);
    // relinearise: (c0, c1) += KeySwitchCore(c2)                                          (base-leveledshe.cpp:207-211)
    keyswitch_core(sizeQl, c2, evk, out, out, ws, st);
}

void Hks::rescale(uint32_t sizeQl, const uint64_t* in, uint64_t* out, uint32_t n_polys, cudaStream_t st) {
    OFHE_REQUIRE(sizeQl >= 2, "rescale needs at least two limbs");
    Level& lv      = level(sizeQl);
    Level& lvm     = level(sizeQl - 1);
    const Ring& r  = *ring;
    const size_t N = r.N;
    const uint32_t l = sizeQl - 1;
    uint64_t* ws   = static_cast<uint64_t*>(r.ctx->get_scratch(sizeof(uint64_t) * N * (1 + (size_t)l), st));
    uint64_t* z    = ws;
    uint64_t* t    = ws + N;
    uint32_t last  = l;
    LimbMap one    = make_map(&last, 1);
    for (uint32_t k = 0; k < n_polys; ++k) {
        const uint64_t* x = in + (size_t)k * sizeQl * N;
        uint64_t* o       = out + (size_t)k * l * N;
        one.pos[0]        = 0;
        launch_ntt_inv(r, x + (size_t)l * N, z, one, st);                       // dcrtpoly-impl.h:696-697
        launch_switch_modulus_scale(r, t, z, l, lv.rescaleA, lvm.map_Ql, st);   // :703-704
        launch_ntt_fwd(r, t, t, lvm.map_Ql, st);                                // :705-706
        launch_mul_scalars_add(r, o, x, lv.rescaleB, t, lvm.map_Ql, st);        // :707-708
    }
}

void Hks::rotate(uint32_t sizeQl, uint32_t k, const uint64_t* ct, const EvalKey& evk, uint64_t* out, cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, P = (size_t)sizeQl * N;
    uint64_t* ws  = static_cast<uint64_t*>(r.ctx->get_scratch(sizeof(uint64_t) * (workspace_words(sizeQl) + 2 * P), st));
    uint64_t* tmp = ws + workspace_words(sizeQl);
    // KeySwitchInPlace on a 2-element ciphertext: (c0 + ab0, ab1), ab = KeySwitchCore(c1)   (keyswitch-hybrid.cpp:199-214)
    // the addend for the second polynomial is absent, so run the two ModDown tails separately via a zeroed addend
    OFHE_CUDA(cudaMemcpyAsync(tmp, ct, sizeof(uint64_t) * P, cudaMemcpyDeviceToDevice, st));
    OFHE_CUDA(cudaMemsetAsync(tmp + P, 0, sizeof(uint64_t) * P, st));
    keyswitch_core(sizeQl, ct + P, evk, tmp, tmp, ws, st);
    // both polynomials through the automorphism gather                                   (base-leveledshe.cpp:414-420)
    const uint32_t* map = ring_automap(ring, k, st);
    launch_automorphism(r, out, tmp, map, 2 * sizeQl, st);
    (void)lv;
}

// EvalFastRotation (base-leveledshe.cpp:433-463): ba = EvalFastKeySwitchCore(digits, key_k) = inner product + ModDown,
// out = (auto_k(ba0 + c0), auto_k(ba1)).  The ModUp that produced `digits` is shared by all rotations of the ciphertext.
void Hks::fast_rotation(uint32_t sizeQl, uint32_t k, const uint64_t* ct, const uint64_t* digits, const EvalKey& evk,
                        uint64_t* out, cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, P = (size_t)sizeQl * N, E = (size_t)(sizeQl + sizeP) * N;
    uint64_t* ws    = static_cast<uint64_t*>(r.ctx->get_scratch(sizeof(uint64_t) * (2 * E + 2 * sizeP * N + 4 * P), st));
    uint64_t* e01   = ws;
    uint64_t* p_tmp = e01 + 2 * E;
    uint64_t* conv  = p_tmp + 2 * (size_t)sizeP * N;
    uint64_t* tmp   = conv + 2 * P;
    OFHE_CUDA(cudaMemcpyAsync(tmp, ct, sizeof(uint64_t) * P, cudaMemcpyDeviceToDevice, st));  // addend (c0, 0)
    OFHE_CUDA(cudaMemsetAsync(tmp + P, 0, sizeof(uint64_t) * P, st));
    inner_product(sizeQl, lv.n_digits, digits, evk, e01, st);                                 // keyswitch-hybrid.cpp:402-435
    mod_down2(sizeQl, e01, tmp, tmp, p_tmp, conv, st);                                        // :389-399
    launch_automorphism(r, out, tmp, ring_automap(ring, k, st), 2 * sizeQl, st);              // base-leveledshe.cpp:455-457
}

// EvalFastRotationExt (ckksrns-leveledshe.cpp:534-582): cTilda = EvalFastKeySwitchCoreExt(digits, key_k) stays in QlP;
// addFirst adds [c0 * (P mod q_i) | 0] to cTilda0; both polynomials through auto_k on all sizeQl + sizeP limbs.
void Hks::fast_rotation_ext(uint32_t sizeQl, uint32_t k, bool add_first, const uint64_t* ct, const uint64_t* digits,
                            const EvalKey& evk, uint64_t* out_ext, cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, E = (size_t)(sizeQl + sizeP) * N;
    uint64_t* e01  = static_cast<uint64_t*>(r.ctx->get_scratch(sizeof(uint64_t) * 2 * E, st));
    inner_product(sizeQl, lv.n_digits, digits, evk, e01, st);
    if (add_first)  // :561-570
        launch_mul_scalars_add(r, e01, ct, pmodq_sv, e01, lv.map_Ql, st);
    launch_automorphism(r, out_ext, e01, ring_automap(ring, k, st), 2 * (sizeQl + sizeP), st);  // :576-577
}

// KeySwitchExt (keyswitch-hybrid.cpp:217-243): lift to QlP as [c_k * (P mod q_i) | 0]; c_0 only if addFirst
void Hks::keyswitch_ext(uint32_t sizeQl, bool add_first, const uint64_t* ct, uint64_t* out_ext, cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, P = (size_t)sizeQl * N, E = (size_t)(sizeQl + sizeP) * N;
    OFHE_CUDA(cudaMemsetAsync(out_ext, 0, sizeof(uint64_t) * 2 * E, st));
    for (uint32_t b = add_first ? 0 : 1; b < 2; ++b)
        launch_mul_scalars(r, out_ext + b * E, ct + b * P, pmodq_sv, lv.map_Ql, st);
}

// ---- double-hoisted BSGS linear transform, one level ------------------------------------------------------------
// Reference: the body of the `for s` loops of FHECKKSRNS::EvalCoeffsToSlots (ckksrns-fhe.cpp:1934-1987) and
// EvalSlotsToCoeffs (:2094-2139); EvalLinearTransform (:1834-1882) is the same scheme with single hoisting.
//   digits      = EvalFastRotationPrecompute(ct)                                              (:1941)
//   fr[j]       = k_in[j] != 0 ? EvalFastRotationExt(ct, k_in[j], digits, addFirst = true)
//                              : KeySwitchExt(ct, addFirst = true)                            (:1943-1946)  [2][QlP][N]
//   for i < b:  inner = sum_j fr[j] (.) A[g i + j]   (EvalMultExt / EvalAddExtInPlace, j with g i + j == skip left out)
//               i == 0        : first = KeySwitchDownFirstElement(inner); outer = inner with element 0 zeroed
//               k_out[i] != 0 : inner' = KeySwitchDown(inner); first += auto_k(inner'_0);
//                               outer += EvalFastRotationExt(inner', k_out[i], Precompute(inner'), addFirst = false)
//               k_out[i] == 0 : first += KeySwitchDownFirstElement(inner); outer += inner with element 0 zeroed
//   out = KeySwitchDown(outer); out_0 += first                                               (:1984-1985)
// k_in / k_out are AUTOMORPHISM indices (FindAutomorphismIndex2nComplex of the rotation; 0 = no rotation); A holds the
// b*g pre-encoded diagonals in the extended basis, [b*g][sizeQl+sizeP][N] EVALUATION, device resident.
// Nothing leaves the device and nothing is allocated: `ws` is the caller's, bsgs_workspace_words() words.
size_t Hks::bsgs_workspace_words(uint32_t sizeQl, uint32_t g) const {
    const size_t N = ring->N, Ql = sizeQl, QlP = sizeQl + sizeP;
    //  digits, inner digits | coeff tmp, first, auto tmp, inner' (2) | fr (2g), inner, outer, rotated (2 each) | P rows (2), conv (2)
    return N * (2 * (size_t)dnum * QlP + 5 * Ql + (2 * (size_t)g + 6) * QlP + 2 * sizeP + 2 * Ql);
}

void Hks::bsgs_level(uint32_t sizeQl, uint32_t g, uint32_t b, const uint32_t* k_in, const uint32_t* k_out,
                     const EvalKey* const* evk_in, const EvalKey* const* evk_out, const uint64_t* A, uint32_t skip,
                     const uint64_t* ct, uint64_t* out, uint64_t* ws, cudaStream_t st) {
    Level& lv      = level(sizeQl);
    const Ring& r  = *ring;
    const size_t N = r.N, P = (size_t)sizeQl * N, E = (size_t)(sizeQl + sizeP) * N;
    OFHE_REQUIRE(g >= 1 && b >= 1, "bsgs: empty step counts");
    uint64_t* digits = ws;
    uint64_t* idig   = digits + (size_t)dnum * E;
    uint64_t* coeff  = idig + (size_t)dnum * E;
    uint64_t* first  = coeff + P;
    uint64_t* atmp   = first + P;
    uint64_t* innerD = atmp + P;           // [2][sizeQl][N]
    uint64_t* fr     = innerD + 2 * P;     // [g][2][QlP][N]
    uint64_t* inner  = fr + 2 * (size_t)g * E;
    uint64_t* outer  = inner + 2 * E;
    uint64_t* rot    = outer + 2 * E;
    uint64_t* p_tmp  = rot + 2 * E;
    uint64_t* conv   = p_tmp + 2 * (size_t)sizeP * N;
    // hoisting: one ModUp of c1 for all g baby-step rotations
    modup(sizeQl, ct + P, digits, coeff, st);
    for (uint32_t j = 0; j < g; ++j) {
        if (k_in[j] != 0)
            fast_rotation_ext(sizeQl, k_in[j], true, ct, digits, *evk_in[j], fr + 2 * (size_t)j * E, st);
        else
            keyswitch_ext(sizeQl, true, ct, fr + 2 * (size_t)j * E, st);
    }
    for (uint32_t i = 0; i < b; ++i) {
        const uint32_t G = g * i;
        for (uint32_t j = 0; j < g; ++j) {
            if (j != 0 && G + j == skip)
                continue;
            const uint64_t* a = A + (size_t)(G + j) * E;
            for (uint32_t e = 0; e < 2; ++e)   // EvalMultExt: both elements times the same plaintext (:2267-2276)
                launch_ew(r, j == 0 ? EW_MUL : EW_MAC, inner + e * E, fr + (2 * (size_t)j + e) * E, a, lv.map_QlP, st);
        }
        if (i == 0) {
            mod_down2(sizeQl, inner, innerD, nullptr, p_tmp, conv, st);  // only element 0 is kept here
            OFHE_CUDA(cudaMemcpyAsync(first, innerD, sizeof(uint64_t) * P, cudaMemcpyDeviceToDevice, st));
            OFHE_CUDA(cudaMemsetAsync(outer, 0, sizeof(uint64_t) * E, st));
            OFHE_CUDA(cudaMemcpyAsync(outer + E, inner + E, sizeof(uint64_t) * E, cudaMemcpyDeviceToDevice, st));
        }
        else if (k_out[i] != 0) {
            mod_down2(sizeQl, inner, innerD, nullptr, p_tmp, conv, st);                       // KeySwitchDown
            launch_automorphism(r, atmp, innerD, ring_automap(ring, k_out[i], st), sizeQl, st);
            launch_ew(r, EW_ADD, first, first, atmp, lv.map_Ql, st);
            modup(sizeQl, innerD + P, idig, coeff, st);                                        // second hoisting
            fast_rotation_ext(sizeQl, k_out[i], false, innerD, idig, *evk_out[i], rot, st);
            for (uint32_t e = 0; e < 2; ++e)
                launch_ew(r, EW_ADD, outer + e * E, outer + e * E, rot + e * E, lv.map_QlP, st);
        }
        else {
            mod_down2(sizeQl, inner, innerD, nullptr, p_tmp, conv, st);
            launch_ew(r, EW_ADD, first, first, innerD, lv.map_Ql, st);
            launch_ew(r, EW_ADD, outer + E, outer + E, inner + E, lv.map_QlP, st);           // element 0 contributes zero
        }
    }
    mod_down2(sizeQl, outer, out, nullptr, p_tmp, conv, st);
    launch_ew(r, EW_ADD, out, out, first, lv.map_Ql, st);
}

EvalKey::~EvalKey() {
    if (a_d)
        cudaFree(a_d);
    if (b_d)
        cudaFree(b_d);
}

}  // namespace ofhe
