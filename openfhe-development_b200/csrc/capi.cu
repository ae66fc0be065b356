// capi.cu -- the extern "C" boundary declared in include/ofhe_b200.h.  No C++ exception crosses it.
#include <cstring>
#include <string>

#include "../../include/ofhe_b200.h"
#include "hks.h"
#include "hostmath.h"
#include "bfv.h"
#include "cereal_io.h"
#include "shard.h"

using namespace ofhe;

struct ofhe_ctx {
    Context* c;
};
struct ofhe_ring {
    Ring* r;
    std::map<std::string, std::unique_ptr<BConv>> bconv_cache;
    std::mutex mu;
};
struct ofhe_hks {
    Hks* h;
    ofhe_ring* ring;
    // device staging buffers of the *_host entry points (allocated on first use)
    uint64_t* stage_d = nullptr;
    size_t stage_words = 0;
    // pipelined batch entry point: copy streams + per-slot events
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_comp = nullptr;
    cudaEvent_t ev_in[3] = {}, ev_comp[3] = {}, ev_out[3] = {};
};
struct ofhe_evk {
    EvalKey* k;
};
struct ofhe_shard {
    Shard* s;
};
struct ofhe_bconv {
    std::unique_ptr<BConv> t;
    Ring* r;
};

namespace {
thread_local std::string g_err;

template <typename F>
int guard(F&& f) {
    try {
        f();
        return 0;
    }
    catch (const std::exception& e) {
        g_err = e.what();
        return 1;
    }
    catch (...) {
        g_err = "unknown error";
        return 1;
    }
}
inline cudaStream_t S(void* s) {
    return static_cast<cudaStream_t>(s);
}
void use(const Ring* r) {
    OFHE_CUDA(cudaSetDevice(r->ctx->device));
}

__global__ void fill_uniform_kernel(RingDev r, uint64_t* dst, LimbMap lm, uint64_t seed) {
    const uint32_t limb = blockIdx.y;
    const uint64_t q    = r.mc[lm.idx[limb]].q;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < r.N; i += (size_t)gridDim.x * blockDim.x) {
        // splitmix64 of (seed, limb, i), then the high word of z*q: uniform on [0,q) up to 2^-64 bias
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (((uint64_t)limb << 32) + i + 1);
        z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        dst[(size_t)limb * r.N + i] = __umul64hi(z, q);
    }
}
}  // namespace

extern "C" {

const char* ofhe_last_error(void) {
    return g_err.c_str();
}
const char* ofhe_version(void) {
    return "ofhe_b200 0.1 (sm_100a)";
}

int ofhe_ctx_create(int device, ofhe_ctx** out) {
    return guard([&] {
        OFHE_REQUIRE(out, "null output pointer");
        *out = new ofhe_ctx{context_create(device)};
    });
}
void ofhe_ctx_destroy(ofhe_ctx* ctx) {
    if (ctx) {
        delete ctx->c;
        delete ctx;
    }
}
int ofhe_ctx_sm_count(const ofhe_ctx* ctx) {
    return ctx ? ctx->c->sm_count : 0;
}
int ofhe_malloc(ofhe_ctx* ctx, size_t bytes, void** dptr) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaMalloc(dptr, bytes));
    });
}
int ofhe_free(ofhe_ctx* ctx, void* dptr) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaFree(dptr));
    });
}
int ofhe_memcpy_h2d(ofhe_ctx* ctx, void* d, const void* h, size_t bytes, void* stream) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, S(stream)));
    });
}
int ofhe_memcpy_d2h(ofhe_ctx* ctx, void* h, const void* d, size_t bytes, void* stream) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, S(stream)));
    });
}
int ofhe_memcpy_d2d(ofhe_ctx* ctx, void* d, const void* s, size_t bytes, void* stream) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, S(stream)));
    });
}
int ofhe_memset0(ofhe_ctx* ctx, void* dptr, size_t bytes, void* stream) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaMemsetAsync(dptr, 0, bytes, S(stream)));
    });
}
int ofhe_stream_sync(ofhe_ctx* ctx, void* stream) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaStreamSynchronize(S(stream)));
    });
}
int ofhe_stream_create(ofhe_ctx* ctx, void** stream_out) {
    return guard([&] {
        OFHE_REQUIRE(ctx && stream_out, "null argument");
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        cudaStream_t st;
        OFHE_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        *stream_out = (void*)st;
    });
}
int ofhe_stream_destroy(ofhe_ctx* ctx, void* stream) {
    return guard([&] {
        OFHE_REQUIRE(ctx && stream, "null argument");
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaStreamDestroy(S(stream)));
    });
}

int ofhe_ring_create(ofhe_ctx* ctx, uint32_t N, uint32_t n_moduli, const uint64_t* moduli_h, const uint64_t* roots_h,
                     ofhe_ring** out) {
    return guard([&] {
        OFHE_REQUIRE(ctx && moduli_h && roots_h && out, "null argument");
        auto* h = new ofhe_ring();
        try {
            h->r = ring_create(ctx->c, N, n_moduli, moduli_h, roots_h);
        }
        catch (...) {
            delete h;
            throw;
        }
        *out = h;
    });
}
void ofhe_ring_destroy(ofhe_ring* ring) {
    if (ring) {
        ring->bconv_cache.clear();
        delete ring->r;
        delete ring;
    }
}
int ofhe_ring_twiddles_host(ofhe_ring* ring, uint32_t mi, uint64_t* T, uint64_t* Tp, uint64_t* Ti, uint64_t* Tip,
                            uint64_t* nInv, uint64_t* nInvPrecon) {
    return guard([&] {
        Ring* r = ring->r;
        use(r);
        OFHE_REQUIRE(mi < r->n_moduli, "modulus index out of range");
        std::vector<Twiddle> f(r->N), v(r->N);
        OFHE_CUDA(cudaMemcpy(f.data(), r->tw_fwd_d + (size_t)mi * r->N, sizeof(Twiddle) * r->N, cudaMemcpyDeviceToHost));
        OFHE_CUDA(cudaMemcpy(v.data(), r->tw_inv_d + (size_t)mi * r->N, sizeof(Twiddle) * r->N, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < r->N; ++i) {
            T[i]   = f[i].w;
            Tp[i]  = f[i].ws;
            Ti[i]  = v[i].w;
            Tip[i] = v[i].ws;
        }
        *nInv       = r->mc_h[mi].nInv;
        *nInvPrecon = r->mc_h[mi].nInvS;
    });
}

static LimbMap checked_map(const Ring* r, uint32_t n, const uint32_t* idx) {
    OFHE_REQUIRE(idx, "null modulus index list");
    for (uint32_t i = 0; i < n; ++i)
        OFHE_REQUIRE(idx[i] < r->n_moduli, "modulus index out of range");
    return make_map(idx, n);
}

int ofhe_ntt_fwd(ofhe_ring* ring, const uint64_t* in_d, uint64_t* out_d, uint32_t n_limbs, const uint32_t* mod_idx_h,
                 void* stream) {
    return guard([&] {
        use(ring->r);
        launch_ntt_fwd(*ring->r, in_d, out_d, checked_map(ring->r, n_limbs, mod_idx_h), S(stream));
    });
}
int ofhe_ntt_inv(ofhe_ring* ring, const uint64_t* in_d, uint64_t* out_d, uint32_t n_limbs, const uint32_t* mod_idx_h,
                 void* stream) {
    return guard([&] {
        use(ring->r);
        launch_ntt_inv(*ring->r, in_d, out_d, checked_map(ring->r, n_limbs, mod_idx_h), S(stream));
    });
}
int ofhe_ew_binop(ofhe_ring* ring, int op, uint64_t* dst_d, const uint64_t* a_d, const uint64_t* b_d, uint32_t n_limbs,
                  const uint32_t* mod_idx_h, void* stream) {
    return guard([&] {
        use(ring->r);
        OFHE_REQUIRE(op >= 0 && op <= 4, "unknown element-wise op");
        launch_ew(*ring->r, (EwOp)op, dst_d, a_d, b_d, checked_map(ring->r, n_limbs, mod_idx_h), S(stream));
    });
}
int ofhe_ew_mul_scalars(ofhe_ring* ring, uint64_t* dst_d, const uint64_t* a_d, const uint64_t* scalars_h,
                        uint32_t n_limbs, const uint32_t* mod_idx_h, void* stream) {
    return guard([&] {
        use(ring->r);
        LimbMap lm = checked_map(ring->r, n_limbs, mod_idx_h);
        ScalarVec sv;
        for (uint32_t i = 0; i < n_limbs; ++i) {
            uint64_t q = ring->r->moduli[mod_idx_h[i]];
            sv.s[i]    = scalars_h[i] % q;  // Times() reduces the scalar first (mubintvecnat.cpp:295-304)
            sv.ss[i]   = hm::shoup_const(sv.s[i], q);
        }
        launch_mul_scalars(*ring->r, dst_d, a_d, sv, lm, S(stream));
    });
}
// ---- BEHZ correction steps (f4): see include/ofhe_b200.h
static ScalarVec scalar_vec(const Ring* r, uint32_t n, const uint32_t* idx, const uint64_t* s_h) {
    ScalarVec sv;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t q = r->moduli[idx[i]];
        OFHE_REQUIRE(s_h[i] < q, "scalar is not reduced modulo its limb");
        sv.s[i]  = s_h[i];
        sv.ss[i] = hm::shoup_const(s_h[i], q);
    }
    return sv;
}
int ofhe_behz_mtilde(ofhe_ring* ring, uint32_t n_q, const uint32_t* q_idx_h, const uint64_t* in_d,
                     const uint64_t* mtildeQHatInvModq_h, const uint64_t* QHatModmtilde_h, uint64_t negQInvModmtilde,
                     uint64_t* r_out_d, void* stream) {
    return guard([&] {
        use(ring->r);
        OFHE_REQUIRE(in_d && mtildeQHatInvModq_h && QHatModmtilde_h && r_out_d, "null argument");
        LimbMap qm = checked_map(ring->r, n_q, q_idx_h);
        ScalarVec hat = scalar_vec(ring->r, n_q, q_idx_h, mtildeQHatInvModq_h), qhm;
        for (uint32_t i = 0; i < n_q; ++i)
            qhm.s[i] = QHatModmtilde_h[i], qhm.ss[i] = 0;
        launch_behz_mtilde(*ring->r, in_d, qm, hat, qhm, negQInvModmtilde, r_out_d, S(stream));
    });
}
int ofhe_behz_mtilde_fix(ofhe_ring* ring, uint32_t n_bsk, const uint32_t* bsk_idx_h, uint64_t* s_d, const uint64_t* r_d,
                         const uint64_t* QModbsk_h, const uint64_t* mtildeInvModbsk_h, void* stream) {
    return guard([&] {
        use(ring->r);
        OFHE_REQUIRE(s_d && r_d && QModbsk_h && mtildeInvModbsk_h, "null argument");
        LimbMap bm = checked_map(ring->r, n_bsk, bsk_idx_h);
        launch_behz_mtilde_fix(*ring->r, s_d, bm, r_d, scalar_vec(ring->r, n_bsk, bsk_idx_h, QModbsk_h),
                               scalar_vec(ring->r, n_bsk, bsk_idx_h, mtildeInvModbsk_h), S(stream));
    });
}
int ofhe_behz_sk_fix(ofhe_ring* ring, uint32_t n_q, const uint32_t* q_idx_h, uint64_t* q_d, const uint64_t* alpha_d,
                     uint64_t msk, const uint64_t* BModq_h, void* stream) {
    return guard([&] {
        use(ring->r);
        OFHE_REQUIRE(q_d && alpha_d && BModq_h, "null argument");
        LimbMap qm = checked_map(ring->r, n_q, q_idx_h);
        launch_behz_sk_fix(*ring->r, q_d, qm, alpha_d, msk, scalar_vec(ring->r, n_q, q_idx_h, BModq_h), S(stream));
    });
}
int ofhe_switch_modulus(ofhe_ring* ring, uint64_t* dst_d, const uint64_t* src_d, uint32_t from_idx, uint32_t to_idx,
                        void* stream) {
    return guard([&] {
        use(ring->r);
        OFHE_REQUIRE(from_idx < ring->r->n_moduli, "modulus index out of range");
        LimbMap lm = checked_map(ring->r, 1, &to_idx);
        ScalarVec sv;
        sv.s[0]  = 1;
        sv.ss[0] = hm::shoup_const(1, ring->r->moduli[to_idx]);
        launch_switch_modulus_scale(*ring->r, dst_d, src_d, from_idx, sv, lm, S(stream));
    });
}
int ofhe_automorphism(ofhe_ring* ring, uint64_t* dst_d, const uint64_t* src_d, uint32_t n_limbs, uint32_t k,
                      void* stream) {
    return guard([&] {
        use(ring->r);
        OFHE_REQUIRE((k & 1) && k < 2 * ring->r->N, "automorphism index must be odd and below 2N");
        OFHE_REQUIRE(dst_d != src_d, "automorphism cannot run in place");
        launch_automorphism(*ring->r, dst_d, src_d, ring_automap(ring->r, k, S(stream)), n_limbs, S(stream));
    });
}
int ofhe_automorphism_coeff(ofhe_ring* ring, uint64_t* dst_d, const uint64_t* src_d, uint32_t n_limbs,
                            const uint32_t* mod_idx_h, uint32_t k, void* stream) {
    return guard([&] {
        use(ring->r);
        OFHE_REQUIRE((k & 1) && k < 2 * ring->r->N, "automorphism index must be odd and below 2N");
        OFHE_REQUIRE(dst_d != src_d, "automorphism cannot run in place");
        launch_automorphism_coeff(*ring->r, dst_d, src_d, k, checked_map(ring->r, n_limbs, mod_idx_h), S(stream));
    });
}
int ofhe_automorphism_map_host(ofhe_ring* ring, uint32_t k, uint32_t* map_h) {
    return guard([&] {
        use(ring->r);
        const uint32_t* d = ring_automap(ring->r, k, nullptr);
        OFHE_CUDA(cudaMemcpy(map_h, d, sizeof(uint32_t) * ring->r->N, cudaMemcpyDeviceToHost));
    });
}

namespace {
std::unique_ptr<BConv> build_bconv(const Ring& r, uint32_t n_src, const uint32_t* src, uint32_t n_dst,
                                   const uint32_t* dst);
}

int ofhe_approx_switch_crt_basis(ofhe_ring* ring, uint32_t n_src, const uint32_t* src_idx_h, uint32_t n_dst,
                                 const uint32_t* dst_idx_h, const uint64_t* in_d, uint64_t* out_d, void* stream) {
    return guard([&] {
        use(ring->r);
        checked_map(ring->r, n_src, src_idx_h);
        checked_map(ring->r, n_dst, dst_idx_h);
        std::string key((const char*)src_idx_h, n_src * 4);
        key.push_back('|');
        key.append((const char*)dst_idx_h, n_dst * 4);
        BConv* t = nullptr;
        {
            std::lock_guard<std::mutex> g(ring->mu);
            auto it = ring->bconv_cache.find(key);
            if (it == ring->bconv_cache.end())
                it = ring->bconv_cache.emplace(key, build_bconv(*ring->r, n_src, src_idx_h, n_dst, dst_idx_h)).first;
            t = it->second.get();
        }
        launch_bconv(*ring->r, *t, in_d, out_d, 1, 0, 0, /*premultiplied=*/false, S(stream));
    });
}

static void upload(void** dptr, const void* h, size_t bytes) {
    OFHE_CUDA(cudaMalloc(dptr, bytes));
    h2d_blocking(*dptr, h, bytes);
}
static std::unique_ptr<BConv> bconv_from_tables(const Ring& r, uint32_t n_src, const uint32_t* src, const uint32_t* src_pos,
                                                const uint64_t* hatinv, uint32_t n_dst, const uint32_t* dst,
                                                const uint64_t* mat_h, bool dst_major, uint32_t mat_cols) {
    OFHE_REQUIRE(n_src >= 1 && n_src <= (uint32_t)kMaxLimbs && n_dst >= 1 && n_dst <= (uint32_t)kMaxLimbs, "base conversion: unsupported basis size");
    std::vector<uint64_t> hinv(2 * (size_t)n_src);
    std::vector<uint32_t> mat(2 * (size_t)n_src * n_dst);
    for (uint32_t i = 0; i < n_src; ++i) {
        const uint64_t q = r.moduli[src[i]];
        hinv[2 * i]      = hatinv ? hatinv[i] % q : 1;
        hinv[2 * i + 1]  = hm::shoup_const(hinv[2 * i], q);
        for (uint32_t j = 0; j < n_dst; ++j) {
            const uint64_t m = (dst_major ? mat_h[(size_t)j * mat_cols + i] : mat_h[(size_t)i * mat_cols + j]);
            OFHE_REQUIRE(m < (uint64_t(1) << 60), "base conversion: table entries must be below 2^60");
            mat[2 * ((size_t)i * n_dst + j)]     = (uint32_t)(m & ((1u << 30) - 1));
            mat[2 * ((size_t)i * n_dst + j) + 1] = (uint32_t)(m >> 30);
        }
    }
    auto t   = std::make_unique<BConv>();
    t->sizeQ = n_src;
    t->sizeP = n_dst;
    t->qmap  = make_map(src, n_src);
    if (src_pos)
        for (uint32_t i = 0; i < n_src; ++i)
            t->qmap.pos[i] = (uint16_t)src_pos[i];
    t->pmap = make_map(dst, n_dst);
    upload((void**)&t->qhatinv_d, hinv.data(), sizeof(uint64_t) * hinv.size());
    upload((void**)&t->mat_d, mat.data(), sizeof(uint32_t) * mat.size());
    bconv_finalize(r, *t);
    return t;
}

int ofhe_bconv_create(ofhe_ring* ring, uint32_t n_src, const uint32_t* src_idx_h, const uint64_t* QHatInvModq_h,
                      uint32_t n_dst, const uint32_t* dst_idx_h, const uint64_t* QHatModp_h, int table_is_dst_major,
                      const uint64_t* alphaQModp_h, const double* qInv_h, ofhe_bconv** out) {
    return guard([&] {
        OFHE_REQUIRE(ring && src_idx_h && dst_idx_h && QHatInvModq_h && QHatModp_h && out, "null argument");
        OFHE_REQUIRE((alphaQModp_h == nullptr) == (qInv_h == nullptr), "alphaQModp and qInv must be given together");
        Ring* r = ring->r;
        use(r);
        checked_map(r, n_src, src_idx_h);
        checked_map(r, n_dst, dst_idx_h);
        auto t = bconv_from_tables(*r, n_src, src_idx_h, nullptr, QHatInvModq_h, n_dst, dst_idx_h, QHatModp_h,
                                   table_is_dst_major != 0, table_is_dst_major ? n_src : n_dst);
        if (alphaQModp_h) {
            t->mode = 1;
            upload((void**)&t->dbl_d, qInv_h, sizeof(double) * n_src);
            upload((void**)&t->alpha_d, alphaQModp_h, sizeof(uint64_t) * (size_t)(n_src + 1) * n_dst);
        }
        *out = new ofhe_bconv{std::move(t), r};
    });
}
int ofhe_scale_round_create(ofhe_ring* ring, uint32_t sizeI, const uint32_t* in_idx_h, uint32_t sizeO,
                            const uint32_t* out_idx_h, int out_limbs_first, const uint64_t* tbl_h, const double* frac_h,
                            ofhe_bconv** out) {
    return guard([&] {
        OFHE_REQUIRE(ring && in_idx_h && out_idx_h && tbl_h && frac_h && out, "null argument");
        Ring* r = ring->r;
        use(r);
        checked_map(r, sizeI, in_idx_h);
        checked_map(r, sizeO, out_idx_h);
        const uint32_t inputIndex = out_limbs_first ? sizeO : 0, outputIndex = out_limbs_first ? 0 : sizeI;
        std::vector<uint32_t> src_pos(sizeI);
        for (uint32_t i = 0; i < sizeI; ++i)
            src_pos[i] = inputIndex + i;  // dcrtpoly-impl.h:1527-1534
        auto t  = bconv_from_tables(*r, sizeI, in_idx_h, src_pos.data(), nullptr, sizeO, out_idx_h, tbl_h, true, sizeI + 1);
        t->mode = 2;
        std::vector<uint64_t> extra(sizeO);
        for (uint32_t j = 0; j < sizeO; ++j) {
            extra[j]          = tbl_h[(size_t)j * (sizeI + 1) + sizeI];
            t->xomap.pos[j]   = (uint16_t)(outputIndex + j);
        }
        upload((void**)&t->dbl_d, frac_h, sizeof(double) * sizeI);
        upload((void**)&t->extra_d, extra.data(), sizeof(uint64_t) * sizeO);
        *out = new ofhe_bconv{std::move(t), r};
    });
}
int ofhe_bconv_apply(ofhe_bconv* t, const uint64_t* in_d, uint64_t* out_d, void* stream) {
    return guard([&] {
        use(t->r);
        launch_bconv(*t->r, *t->t, in_d, out_d, 1, 0, 0, /*premultiplied=*/false, S(stream));
    });
}
void ofhe_bconv_destroy(ofhe_bconv* t) {
    delete t;
}

// ------------------------------------------------------------------------------------------------ hybrid key switch
int ofhe_hks_create(ofhe_ring* ring, uint32_t sizeQ, uint32_t sizeP, uint32_t numPartQ, ofhe_hks** out) {
    return guard([&] {
        OFHE_REQUIRE(ring && out, "null argument");
        use(ring->r);
        auto* h = new ofhe_hks();
        try {
            h->h = new Hks(ring->r, sizeQ, sizeP, numPartQ);
        }
        catch (...) {
            delete h;
            throw;
        }
        h->ring = ring;
        *out    = h;
    });
}
void ofhe_hks_destroy(ofhe_hks* hks) {
    if (hks) {
        if (hks->stage_d)
            cudaFree(hks->stage_d);
        for (cudaStream_t st : {hks->s_h2d, hks->s_d2h, hks->s_comp})
            if (st)
                cudaStreamDestroy(st);
        for (int i = 0; i < 3; ++i)
            for (cudaEvent_t e : {hks->ev_in[i], hks->ev_comp[i], hks->ev_out[i]})
                if (e)
                    cudaEventDestroy(e);
        delete hks->h;
        delete hks;
    }
}
uint32_t ofhe_hks_num_per_part(const ofhe_hks* hks) {
    return hks->h->alpha;
}
uint32_t ofhe_hks_num_digits(const ofhe_hks* hks, uint32_t sizeQl) {
    return std::min((sizeQl + hks->h->alpha - 1) / hks->h->alpha, hks->h->dnum);
}

int ofhe_hks_table_host(ofhe_hks* hks, int which, uint32_t a, uint32_t b, uint64_t* out, size_t* n_words) {
    return guard([&] {
        Hks& h        = *hks->h;
        const Ring& r = *h.ring;
        const uint64_t* Q = r.moduli.data();
        const uint64_t* P = Q + h.sizeQ;
        std::vector<uint64_t> v;
        switch (which) {
            case 0: v = h.PInvModq; break;
            case 1: v = h.PModq; break;
            case 2:
                for (uint32_t j = 0; j < h.sizeP; ++j)
                    v.push_back(hm::inv(hm::prod_mod(P, h.sizeP, j, P[j]), P[j]));
                break;
            case 3:
                for (uint32_t j = 0; j < h.sizeP; ++j)
                    for (uint32_t i = 0; i < h.sizeQ; ++i)
                        v.push_back(hm::prod_mod(P, h.sizeP, j, Q[i]));
                break;
            case 4: {  // digit a truncated to b limbs
                OFHE_REQUIRE(a < h.dnum && b >= 1 && a * h.alpha + b <= h.sizeQ, "bad digit");
                const uint64_t* d = Q + a * h.alpha;
                for (uint32_t i = 0; i < b; ++i)
                    v.push_back(hm::inv(hm::prod_mod(d, b, i, d[i]), d[i]));
                break;
            }
            case 5: {  // sizeQl = a, part = b: [digit_size][sizeCompl], complementary basis order
                OFHE_REQUIRE(a >= 1 && a <= h.sizeQ && b * h.alpha < a, "bad level / digit");
                uint32_t s = b * h.alpha, ds = std::min(h.alpha, a - s);
                for (uint32_t i = 0; i < ds; ++i) {
                    for (uint32_t k = 0; k < a; ++k)
                        if (k < s || k >= s + ds)
                            v.push_back(hm::prod_mod(Q + s, ds, i, Q[k]));
                    for (uint32_t j = 0; j < h.sizeP; ++j)
                        v.push_back(hm::prod_mod(Q + s, ds, i, P[j]));
                }
                break;
            }
            case 6:
            case 7: {
                OFHE_REQUIRE(a >= 2 && a <= h.sizeQ, "bad level");
                Hks::Level& lv = h.level(a);
                for (uint32_t i = 0; i + 1 < a; ++i)
                    v.push_back(which == 6 ? lv.rescaleA.s[i] : lv.rescaleB.s[i]);
                break;
            }
            default: throw Error("unknown table id");
        }
        if (out)
            std::memcpy(out, v.data(), v.size() * sizeof(uint64_t));
        if (n_words)
            *n_words = v.size();
    });
}

int ofhe_evk_create(ofhe_hks* hks, const uint64_t* a_h, const uint64_t* b_h, ofhe_evk** out) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        size_t words = (size_t)h.dnum * (h.sizeQ + h.sizeP) * h.ring->N;
        auto k       = std::make_unique<EvalKey>();
        k->hks       = &h;
        k->dnum      = h.dnum;
        OFHE_CUDA(cudaMalloc(&k->a_d, words * sizeof(uint64_t)));
        OFHE_CUDA(cudaMalloc(&k->b_d, words * sizeof(uint64_t)));
        h2d_blocking(k->a_d, a_h, words * sizeof(uint64_t));
        h2d_blocking(k->b_d, b_h, words * sizeof(uint64_t));
        *out = new ofhe_evk{k.release()};
    });
}
int ofhe_evk_create_random(ofhe_hks* hks, uint64_t seed, ofhe_evk** out) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        const uint32_t qp = h.sizeQ + h.sizeP;
        size_t words      = (size_t)h.dnum * qp * h.ring->N;
        auto k            = std::make_unique<EvalKey>();
        k->hks            = &h;
        k->dnum           = h.dnum;
        OFHE_CUDA(cudaMalloc(&k->a_d, words * sizeof(uint64_t)));
        OFHE_CUDA(cudaMalloc(&k->b_d, words * sizeof(uint64_t)));
        std::vector<uint32_t> idx(qp);
        for (uint32_t i = 0; i < qp; ++i)
            idx[i] = i;
        LimbMap lm = make_map(idx.data(), qp);
        for (uint32_t j = 0; j < h.dnum; ++j) {
            dim3 g(64, qp);
            fill_uniform_kernel<<<g, 256>>>(h.ring->dev(), k->a_d + (size_t)j * qp * h.ring->N, lm, seed + 2 * j);
            fill_uniform_kernel<<<g, 256>>>(h.ring->dev(), k->b_d + (size_t)j * qp * h.ring->N, lm, seed + 2 * j + 1);
        }
        OFHE_CUDA(cudaGetLastError());
        OFHE_CUDA(cudaDeviceSynchronize());
        *out = new ofhe_evk{k.release()};
    });
}
int ofhe_evk_copy_to_device(const ofhe_evk* evk, int which, uint64_t* dst_d, void* stream) {
    return guard([&] {
        OFHE_REQUIRE(evk && dst_d && (which == 0 || which == 1), "evk copy: bad arguments");
        const EvalKey& k = *evk->k;
        use(k.hks->ring);
        const size_t words = (size_t)k.dnum * (k.hks->sizeQ + k.hks->sizeP) * k.hks->ring->N;
        OFHE_CUDA(cudaMemcpyAsync(dst_d, which == 0 ? k.a_d : k.b_d, words * sizeof(uint64_t), cudaMemcpyDeviceToDevice,
                                  S(stream)));
    });
}
void ofhe_evk_destroy(ofhe_evk* evk) {
    if (evk) {
        delete evk->k;
        delete evk;
    }
}

int ofhe_ntt_cluster_occupancy(ofhe_ctx* ctx, uint32_t logN, int* fwd_inv2) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        ntt_cluster_occupancy(logN, fwd_inv2);
    });
}
int ofhe_hks_reserve(ofhe_hks* hks, uint32_t sizeQl, void* stream) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        h.level(sizeQl);
        if (sizeQl >= 2)
            h.level(sizeQl - 1);  // rescale reads the next level's limb map
        h.ring->ctx->get_scratch(sizeof(uint64_t) * h.max_workspace_words(sizeQl), S(stream));
    });
}
int ofhe_hks_modup(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* c_d, uint64_t* digits_d, void* stream) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        uint64_t* tmp = static_cast<uint64_t*>(h.ring->ctx->get_scratch(sizeof(uint64_t) * h.workspace_words(sizeQl), S(stream)));
        h.modup(sizeQl, c_d, digits_d, tmp, S(stream));
    });
}
int ofhe_hks_inner_product(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* digits_d, const ofhe_evk* evk,
                           uint64_t* e01_d, void* stream) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        h.inner_product(sizeQl, ofhe_hks_num_digits(hks, sizeQl), digits_d, *evk->k, e01_d, S(stream));
    });
}
int ofhe_hks_approx_mod_down(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* in_d, uint64_t* out_d, uint32_t n_polys,
                             void* stream) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        OFHE_REQUIRE(n_polys == 1 || n_polys == 2, "n_polys must be 1 or 2");
        const size_t N = h.ring->N, sizeQlP = sizeQl + h.sizeP;
        uint64_t* ws = static_cast<uint64_t*>(h.ring->ctx->get_scratch(sizeof(uint64_t) * h.workspace_words(sizeQl), S(stream)));
        if (n_polys == 2) {
            h.mod_down2(sizeQl, in_d, out_d, nullptr, ws, ws + 2 * (size_t)h.sizeP * N, S(stream));
        }
        else {
            // run the 2-polynomial pipeline on a duplicated input and keep the first result
            uint64_t* e01  = ws;
            uint64_t* o01  = e01 + 2 * sizeQlP * N;
            uint64_t* ptmp = o01 + 2 * (size_t)sizeQl * N;
            uint64_t* conv = ptmp + 2 * (size_t)h.sizeP * N;
            for (int b = 0; b < 2; ++b)
                OFHE_CUDA(cudaMemcpyAsync(e01 + b * sizeQlP * N, in_d, sizeof(uint64_t) * sizeQlP * N,
                                          cudaMemcpyDeviceToDevice, S(stream)));
            h.mod_down2(sizeQl, e01, o01, nullptr, ptmp, conv, S(stream));
            OFHE_CUDA(cudaMemcpyAsync(out_d, o01, sizeof(uint64_t) * sizeQl * N, cudaMemcpyDeviceToDevice, S(stream)));
        }
    });
}
int ofhe_hks_keyswitch_core(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* c_d, const ofhe_evk* evk,
                            uint64_t* out01_d, void* stream) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        uint64_t* ws = static_cast<uint64_t*>(h.ring->ctx->get_scratch(sizeof(uint64_t) * h.workspace_words(sizeQl), S(stream)));
        h.keyswitch_core(sizeQl, c_d, *evk->k, out01_d, nullptr, ws, S(stream));
    });
}
int ofhe_ckks_evalmult(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* ct1_d, const uint64_t* ct2_d,
                       const ofhe_evk* evk, uint64_t* out_d, void* stream) {
    return guard([&] {
        use(hks->h->ring);
        hks->h->evalmult(sizeQl, ct1_d, ct2_d, *evk->k, out_d, S(stream));
    });
}
int ofhe_ckks_evalmult_host(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* ct1_h, const uint64_t* ct2_h,
                            const ofhe_evk* evk, uint64_t* out_h) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        const size_t words = 2 * (size_t)sizeQl * h.ring->N;
        if (hks->stage_words < 3 * words) {
            if (hks->stage_d) {
                OFHE_CUDA(cudaDeviceSynchronize());
                OFHE_CUDA(cudaFree(hks->stage_d));
                hks->stage_d = nullptr;
            }
            OFHE_CUDA(cudaMalloc(&hks->stage_d, 3 * words * sizeof(uint64_t)));
            hks->stage_words = 3 * words;
        }
        uint64_t *d1 = hks->stage_d, *d2 = d1 + words, *d3 = d2 + words;
        cudaStream_t st = nullptr;
        OFHE_CUDA(cudaMemcpyAsync(d1, ct1_h, words * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        OFHE_CUDA(cudaMemcpyAsync(d2, ct2_h, words * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        h.evalmult(sizeQl, d1, d2, *evk->k, d3, st);
        OFHE_CUDA(cudaMemcpyAsync(out_h, d3, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        OFHE_CUDA(cudaStreamSynchronize(st));
    });
}
int ofhe_ckks_evalmult_host_batch(ofhe_hks* hks, uint32_t sizeQl, uint32_t n, const uint64_t* const* ct1_h,
                                  const uint64_t* const* ct2_h, const ofhe_evk* evk, uint64_t* const* out_h) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        constexpr int S_   = 3;
        const size_t words = 2 * (size_t)sizeQl * h.ring->N, bytes = words * sizeof(uint64_t);
        if (hks->stage_words < 3 * S_ * words) {
            if (hks->stage_d) {
                OFHE_CUDA(cudaDeviceSynchronize());
                OFHE_CUDA(cudaFree(hks->stage_d));
                hks->stage_d = nullptr;
            }
            OFHE_CUDA(cudaMalloc(&hks->stage_d, 3 * S_ * words * sizeof(uint64_t)));
            hks->stage_words = 3 * S_ * words;
        }
        if (!hks->s_comp) {
            OFHE_CUDA(cudaStreamCreateWithFlags(&hks->s_h2d, cudaStreamNonBlocking));
            OFHE_CUDA(cudaStreamCreateWithFlags(&hks->s_d2h, cudaStreamNonBlocking));
            OFHE_CUDA(cudaStreamCreateWithFlags(&hks->s_comp, cudaStreamNonBlocking));
            for (int i = 0; i < S_; ++i) {
                OFHE_CUDA(cudaEventCreateWithFlags(&hks->ev_in[i], cudaEventDisableTiming));
                OFHE_CUDA(cudaEventCreateWithFlags(&hks->ev_comp[i], cudaEventDisableTiming));
                OFHE_CUDA(cudaEventCreateWithFlags(&hks->ev_out[i], cudaEventDisableTiming));
            }
        }
        OFHE_CUDA(cudaDeviceSynchronize());  // the staging buffers / scratch may be in use by earlier stream work
        for (uint32_t k = 0; k < n; ++k) {
            const int sl = k % S_;
            uint64_t *d1 = hks->stage_d + (size_t)sl * 3 * words, *d2 = d1 + words, *d3 = d2 + words;
            if (k >= (uint32_t)S_)  // slot inputs are free once the multiplication that used them has run
                OFHE_CUDA(cudaStreamWaitEvent(hks->s_h2d, hks->ev_comp[sl], 0));
            OFHE_CUDA(cudaMemcpyAsync(d1, ct1_h[k], bytes, cudaMemcpyHostToDevice, hks->s_h2d));
            OFHE_CUDA(cudaMemcpyAsync(d2, ct2_h[k], bytes, cudaMemcpyHostToDevice, hks->s_h2d));
            OFHE_CUDA(cudaEventRecord(hks->ev_in[sl], hks->s_h2d));
            OFHE_CUDA(cudaStreamWaitEvent(hks->s_comp, hks->ev_in[sl], 0));
            if (k >= (uint32_t)S_)  // the slot's previous result must have left before it is overwritten
                OFHE_CUDA(cudaStreamWaitEvent(hks->s_comp, hks->ev_out[sl], 0));
            h.evalmult(sizeQl, d1, d2, *evk->k, d3, hks->s_comp);
            OFHE_CUDA(cudaEventRecord(hks->ev_comp[sl], hks->s_comp));
            OFHE_CUDA(cudaStreamWaitEvent(hks->s_d2h, hks->ev_comp[sl], 0));
            OFHE_CUDA(cudaMemcpyAsync(out_h[k], d3, bytes, cudaMemcpyDeviceToHost, hks->s_d2h));
            OFHE_CUDA(cudaEventRecord(hks->ev_out[sl], hks->s_d2h));
        }
        OFHE_CUDA(cudaStreamSynchronize(hks->s_d2h));
        OFHE_CUDA(cudaStreamSynchronize(hks->s_comp));
    });
}
int ofhe_ckks_rescale(ofhe_hks* hks, uint32_t sizeQl, const uint64_t* in_d, uint64_t* out_d, uint32_t n_polys,
                      void* stream) {
    return guard([&] {
        use(hks->h->ring);
        hks->h->rescale(sizeQl, in_d, out_d, n_polys, S(stream));
    });
}
int ofhe_ckks_rotate(ofhe_hks* hks, uint32_t sizeQl, uint32_t k, const uint64_t* ct_d, const ofhe_evk* evk,
                     uint64_t* out_d, void* stream) {
    return guard([&] {
        use(hks->h->ring);
        OFHE_REQUIRE((k & 1) && k < 2 * hks->h->ring->N, "automorphism index must be odd and below 2N");
        hks->h->rotate(sizeQl, k, ct_d, *evk->k, out_d, S(stream));
    });
}
int ofhe_ckks_fast_rotation(ofhe_hks* hks, uint32_t sizeQl, uint32_t k, const uint64_t* ct_d, const uint64_t* digits_d,
                            const ofhe_evk* evk, uint64_t* out_d, void* stream) {
    return guard([&] {
        use(hks->h->ring);
        OFHE_REQUIRE((k & 1) && k < 2 * hks->h->ring->N, "automorphism index must be odd and below 2N");
        hks->h->fast_rotation(sizeQl, k, ct_d, digits_d, *evk->k, out_d, S(stream));
    });
}
int ofhe_ckks_fast_rotation_ext(ofhe_hks* hks, uint32_t sizeQl, uint32_t k, int add_first, const uint64_t* ct_d,
                                const uint64_t* digits_d, const ofhe_evk* evk, uint64_t* out_ext_d, void* stream) {
    return guard([&] {
        use(hks->h->ring);
        OFHE_REQUIRE((k & 1) && k < 2 * hks->h->ring->N, "automorphism index must be odd and below 2N");
        hks->h->fast_rotation_ext(sizeQl, k, add_first != 0, ct_d, digits_d, *evk->k, out_ext_d, S(stream));
    });
}
int ofhe_hks_keyswitch_ext(ofhe_hks* hks, uint32_t sizeQl, int add_first, const uint64_t* ct_d, uint64_t* out_ext_d,
                           void* stream) {
    return guard([&] {
        use(hks->h->ring);
        hks->h->keyswitch_ext(sizeQl, add_first != 0, ct_d, out_ext_d, S(stream));
    });
}
size_t ofhe_ckks_bsgs_workspace_words(const ofhe_hks* hks, uint32_t sizeQl, uint32_t g) {
    return hks->h->bsgs_workspace_words(sizeQl, g);
}
int ofhe_ckks_bsgs_level(ofhe_hks* hks, uint32_t sizeQl, uint32_t g, uint32_t b, const uint32_t* k_in_h,
                         const uint32_t* k_out_h, const ofhe_evk* const* evk_in_h, const ofhe_evk* const* evk_out_h,
                         const uint64_t* A_d, uint32_t skip, const uint64_t* ct_d, uint64_t* out_d, uint64_t* ws_d,
                         void* stream) {
    return guard([&] {
        Hks& h = *hks->h;
        use(h.ring);
        OFHE_REQUIRE(k_in_h && k_out_h && evk_in_h && evk_out_h && A_d && ct_d && out_d && ws_d, "null argument");
        OFHE_REQUIRE(g >= 1 && g <= 64 && b >= 1 && b <= 64, "bsgs: step counts out of range");
        std::vector<const EvalKey*> ki(g, nullptr), ko(b, nullptr);
        for (uint32_t j = 0; j < g; ++j) {
            OFHE_REQUIRE(k_in_h[j] == 0 || ((k_in_h[j] & 1) && k_in_h[j] < 2 * h.ring->N && evk_in_h[j]),
                         "bsgs: baby-step automorphism index must be 0 or odd and below 2N, with a key");
            ki[j] = k_in_h[j] ? evk_in_h[j]->k : nullptr;
        }
        for (uint32_t i = 0; i < b; ++i) {
            OFHE_REQUIRE(k_out_h[i] == 0 || ((k_out_h[i] & 1) && k_out_h[i] < 2 * h.ring->N && evk_out_h[i]),
                         "bsgs: giant-step automorphism index must be 0 or odd and below 2N, with a key");
            ko[i] = k_out_h[i] ? evk_out_h[i]->k : nullptr;
        }
        h.bsgs_level(sizeQl, g, b, k_in_h, k_out_h, ki.data(), ko.data(), A_d, skip, ct_d, out_d, ws_d, S(stream));
    });
}
uint32_t ofhe_find_automorphism_index(int32_t index, uint32_t m) {
    // FindAutomorphismIndex2nComplex (core/lib/math/nbtheory2.cpp:243-262): 5^index mod m (5^-1 for negative index)
    if (index == 0)
        return 1;
    if (index == (int32_t)m - 1)
        return (uint32_t)index;
    uint64_t g0 = 5;
    if (index < 0) {
        // inverse of 5 modulo the power of two m
        uint64_t x = 1;
        for (int i = 0; i < 6; ++i)
            x = x * (2 - 5 * x);
        g0 = x & (m - 1);
    }
    uint64_t g  = g0;
    uint32_t n  = (uint32_t)(index < 0 ? -index : index);
    for (uint32_t j = 1; j < n; ++j)
        g = (g * g0) & (m - 1);
    return (uint32_t)g;
}

int ofhe_ckks_paramgen(uint32_t N, uint32_t depth, uint32_t scale_bits, uint32_t first_bits, uint32_t dnum,
                       int scal_tech, uint64_t* moduli_out, uint64_t* roots_out, uint32_t* shape_out) {
    return guard([&] {
        OFHE_REQUIRE(scale_bits <= 60 && first_bits <= 60 && scale_bits >= 14, "modulus sizes must be in [14, 60] bits");
        hm::CkksChain c = hm::ckks_chain(N, depth, scale_bits, first_bits, dnum, scal_tech);
        std::memcpy(moduli_out, c.moduli.data(), c.moduli.size() * sizeof(uint64_t));
        std::memcpy(roots_out, c.roots.data(), c.roots.size() * sizeof(uint64_t));
        shape_out[0] = c.sizeQ;
        shape_out[1] = c.sizeP;
        shape_out[2] = c.dnum;
        shape_out[3] = c.alpha;
    });
}
uint64_t ofhe_root_of_unity(uint64_t m, uint64_t q) {
    uint64_t r = 0;
    guard([&] { r = hm::minimal_root_of_unity(m, q); });
    return r;
}
uint64_t ofhe_first_prime(uint32_t bits, uint64_t m) {
    return hm::first_prime(bits, m);
}
uint64_t ofhe_last_prime(uint32_t bits, uint64_t m) {
    uint64_t r = 0;
    guard([&] { r = hm::last_prime(bits, m); });
    return r;
}
uint64_t ofhe_previous_prime(uint64_t q, uint64_t m) {
    return hm::previous_prime(q, m);
}
uint64_t ofhe_next_prime(uint64_t q, uint64_t m) {
    return hm::next_prime(q, m);
}

int ofhe_shard_create(ofhe_ring* ring, uint32_t sizeQ, uint32_t sizeP, uint32_t numPartQ, uint32_t rank, uint32_t world,
                      ofhe_shard** out) {
    return guard([&] {
        OFHE_REQUIRE(ring && out, "null argument");
        use(ring->r);
        *out = new ofhe_shard{new Shard(ring->r, sizeQ, sizeP, numPartQ, rank, world)};
    });
}
void ofhe_shard_destroy(ofhe_shard* s) {
    if (s) {
        delete s->s;
        delete s;
    }
}
int ofhe_shard_info(ofhe_shard* s, uint32_t sizeQl, uint32_t* o) {
    return guard([&] {
        use(s->s->ring);
        Shard::Level& lv = s->s->level(sizeQl);
        o[0] = lv.nOwnQ;
        o[1] = (uint32_t)s->s->ownP.size();
        o[2] = lv.slotsQ;
        o[3] = s->s->slotsP;
        o[4] = s->s->nOwnQFull + (uint32_t)s->s->ownP.size();
        o[5] = lv.n_digits;
    });
}
int ofhe_shard_owned(ofhe_shard* s, uint32_t sizeQl, uint32_t* q, uint32_t* p) {
    return guard([&] {
        use(s->s->ring);
        Shard::Level& lv = s->s->level(sizeQl);
        for (size_t t = 0; t < lv.ownQ.size(); ++t)
            q[t] = lv.ownQ[t];
        for (size_t t = 0; t < s->s->ownP.size(); ++t)
            p[t] = s->s->ownP[t];
    });
}
int ofhe_shard_evalmult_a(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct1_d, const uint64_t* ct2_d, uint64_t* out_d,
                          uint64_t* gatherQ_d, void* stream) {
    return guard([&] {
        use(s->s->ring);
        s->s->evalmult_a(sizeQl, ct1_d, ct2_d, out_d, gatherQ_d, S(stream));
    });
}
int ofhe_shard_evalmult_b(ofhe_shard* s, uint32_t sizeQl, const uint64_t* gatherQ_d, const uint64_t* evk_a_d,
                          const uint64_t* evk_b_d, uint64_t* gatherP_d, void* stream) {
    return guard([&] {
        use(s->s->ring);
        s->s->evalmult_b(sizeQl, gatherQ_d, evk_a_d, evk_b_d, gatherP_d, S(stream));
    });
}
int ofhe_shard_evalmult_c(ofhe_shard* s, uint32_t sizeQl, const uint64_t* gatherP_d, uint64_t* out_d, void* stream) {
    return guard([&] {
        use(s->s->ring);
        s->s->evalmult_c(sizeQl, gatherP_d, out_d, S(stream));
    });
}

namespace {
PeerOut peer_list(const Shard& sh, uint64_t* const* of_rank) {
    if (sh.world > kMaxPeers)
        throw Error("peer push supports at most 8 ranks");
    PeerOut po;
    po.n = sh.world;
    for (uint32_t r = 0; r < sh.world; ++r) {
        if (!of_rank[r])
            throw Error("null peer buffer");
        po.base[r] = of_rank[r];
    }
    return po;
}
struct PeerFlags {
    uint64_t* base[kMaxPeers];
};
// flags layout per rank: [n_barriers][world + 2] = arrivals from each rank | this rank's epoch | timeout count
__global__ void peer_barrier_kernel(PeerFlags pf, uint32_t world, uint32_t rank, uint32_t index) {
    const uint32_t row = index * (world + 2);
    uint64_t* mine     = pf.base[rank] + row;
    __shared__ uint64_t epoch_s;
    if (threadIdx.x == 0) {
        epoch_s     = mine[world] + 1;
        mine[world] = epoch_s;
    }
    __syncthreads();
    const uint64_t epoch = epoch_s;
    if (threadIdx.x < world) {
        // everything this rank stored to its peers (previous kernels on this stream) is ordered before the flag
        __threadfence_system();
        uint64_t* theirs = pf.base[threadIdx.x] + row + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
        const uint64_t* arrive = mine + threadIdx.x;
        uint64_t t0, now, seen = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(arrive) : "memory");
            if (seen >= epoch)
                break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 10000000000ull) {  // 10 s of wall clock: a peer is gone
                // count it, then abort the context: the phases behind this barrier would consume partially written
                // gather buffers, so the failure must be LOUD (every later CUDA call of this rank returns an error)
                // rather than a silently wrong ciphertext
                atomicAdd(reinterpret_cast<unsigned long long*>(mine + world + 1), 1ULL);
                __threadfence_system();
                __trap();
            }
            __nanosleep(64);
        }
    }
}
}  // namespace

int ofhe_shard_evalmult_a_push(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct1_d, const uint64_t* ct2_d,
                               uint64_t* out_d, uint64_t* const* gatherQ_of_rank_h, void* stream) {
    return guard([&] {
        use(s->s->ring);
        const PeerOut po = peer_list(*s->s, gatherQ_of_rank_h);
        s->s->evalmult_a(sizeQl, ct1_d, ct2_d, out_d, po.base[s->s->rank], S(stream), &po);
    });
}
int ofhe_shard_evalmult_b_push(ofhe_shard* s, uint32_t sizeQl, const uint64_t* evk_a_d, const uint64_t* evk_b_d,
                               uint64_t* const* gatherQ_of_rank_h, uint64_t* const* gatherP_of_rank_h, void* stream) {
    return guard([&] {
        use(s->s->ring);
        const PeerOut po = peer_list(*s->s, gatherP_of_rank_h);
        s->s->evalmult_b(sizeQl, gatherQ_of_rank_h[s->s->rank], evk_a_d, evk_b_d, po.base[s->s->rank], S(stream), &po);
    });
}
int ofhe_shard_rotate_a(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, uint64_t* tmp_d, uint64_t* gatherQ_d,
                        void* stream) {
    return guard([&] {
        use(s->s->ring);
        s->s->rotate_a(sizeQl, ct_d, tmp_d, gatherQ_d, S(stream));
    });
}
int ofhe_shard_rotate_a_push(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, uint64_t* tmp_d,
                             uint64_t* const* gatherQ_of_rank_h, void* stream) {
    return guard([&] {
        use(s->s->ring);
        const PeerOut po = peer_list(*s->s, gatherQ_of_rank_h);
        s->s->rotate_a(sizeQl, ct_d, tmp_d, po.base[s->s->rank], S(stream), &po);
    });
}
int ofhe_shard_rotate_c(ofhe_shard* s, uint32_t sizeQl, uint32_t k, const uint64_t* gatherP_d, uint64_t* tmp_d,
                        uint64_t* out_d, void* stream) {
    return guard([&] {
        use(s->s->ring);
        OFHE_REQUIRE((k & 1u) && k < 2 * s->s->ring->N, "automorphism index must be odd and below 2N");
        s->s->rotate_c(sizeQl, k, gatherP_d, tmp_d, out_d, S(stream));
    });
}
int ofhe_shard_rescale_a(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, uint64_t* gatherL_d, void* stream) {
    return guard([&] {
        use(s->s->ring);
        s->s->rescale_a(sizeQl, ct_d, gatherL_d, S(stream));
    });
}
int ofhe_shard_rescale_a_push(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, uint64_t* const* gatherL_of_rank_h,
                              void* stream) {
    return guard([&] {
        use(s->s->ring);
        const PeerOut po = peer_list(*s->s, gatherL_of_rank_h);
        s->s->rescale_a(sizeQl, ct_d, po.base[s->s->rank], S(stream), &po);
    });
}
int ofhe_shard_rescale_b(ofhe_shard* s, uint32_t sizeQl, const uint64_t* ct_d, const uint64_t* gatherL_d, uint64_t* out_d,
                         void* stream) {
    return guard([&] {
        use(s->s->ring);
        s->s->rescale_b(sizeQl, ct_d, gatherL_d, out_d, S(stream));
    });
}
size_t ofhe_peer_flags_words(uint32_t world, uint32_t n_barriers) {
    return (size_t)n_barriers * (world + 2);
}
int ofhe_peer_barrier(ofhe_ctx* ctx, uint64_t* const* flags_of_rank_h, uint32_t world, uint32_t rank, uint32_t index,
                      void* stream) {
    return guard([&] {
        if (world == 0 || world > kMaxPeers || rank >= world)
            throw Error("peer barrier: bad world / rank");
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        PeerFlags pf;
        for (uint32_t r = 0; r < world; ++r)
            pf.base[r] = flags_of_rank_h[r];
        peer_barrier_kernel<<<1, 32, 0, S(stream)>>>(pf, world, rank, index);
        OFHE_CUDA(cudaGetLastError());
    });
}
int ofhe_peer_barrier_timeouts(ofhe_ctx* ctx, const uint64_t* flags_d, uint32_t world, uint32_t n_barriers,
                               uint64_t* count_out) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        std::vector<uint64_t> h(ofhe_peer_flags_words(world, n_barriers));
        OFHE_CUDA(cudaMemcpy(h.data(), flags_d, sizeof(uint64_t) * h.size(), cudaMemcpyDeviceToHost));
        uint64_t n = 0;
        for (uint32_t b = 0; b < n_barriers; ++b)
            n += h[(size_t)b * (world + 2) + world + 1];
        *count_out = n;
    });
}
int ofhe_ipc_export(ofhe_ctx* ctx, void* dptr, void* handle64_out) {
    return guard([&] {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        cudaIpcMemHandle_t h;
        OFHE_CUDA(cudaIpcGetMemHandle(&h, dptr));
        std::memcpy(handle64_out, &h, sizeof(h));
    });
}
int ofhe_ipc_import(ofhe_ctx* ctx, const void* handle64, void** dptr_out) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, sizeof(h));
        OFHE_CUDA(cudaIpcOpenMemHandle(dptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    });
}
int ofhe_ipc_release(ofhe_ctx* ctx, void* dptr) {
    return guard([&] {
        OFHE_CUDA(cudaSetDevice(ctx->c->device));
        OFHE_CUDA(cudaIpcCloseMemHandle(dptr));
    });
}

struct ofhe_bfv {
    BfvMult* b;
};
int ofhe_bfv_tables_host(uint32_t sizeQ, const uint64_t* moduliQ_h, uint32_t sizeR, const uint64_t* moduliR_h, uint64_t t,
                         int which, void* out_h, size_t* count_out) {
    return guard([&] {
        OFHE_REQUIRE(sizeQ >= 1 && sizeR >= 1 && sizeQ <= 64 && sizeR <= 64, "BFV basis sizes out of range");
        for (uint32_t i = 0; i < sizeQ; ++i)
            for (uint32_t j = 0; j < sizeR; ++j)
                OFHE_REQUIRE(moduliQ_h[i] != moduliR_h[j], "Q and R must be disjoint");
        const size_t n = hm::bfv_table(moduliQ_h, sizeQ, moduliR_h, sizeR, t, which, out_h);
        OFHE_REQUIRE(n != 0, "unknown BFV table");
        if (count_out)
            *count_out = n;
    });
}
int ofhe_bfv_create(ofhe_ring* ring, uint32_t sizeQ, uint32_t sizeR, const ofhe_bconv* q_to_r_exact,
                    const ofhe_bconv* q_to_r_approx, const ofhe_bconv* r_to_q_exact, const ofhe_bconv* scale_round,
                    ofhe_bfv** out) {
    return guard([&] {
        use(ring->r);
        OFHE_REQUIRE(q_to_r_exact && q_to_r_approx && r_to_q_exact && scale_round, "null conversion table");
        for (const ofhe_bconv* t : {q_to_r_exact, q_to_r_approx, r_to_q_exact, scale_round})
            OFHE_REQUIRE(t->r == ring->r, "conversion table belongs to another ring");
        OFHE_REQUIRE(q_to_r_exact->t->mode == 1 && q_to_r_approx->t->mode == 0 && r_to_q_exact->t->mode == 1 &&
                         scale_round->t->mode == 2,
                     "conversion tables have the wrong kind (exact, approximate, exact, scale-and-round expected)");
        auto* h = new ofhe_bfv();
        h->b    = new BfvMult(ring->r, sizeQ, sizeR, q_to_r_exact->t.get(), q_to_r_approx->t.get(), r_to_q_exact->t.get(),
                              scale_round->t.get());
        *out    = h;
    });
}
void ofhe_bfv_destroy(ofhe_bfv* b) {
    if (b) {
        delete b->b;
        delete b;
    }
}
int ofhe_bfv_evalmult_norelin(ofhe_bfv* b, const uint64_t* ct1_d, const uint64_t* ct2_d, uint64_t* out3_d, void* stream) {
    return guard([&] {
        use(b->b->ring);
        b->b->evalmult_norelin(ct1_d, ct2_d, out3_d, S(stream));
    });
}
int ofhe_bfv_evalmult(ofhe_bfv* b, ofhe_hks* hks, const uint64_t* ct1_d, const uint64_t* ct2_d, const ofhe_evk* evk,
                      uint64_t* out_d, void* stream) {
    return guard([&] {
        use(b->b->ring);
        b->b->evalmult(*hks->h, ct1_d, ct2_d, *evk->k, out_d, S(stream));
    });
}

// ---- the reference's wire format for DCRTPoly (cereal portable binary), see cereal_io.cpp
int ofhe_cereal_dcrtpoly_scan(const void* buf_h, size_t len, uint32_t max_towers, uint32_t* shape_out,
                              uint64_t* moduli_out, uint64_t* roots_out, size_t* offsets_out) {
    return guard([&] {
        std::vector<uint64_t> q, r;
        std::vector<size_t> off;
        const cio::DcrtInfo info = cio::scan_dcrtpoly(buf_h, len, &q, &r, &off);
        OFHE_REQUIRE(info.n_towers <= max_towers, "serialized polynomial has more towers than the caller allowed for");
        shape_out[0] = info.N, shape_out[1] = info.n_towers, shape_out[2] = info.format;
        for (uint32_t t = 0; t < info.n_towers; ++t) {
            if (moduli_out)
                moduli_out[t] = q[t];
            if (roots_out)
                roots_out[t] = r[t];
            if (offsets_out)
                offsets_out[t] = off[t];
        }
    });
}
int ofhe_cereal_dcrtpoly_write(uint32_t N, uint32_t n_towers, const uint64_t* moduli_h, const uint64_t* roots_h,
                               uint32_t format, const uint64_t* data_h, void* out_h, size_t cap, size_t* written_out) {
    return guard([&] {
        std::vector<uint8_t> bytes;
        cio::write_dcrtpoly(N, n_towers, moduli_h, roots_h, format, data_h, bytes);
        *written_out = bytes.size();  // reported even when `cap` is too small, so the caller can size the buffer
        OFHE_REQUIRE(bytes.size() <= cap, "output buffer too small for the serialized polynomial");
        std::memcpy(out_h, bytes.data(), bytes.size());
    });
}
int ofhe_dcrtpoly_upload_cereal(ofhe_ring* ring, const uint32_t* mod_idx_h, const void* buf_h, size_t len,
                                uint32_t max_towers, int check_residues, uint64_t* dst_d, uint32_t* shape_out,
                                void* stream) {
    return guard([&] {
        use(ring->r);
        std::vector<size_t> off;
        std::vector<uint64_t> moduli;
        const cio::DcrtInfo info = cio::scan_dcrtpoly(buf_h, len, &moduli, nullptr, &off);
        // the archive is untrusted input: it must describe exactly the rows the caller sized dst_d for
        OFHE_REQUIRE(info.N == ring->r->N, "serialized polynomial has a different ring dimension than the target ring");
        OFHE_REQUIRE(info.n_towers <= max_towers, "serialized polynomial has more towers than the destination slab");
        for (uint32_t t = 0; t < info.n_towers; ++t) {
            const uint32_t mi = mod_idx_h ? mod_idx_h[t] : t;
            OFHE_REQUIRE(mi < ring->r->n_moduli, "modulus index out of range");
            OFHE_REQUIRE(moduli[t] == ring->r->moduli[mi], "serialized tower modulus differs from the target ring's");
            if (check_residues) {  // the kernels assume canonical residues (Barrett / 30-bit split bounds)
                const uint64_t* w = reinterpret_cast<const uint64_t*>(static_cast<const uint8_t*>(buf_h) + off[t]);
                uint64_t bad = 0;
                for (uint32_t j = 0; j < info.N; ++j) {
                    uint64_t v;
                    std::memcpy(&v, w + j, 8);  // payload offsets are not 8-byte aligned in general
                    bad |= (uint64_t)(v >= moduli[t]);
                }
                OFHE_REQUIRE(!bad, "serialized tower holds a residue >= its modulus");
            }
        }
        if (shape_out)
            shape_out[0] = info.N, shape_out[1] = info.n_towers, shape_out[2] = info.format;
        // every tower payload is N raw little-endian words: straight from the (ideally pinned) archive to its device row
        for (uint32_t t = 0; t < info.n_towers; ++t)
            OFHE_CUDA(cudaMemcpyAsync(dst_d + (size_t)t * info.N, static_cast<const uint8_t*>(buf_h) + off[t],
                                      sizeof(uint64_t) * info.N, cudaMemcpyHostToDevice, S(stream)));
    });
}

int ofhe_fill_uniform(ofhe_ring* ring, uint64_t* dst_d, uint32_t n_limbs, const uint32_t* mod_idx_h, uint64_t seed,
                      void* stream) {
    return guard([&] {
        use(ring->r);
        LimbMap lm = checked_map(ring->r, n_limbs, mod_idx_h);
        if (n_limbs == 0)
            return;
        dim3 g(std::max(1u, std::min(64u, ring->r->N / 256)), n_limbs);
        fill_uniform_kernel<<<g, 256, 0, S(stream)>>>(ring->r->dev(), dst_d, lm, seed);
        OFHE_CUDA(cudaGetLastError());
    });
}

}  // extern "C"

namespace {
std::unique_ptr<BConv> build_bconv(const Ring& r, uint32_t n_src, const uint32_t* src, uint32_t n_dst,
                                   const uint32_t* dst) {
    OFHE_REQUIRE(n_src >= 1 && n_src <= (uint32_t)kMaxLimbs && n_dst >= 1 && n_dst <= (uint32_t)kMaxLimbs,
                 "base conversion: unsupported basis size");
    std::vector<uint64_t> q(n_src);
    for (uint32_t i = 0; i < n_src; ++i)
        q[i] = r.moduli[src[i]];
    std::vector<uint64_t> hinv(2 * (size_t)n_src);
    std::vector<uint32_t> mat(2 * (size_t)n_src * n_dst);
    for (uint32_t i = 0; i < n_src; ++i) {
        uint64_t w      = hm::inv(hm::prod_mod(q.data(), n_src, i, q[i]), q[i]);
        hinv[2 * i]     = w;
        hinv[2 * i + 1] = hm::shoup_const(w, q[i]);
        for (uint32_t j = 0; j < n_dst; ++j) {
            uint64_t m                         = hm::prod_mod(q.data(), n_src, i, r.moduli[dst[j]]);
            mat[2 * ((size_t)i * n_dst + j)]     = (uint32_t)(m & ((1u << 30) - 1));
            mat[2 * ((size_t)i * n_dst + j) + 1] = (uint32_t)(m >> 30);
        }
    }
    auto t   = std::make_unique<BConv>();
    t->sizeQ = n_src;
    t->sizeP = n_dst;
    t->qmap  = make_map(src, n_src);
    t->pmap  = make_map(dst, n_dst);
    OFHE_CUDA(cudaMalloc(&t->qhatinv_d, sizeof(uint64_t) * hinv.size()));
    OFHE_CUDA(cudaMalloc(&t->mat_d, sizeof(uint32_t) * mat.size()));
    h2d_blocking(t->qhatinv_d, hinv.data(), sizeof(uint64_t) * hinv.size());
    h2d_blocking(t->mat_d, mat.data(), sizeof(uint32_t) * mat.size());
    bconv_finalize(r, *t);
    return t;
}
}  // namespace
