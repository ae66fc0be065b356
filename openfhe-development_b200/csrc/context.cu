// context.cu -- device context and ring registration (one-time table construction + upload).
//
// A Ring is the device-resident replacement for the reference's static per-modulus twiddle caches
// (ChineseRemainderTransformFTTNat::PreCompute, math/hal/intnat/transformnat-impl.h:714-756, guarded there by
// `#pragma omp critical`): here they are built once when the ring is registered and are immutable afterwards.
#include <omp.h>

#include "hostmath.h"
#include "internal.h"

namespace ofhe {

void* Context::get_scratch(size_t bytes, cudaStream_t st) {
    std::lock_guard<std::mutex> g(mu);
    Arena& a = arenas[st];
    if (bytes > a.bytes) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (st != nullptr)
            OFHE_CUDA(cudaStreamIsCapturing(st, &cs));
        OFHE_REQUIRE(cs == cudaStreamCaptureStatusNone,
                     "workspace of this stream must grow during stream capture: call ofhe_hks_reserve (or run the "
                     "operation once eagerly) on the stream before capturing");
        if (a.p)
            retired.push_back(a.p);  // queued kernels / captured graphs may still reference it
        a.p = nullptr;
        a.bytes = 0;
        OFHE_CUDA(cudaMalloc(&a.p, bytes));
        a.bytes = bytes;
    }
    return a.p;
}
Context::~Context() {
    for (auto& kv : arenas)
        if (kv.second.p)
            cudaFree(kv.second.p);
    for (void* p : retired)
        cudaFree(p);
}

Ring::~Ring() {
    if (mc_d)
        cudaFree(mc_d);
    if (tw_fwd_d)
        cudaFree(tw_fwd_d);
    if (tw_inv_d)
        cudaFree(tw_inv_d);
    for (auto& kv : automaps)
        cudaFree(kv.second);
}

static uint32_t bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; ++i)
        r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

Ring* ring_create(Context* ctx, uint32_t N, uint32_t n_moduli, const uint64_t* moduli, const uint64_t* roots) {
    OFHE_REQUIRE(N >= 2 && (N & (N - 1)) == 0, "ring dimension must be a power of two >= 2");
    OFHE_REQUIRE(N <= (1u << 17), "ring dimension above 2^17 is not supported");
    OFHE_REQUIRE(n_moduli >= 1, "at least one modulus is required");
    uint32_t logN = 0;
    while ((1u << logN) < N)
        ++logN;
    for (uint32_t i = 0; i < n_moduli; ++i) {
        OFHE_REQUIRE(moduli[i] > 2 && moduli[i] < (uint64_t(1) << 60), "moduli must be odd primes below 2^60");
        OFHE_REQUIRE((moduli[i] - 1) % (2 * (uint64_t)N) == 0, "modulus is not 1 mod 2N");
        OFHE_REQUIRE(hm::pow(roots[i], N, moduli[i]) == moduli[i] - 1, "root is not a primitive 2N-th root of unity");
    }
    OFHE_CUDA(cudaSetDevice(ctx->device));
    auto* r     = new Ring();
    r->ctx      = ctx;
    r->N        = N;
    r->logN     = logN;
    r->n_moduli = n_moduli;
    r->moduli.assign(moduli, moduli + n_moduli);
    r->roots.assign(roots, roots + n_moduli);
    r->mc_h.resize(n_moduli);

    std::vector<Twiddle> fwd((size_t)n_moduli * N), inv((size_t)n_moduli * N);
#pragma omp parallel for schedule(dynamic)
    for (long mi = 0; mi < (long)n_moduli; ++mi) {
        const uint64_t q = moduli[mi], psi = roots[mi], psi_inv = hm::inv(psi, q);
        Twiddle* f = fwd.data() + (size_t)mi * N;
        Twiddle* v = inv.data() + (size_t)mi * N;
        uint64_t x = 1, xi = 1;
        for (uint32_t i = 0; i < N; ++i) {
            uint32_t j = bitrev(i, logN);
            f[j]       = Twiddle{x, hm::shoup_const(x, q)};
            v[j]       = Twiddle{xi, hm::shoup_const(xi, q)};
            x          = hm::mul(x, psi, q);
            xi         = hm::mul(xi, psi_inv, q);
        }
        ModConst& mc = r->mc_h[mi];
        mc.q         = q;
        mc.k         = hm::bit_length(q);
        mc.pad       = 0;
        mc.mu        = (uint64_t)(((hm::u128)1 << (63 + mc.k)) / q);
        hm::u128 mu128 = ~(hm::u128)0 / q;  // == floor(2^128 / q) for q not a power of two
        mc.mu128_hi  = (uint64_t)(mu128 >> 64);
        mc.mu128_lo  = (uint64_t)mu128;
        mc.r30       = (uint64_t(1) << 30) % q;
        mc.r30s      = hm::shoup_const(mc.r30, q);
        mc.r60       = (uint64_t(1) << 60) % q;
        mc.r60s      = hm::shoup_const(mc.r60, q);
        mc.one_s     = hm::shoup_const(1 % q, q);
        mc.nInv      = hm::inv(N % q, q);
        mc.nInvS     = hm::shoup_const(mc.nInv, q);
        mc.w1nInv    = hm::mul(v[1].w, mc.nInv, q);
        mc.w1nInvS   = hm::shoup_const(mc.w1nInv, q);
    }
    try {
        OFHE_CUDA(cudaMalloc(&r->mc_d, sizeof(ModConst) * n_moduli));
        OFHE_CUDA(cudaMalloc(&r->tw_fwd_d, sizeof(Twiddle) * fwd.size()));
        OFHE_CUDA(cudaMalloc(&r->tw_inv_d, sizeof(Twiddle) * inv.size()));
        h2d_blocking(r->mc_d, r->mc_h.data(), sizeof(ModConst) * n_moduli);
        h2d_blocking(r->tw_fwd_d, fwd.data(), sizeof(Twiddle) * fwd.size());
        h2d_blocking(r->tw_inv_d, inv.data(), sizeof(Twiddle) * inv.size());
    }
    catch (...) {
        delete r;
        throw;
    }
    return r;
}

void make_inv_fold(const Ring& r, uint32_t mi, uint64_t s, uint64_t out4[4]) {
    const ModConst& mc = r.mc_h[mi];
    out4[0] = hm::mul(mc.nInv, s % mc.q, mc.q);
    out4[1] = hm::shoup_const(out4[0], mc.q);
    out4[2] = hm::mul(mc.w1nInv, s % mc.q, mc.q);
    out4[3] = hm::shoup_const(out4[2], mc.q);
}

Context* context_create(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                    "): this backend has no CPU fallback");
    OFHE_REQUIRE(device >= 0 && device < count, "invalid device ordinal");
    OFHE_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop{};
    OFHE_CUDA(cudaGetDeviceProperties(&prop, device));
    auto* c     = new Context();
    c->device   = device;
    c->sm_count = prop.multiProcessorCount;
    return c;
}

const uint32_t* ring_automap(Ring* r, uint32_t k, cudaStream_t st) {
    std::lock_guard<std::mutex> g(r->ctx->mu);
    auto it = r->automaps.find(k);
    if (it != r->automaps.end())
        return it->second;
    uint32_t* d = nullptr;
    OFHE_CUDA(cudaMalloc(&d, sizeof(uint32_t) * r->N));
    // one-time build on a PRIVATE stream, finished before the pointer is handed out: later calls may come from any
    // stream, and the caller's stream must not be waited for here -- it may hold a device barrier whose partner
    // ranks this very host thread has yet to launch (virtual ranks of one process, tests/test_gpu_shard.py)
    (void)st;
    cudaStream_t tmp;
    OFHE_CUDA(cudaStreamCreateWithFlags(&tmp, cudaStreamNonBlocking));
    launch_build_automap(r->N, r->logN, k, d, tmp);
    const cudaError_t e = cudaStreamSynchronize(tmp);
    cudaStreamDestroy(tmp);
    OFHE_CUDA(e);
    r->automaps[k] = d;
    return d;
}

}  // namespace ofhe
