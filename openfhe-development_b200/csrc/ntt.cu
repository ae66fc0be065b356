// ntt.cu -- negacyclic NTT / inverse NTT over [limbs x N] slabs of 64-bit residues, sm_100a.
//
// Replaces ChineseRemainderTransformFTTNat::{ForwardTransformToBitReverseInPlace,
// InverseTransformFromBitReverseInPlace} (src/core/include/math/hal/intnat/transformnat-impl.h:303-374,
// 512-625, 648-691) and the limb loop of DCRTPolyImpl::SwitchFormat (lattice/hal/default/dcrtpoly-impl.h:1932).
// Same contract: forward = natural order in, bit-reversed order out; inverse = the opposite; canonical
// residues in [0,q) on exit; twiddle tables T[brev(i)] = psi^i with Shoup constants.
//
// B200 decomposition (not the reference's stage-by-stage loop), N = 2^(12+A), A = 0..4:
//   ONE kernel, ONE pass over HBM.  A limb is transformed by a thread-block CLUSTER of 2^A CTAs; CTA c owns
//   the contiguous 4096-point tile c (32 KiB of shared memory), 256 threads x 16 points.
//   forward : phase A -- the A stages whose butterfly span is >= 4096: every thread loads 16 points with stride
//                        4096 straight from HBM (a warp reads 256 contiguous bytes per point), runs a radix-2^A
//                        register round and scatters result k into tile k of CTA k through DISTRIBUTED SHARED
//                        MEMORY (st.shared::cluster, 256 contiguous bytes per warp and destination);
//             cluster barrier;
//             phase B -- the last 12 stages on the CTA's own tile: three radix-16 register rounds with two
//                        shared-memory exchanges (the second XOR-swizzled), 128-bit stores of canonical residues.
//   inverse : the mirror image (phase B first, DSMEM scatter, cluster barrier, strided radix-2^A round with
//             N^-1 -- and optionally one more per-limb scalar -- folded into the last stage, optionally storing the
//             finished rows into several destination buffers at once: NVLink peer pushes of shard.cu).
//   N = 2^17 (a limb does not fit 16 x 32 KiB) keeps the two-kernel split [strided | tile] through L2;
//   N < 4096 (the reference's unit-test sizes) runs one CTA per limb in shared memory.
// Butterflies: see ntt_core.cuh (truncated Shoup quotient, compile-time lazy range schedule).
#include <algorithm>
#include <cstdlib>

#include "ntt_core.cuh"

namespace ofhe {

namespace {
using namespace nttc;
// NOTE: `in` and `out` may be the same slab (in-place transforms), so data pointers carry no __restrict__.

constexpr int kTile = 4096;
__device__ __forceinline__ uint32_t swz(uint32_t c) {
    return c ^ ((c >> 4) & 15u);
}

// finished rows -> one or several destination slabs (PeerOut); constant indices keep the pointer array in the
// kernel parameter bank (a run-time index would force a local-memory copy)
template <typename F>
__device__ __forceinline__ void for_each_dest(uint64_t* out, const PeerOut& po, F&& f) {
    if (po.n == 0) {
        f(out);
        return;
    }
#pragma unroll
    for (uint32_t d = 0; d < kMaxPeers; ++d)
        if (d < po.n)
            f(po.base[d]);
}

// ---- phase B of the forward transform: stages S0 .. S0+11 on the tile held in sm (linear layout on entry) ----
// FROM_GLOBAL: S0 == 0 (N = 4096), round A reads the canonical input straight from `in`.
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
};
// `after_first_barrier` runs right behind the first block barrier of the tile (the streaming kernels refill the
// OTHER staging buffer there: every thread has then finished reading it)
#if defined(OFHE_NTT_EXP) && OFHE_NTT_EXP == 3  // TIMING EXPERIMENT ONLY (wrong results): no block barriers in the forward rounds
#define TILE_SYNC() __syncwarp()
#else
#define TILE_SYNC() __syncthreads()
#endif
template <int LOGN, typename Hook = NoHook>
__device__ __forceinline__ void fwd_tile_rounds(uint64_t* sm, uint64_t (&x)[16], TwWin& w, const Twiddle* tw,
                                                uint32_t tile, const Mod& m, uint32_t t,
                                                Hook after_first_barrier = Hook{}) {
    constexpr int S0 = LOGN - 12;
    const uint32_t bh = t >> 4, bl = t & 15u, tb = (1u << S0) + tile;
    // ---- round A: tile bits 11..8, thread t holds c = t + 256k (already in x); w = tw_preload<false>(tw, tb)
    ct_round16_df<LOGN, S0>(x, w, tw, tb, m);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sm[t + 256 * k] = x[k];
    tw_preload<false>(w, tw, tb * 16u + bh);
    TILE_SYNC();
    after_first_barrier();
    // ---- round B: tile bits 7..4, c = (t>>4)*256 + (t&15) + 16k
    prefetch_round(tw, tb * 256u + t);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = sm[bh * 256 + bl + 16 * k];
    ct_round16_df<LOGN, S0 + 4>(x, w, tw, tb * 16u + bh, m);
    TILE_SYNC();
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sm[swz(bh * 256 + bl + 16 * k)] = x[k];
    tw_preload<false>(w, tw, tb * 256u + t);
    TILE_SYNC();
    // ---- round C: tile bits 3..0, c = 16t + k (128 contiguous bytes per thread)
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = sm[swz(16 * t + k)];
    ct_round16_df<LOGN, S0 + 8>(x, w, tw, tb * 256u + t, m);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = canon8(x[k], m);
}
// ---- phase B of the inverse transform: stages LOGN-1 .. LOGN-12 (mirror order); on exit thread t holds tile
// offsets t + 256k.  LAST: N = 4096, round A contains the folded final stage.
template <bool LAST, typename Hook = NoHook>
__device__ __forceinline__ void inv_tile_rounds(uint64_t* sm, uint64_t (&x)[16], TwWin& w, const Twiddle* tw,
                                                uint32_t tbase, const LastConst& lc, const Mod& m, uint32_t t,
                                                Hook after_first_barrier = Hook{}) {
    // tbase = 2^S0 + tile; w = tw_preload<true>(tw, tbase * 256 + t)
    const uint32_t bh = t >> 4, bl = t & 15u;
    prefetch_round(tw, tbase * 16u + bh);
    gs_round16_df<false>(x, w, tw, tbase * 256u + t, lc, m);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sm[swz(16 * t + k)] = x[k];
    tw_preload<true>(w, tw, tbase * 16u + bh);
    __syncthreads();
    after_first_barrier();
    prefetch_round(tw, tbase);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = sm[swz(bh * 256 + bl + 16 * k)];
    gs_round16_df<false>(x, w, tw, tbase * 16u + bh, lc, m);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sm[bh * 256 + bl + 16 * k] = x[k];
    tw_preload<true>(w, tw, tbase);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = sm[t + 256 * k];
    gs_round16_df<LAST>(x, w, tw, tbase, lc, m);
}

// phase A register rounds: G independent columns of 2^A strided points each, x[g * 2^A + k]
template <int LOGN, int A, int G, int g = 0>
__device__ __forceinline__ void fwd_columns(uint64_t* x, const Twiddle* tw, const Mod& m) {
    if constexpr (g < G) {
        ct_round<LOGN, 0, A, g * (1 << A)>(x, tw, 1u, m);
        fwd_columns<LOGN, A, G, g + 1>(x, tw, m);
    }
}
template <int A, int G, int g = 0>
__device__ __forceinline__ void inv_columns(uint64_t* x, const Twiddle* tw, const LastConst& lc, const Mod& m) {
    if constexpr (g < G) {
        gs_round<A, true, g * (1 << A)>(x, tw, 1u, lc, m);
        inv_columns<A, G, g + 1>(x, tw, lc, m);
    }
}

// ------------------------------------------------------------------------------------------------
// Cluster kernels.  grid = (2^A, limbs), cluster = (2^A, 1, 1), 256 threads.  CL = 2^A tiles, G = 16 / CL
// columns of CL strided points per thread in phase A.
#ifndef OFHE_NTT_MINB
#define OFHE_NTT_MINB 3  // resident CTAs per SM the kernels are compiled for (register cap 65536 / (256 * MINB))
#endif
template <int A>
__global__ void __launch_bounds__(256, OFHE_NTT_MINB) ntt_fwd_cluster(RingDev r, const uint64_t* in, uint64_t* out,
                                                                        LimbMap lm, PosMap op) {
    constexpr int LOGN = 12 + A, CL = 1 << A, G = 16 / CL;
    __shared__ __align__(16) uint64_t sm[kTile];
    const uint32_t limb = blockIdx.y, t = threadIdx.x;
    const uint32_t c    = A ? cluster_ctarank() : 0u;
    const uint32_t mi   = lm.idx[limb];
    const Mod m         = make_mod(r.mc[mi]);
    const Twiddle* tw   = r.tw_fwd + (size_t)mi * r.N;
    const uint64_t* src = in + (size_t)lm.pos[limb] * r.N;
    uint64_t* dst       = out + (size_t)op.pos[limb] * r.N + (size_t)c * kTile;
    uint64_t x[16];
    prefetch_round(tw, (1u << A) + c);
    if constexpr (A == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            x[k] = __ldcs(src + t + 256 * k);
    }
    else {
        cluster_arrive_relaxed();  // every CTA of the cluster is running before anyone touches its shared memory
        // ---- phase A: columns j = (c*G + g)*256 + t, points j + 4096k
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int k = 0; k < CL; ++k)
                x[g * CL + k] = __ldcs(src + (c * G + g) * 256 + t + (size_t)kTile * k);
        fwd_columns<LOGN, A, G>(x, tw, m);
        // ---- scatter: point k of column j belongs to tile k, offset j
        const uint32_t sm0 = smem_u32(sm);
        cluster_wait();
#pragma unroll
        for (int k = 0; k < CL; ++k) {
            const uint32_t remote = mapa(sm0, (uint32_t)k);
#pragma unroll
            for (int g = 0; g < G; ++g)
                st_cluster(remote + 8u * ((c * G + g) * 256 + t), x[g * CL + k]);
        }
        cluster_arrive();
        cluster_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k)
            x[k] = sm[t + 256 * k];
    }
    TwWin ha;
    tw_preload<false>(ha, tw, (1u << A) + c);
    fwd_tile_rounds<LOGN>(sm, x, ha, tw, c, m, t);
    ulonglong2* d2 = reinterpret_cast<ulonglong2*>(dst + 16 * t);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        __stcs(d2 + k, make_ulonglong2(x[2 * k], x[2 * k + 1]));
}

template <int A>
__global__ void __launch_bounds__(256, OFHE_NTT_MINB) ntt_inv_cluster(RingDev r, const uint64_t* in, uint64_t* out,
                                                                        LimbMap lm, PosMap op,
                                                                        const uint64_t* __restrict__ fold, PeerOut po) {
    constexpr int CL = 1 << A, G = 16 / CL;
    extern __shared__ __align__(16) uint64_t dsm[];  // [tile 4096][receive 4096 if A > 0]
    uint64_t* sm   = dsm;
    uint64_t* rbuf = dsm + kTile;
    const uint32_t limb = blockIdx.y, t = threadIdx.x;
    const uint32_t c    = A ? cluster_ctarank() : 0u;
    const uint32_t mi   = lm.idx[limb];
    const ModConst mc   = r.mc[mi];
    const Mod m         = make_mod(mc);
    const Twiddle* tw   = r.tw_inv + (size_t)mi * r.N;
    const LastConst lc  = last_const(mc, fold, limb);
    const size_t obase  = (size_t)op.pos[limb] * r.N;
    if constexpr (A > 0)
        cluster_arrive_relaxed();
    prefetch_round(tw, ((1u << A) + c) * 256u + t);
    uint64_t x[16];
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(in + (size_t)lm.pos[limb] * r.N + (size_t)c * kTile + 16 * t);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const ulonglong2 v = __ldcs(src + k);
        x[2 * k]           = v.x;
        x[2 * k + 1]       = v.y;
    }
    TwWin ha;
    tw_preload<true>(ha, tw, ((1u << A) + c) * 256u + t);
    inv_tile_rounds<A == 0>(sm, x, ha, tw, (1u << A) + c, lc, m, t);
    if constexpr (A == 0) {
        for_each_dest(out, po, [&](uint64_t* o) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                o[obase + t + 256 * k] = x[k];
        });
    }
    else {
        // thread t holds tile offsets t + 256k of tile c: offset 256k + t is column (k % G)*256 + t of CTA k / G
        const uint32_t r0 = smem_u32(rbuf);
        cluster_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k)
            st_cluster(mapa(r0, (uint32_t)(k / G)) + 8u * (((k % G) * CL + c) * 256 + t), x[k]);
        cluster_arrive();
        cluster_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
            x[i] = rbuf[i * 256 + t];
        inv_columns<A, G>(x, tw, lc, m);
        for_each_dest(out, po, [&](uint64_t* o) {
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int k = 0; k < CL; ++k)
                    __stcs(o + obase + (c * G + g) * 256 + t + (size_t)kTile * k, x[g * CL + k]);
        });
    }
}

// ------------------------------------------------------------------------------------------------
// Split kernels (N = 2^17 only): the A = 5 strided stages as a register round straight from/to HBM, then the
// tile kernel; the lazy intermediate ([0,16q) forward, [0,4q) inverse) stays in L2 for slabs that fit it.
template <int A>
__global__ void __launch_bounds__(256) ntt_fwd_strided(RingDev r, const uint64_t* in, uint64_t* out, LimbMap lm, PosMap op) {
    const uint32_t limb   = blockIdx.y;
    const uint32_t mi     = lm.idx[limb];
    const uint32_t stride = r.N >> A;
    const uint32_t c      = blockIdx.x * blockDim.x + threadIdx.x;
    const Mod m           = make_mod(r.mc[mi]);
    const uint64_t* src   = in + (size_t)lm.pos[limb] * r.N + c;
    uint64_t* dst         = out + (size_t)op.pos[limb] * r.N + c;
    uint64_t x[1 << A];
#pragma unroll
    for (int k = 0; k < (1 << A); ++k)
        x[k] = src[(size_t)k * stride];
    ct_round<12 + A, 0, A>(x, r.tw_fwd + (size_t)mi * r.N, 1u, m);
#pragma unroll
    for (int k = 0; k < (1 << A); ++k)
        dst[(size_t)k * stride] = x[k];
}
template <int A>
__global__ void __launch_bounds__(256, OFHE_NTT_MINB) ntt_fwd_tile(RingDev r, uint64_t* data, LimbMap lm) {
    __shared__ __align__(16) uint64_t sm[kTile];
    const uint32_t limb = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
    const uint32_t mi   = lm.idx[limb];
    const Mod m         = make_mod(r.mc[mi]);
    uint64_t* p         = data + (size_t)lm.pos[limb] * r.N + (size_t)tile * kTile;
    const Twiddle* tw   = r.tw_fwd + (size_t)mi * r.N;
    uint64_t x[16];
    prefetch_round(tw, (1u << A) + tile);
#pragma unroll
    for (int k = 0; k < 16; ++k)
#if defined(OFHE_NTT_EXP) && OFHE_NTT_EXP == 5  // TIMING EXPERIMENT ONLY (wrong results): no global data load
        x[k] = (uint64_t)(t + 256 * k) * 0x9E3779B97F4A7C15ull >> 5;
#else
        x[k] = p[t + 256 * k];
#endif
    TwWin ha;
    tw_preload<false>(ha, tw, (1u << A) + tile);
    fwd_tile_rounds<12 + A>(sm, x, ha, tw, tile, m, t);
    ulonglong2* d2 = reinterpret_cast<ulonglong2*>(p + 16 * t);
#if defined(OFHE_NTT_EXP) && OFHE_NTT_EXP == 6  // TIMING EXPERIMENT ONLY (wrong results): one store instead of eight
    uint64_t acc = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        acc ^= x[k];
    if (acc == 0x1234567)
        d2[0] = make_ulonglong2(acc, acc);
#else
#pragma unroll
    for (int k = 0; k < 8; ++k)
        d2[k] = make_ulonglong2(x[2 * k], x[2 * k + 1]);
#endif
}
// the same tile pass with the ApproxModDown tail as its epilogue (NttTail, internal.h): thread t ends with the 16
// canonical values of tile offsets 16t .. 16t+15 in registers and combines them with the minuend / addend rows on
// their way out -- the transform's result itself is never stored
template <int A>
__global__ void __launch_bounds__(256, OFHE_NTT_MINB) ntt_fwd_tile_tail(RingDev r, const uint64_t* data, LimbMap lm, NttTail tl) {
    __shared__ __align__(16) uint64_t sm[kTile];
    const uint32_t limb = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
    const uint32_t mi   = lm.idx[limb];
    const Mod m         = make_mod(r.mc[mi]);
    const uint64_t* p   = data + (size_t)lm.pos[limb] * r.N + (size_t)tile * kTile;
    const Twiddle* tw   = r.tw_fwd + (size_t)mi * r.N;
    uint64_t x[16];
    prefetch_round(tw, (1u << A) + tile);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = p[t + 256 * k];
    TwWin ha;
    tw_preload<false>(ha, tw, (1u << A) + tile);
    fwd_tile_rounds<12 + A>(sm, x, ha, tw, tile, m, t);
    const uint64_t q = r.mc[mi].q, s = tl.scal[2 * limb], ss = tl.scal[2 * limb + 1];
    const size_t off = (size_t)tile * kTile + 16 * t;
    const ulonglong2* e2 = reinterpret_cast<const ulonglong2*>(tl.e + (size_t)tl.erow.pos[limb] * r.N + off);
    const ulonglong2* a2 =
        tl.addend ? reinterpret_cast<const ulonglong2*>(tl.addend + (size_t)tl.drow.pos[limb] * r.N + off) : nullptr;
    ulonglong2* d2 = reinterpret_cast<ulonglong2*>(tl.dst + (size_t)tl.drow.pos[limb] * r.N + off);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const ulonglong2 ev = e2[k];
        ulonglong2 o;
        o.x = shoup(submod(ev.x, x[2 * k], q), s, ss, q);
        o.y = shoup(submod(ev.y, x[2 * k + 1], q), s, ss, q);
        if (a2) {
            const ulonglong2 z = a2[k];
            o.x = addmod(o.x, z.x, q);
            o.y = addmod(o.y, z.y, q);
        }
        d2[k] = o;
    }
}
template <int A>
__global__ void __launch_bounds__(256, OFHE_NTT_MINB) ntt_inv_tile(RingDev r, const uint64_t* in, uint64_t* out,
                                                                     LimbMap lm, PosMap op) {
    __shared__ __align__(16) uint64_t sm[kTile];
    const uint32_t limb = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
    const uint32_t mi   = lm.idx[limb];
    const ModConst mc   = r.mc[mi];
    const Mod m         = make_mod(mc);
    const Twiddle* tw   = r.tw_inv + (size_t)mi * r.N;
    uint64_t x[16];
    prefetch_round(tw, ((1u << A) + tile) * 256u + t);
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(in + (size_t)lm.pos[limb] * r.N + (size_t)tile * kTile + 16 * t);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const ulonglong2 v = src[k];
        x[2 * k]           = v.x;
        x[2 * k + 1]       = v.y;
    }
    TwWin ha;
    tw_preload<true>(ha, tw, ((1u << A) + tile) * 256u + t);
    inv_tile_rounds<false>(sm, x, ha, tw, (1u << A) + tile, LastConst{}, m, t);
    uint64_t* o = out + (size_t)op.pos[limb] * r.N + (size_t)tile * kTile;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        o[t + 256 * k] = x[k];
}
template <int A>
__global__ void __launch_bounds__(256) ntt_inv_strided(RingDev r, uint64_t* data, LimbMap lm,
                                                        const uint64_t* __restrict__ fold, PeerOut po) {
    const uint32_t limb   = blockIdx.y;
    const uint32_t mi     = lm.idx[limb];
    const uint32_t stride = r.N >> A;
    const uint32_t c      = blockIdx.x * blockDim.x + threadIdx.x;
    const ModConst mc     = r.mc[mi];
    const Mod m           = make_mod(mc);
    const size_t off      = (size_t)lm.pos[limb] * r.N + c;
    uint64_t x[1 << A];
#pragma unroll
    for (int k = 0; k < (1 << A); ++k)
        x[k] = data[off + (size_t)k * stride];
    gs_round<A, true>(x, r.tw_inv + (size_t)mi * r.N, 1u, last_const(mc, fold, limb), m);
    for_each_dest(data, po, [&](uint64_t* o) {
#pragma unroll
        for (int k = 0; k < (1 << A); ++k)
            o[off + (size_t)k * stride] = x[k];
    });
}

// ------------------------------------------------------------------------------------------------
// Streaming tile kernels: the tile half of the split pair as PERSISTENT CTAs.  CTA b transforms tiles b, b + grid,
// b + 2 grid, ... of the launch's (limb, tile) list and keeps two 32 KiB staging buffers: while the warps run the
// twelve stages of one tile, the copy engine fills the other buffer with the next one, so no warp ever sits on a
// global load with its 80 registers allocated and there is no CTA launch/drain between tiles.
//   forward: one 32 KiB 1-D TMA bulk copy per tile (cp.async.bulk, completion by mbarrier transaction count); the
//            radix-16 column read t + 256k of the linear tile is bank-conflict free.
//   inverse: the first round needs 16 CONTIGUOUS points per thread, which conflicts 8-way in a linear tile, so the
//            tile is staged by 16-byte cp.async copies (coalesced on the global side) into 128-byte rows whose
//            eight chunks are XOR-swizzled by the row index; each thread then reads its row with 128-bit loads.
// The refill of a buffer is issued behind the first block barrier of the FOLLOWING tile (every thread is then done
// reading it) and has two thirds of a tile's compute time to land.
constexpr uint32_t kTileBytes = kTile * sizeof(uint64_t);

template <int A>
__global__ void __launch_bounds__(256, OFHE_NTT_MINB) ntt_fwd_tile_stream(RingDev r, uint64_t* data, LimbMap lm,
                                                                            uint32_t n_items) {
    extern __shared__ __align__(1024) uint64_t stg[];  // [2][kTile]
    __shared__ __align__(8) uint64_t full[2];
    const uint32_t t = threadIdx.x;
    auto tile_ptr = [&](uint32_t item) {
        return data + (size_t)lm.pos[item >> A] * r.N + (size_t)(item & ((1u << A) - 1u)) * kTile;
    };
    auto refill = [&](uint32_t item, uint32_t b) {  // one thread
        mbar_expect_tx(&full[b], kTileBytes);
        tma_load_1d(stg + b * kTile, tile_ptr(item), kTileBytes, &full[b]);
    };
    if (t == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
        if (blockIdx.x < n_items)
            refill(blockIdx.x, 0);
    }
    __syncthreads();
    uint32_t it = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t b = it & 1u, limb = item >> A, tile = item & ((1u << A) - 1u);
        uint64_t* sm      = stg + b * kTile;
        const uint32_t mi = lm.idx[limb];
        const Mod m       = make_mod(r.mc[mi]);
        const Twiddle* tw = r.tw_fwd + (size_t)mi * r.N;
        TwWin ha;
        tw_preload<false>(ha, tw, (1u << A) + tile);
        mbar_wait(&full[b], (it >> 1) & 1u);
        uint64_t x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            x[k] = sm[t + 256 * k];
        fwd_tile_rounds<12 + A>(sm, x, ha, tw, tile, m, t, [&] {
            if (t == 0 && item + gridDim.x < n_items)
                refill(item + gridDim.x, b ^ 1u);
        });
        // this thread's generic-proxy accesses to the buffer are ordered before the copy engine's next refill of it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        ulonglong2* d2 = reinterpret_cast<ulonglong2*>(tile_ptr(item) + 16 * t);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            d2[k] = make_ulonglong2(x[2 * k], x[2 * k + 1]);
    }
}

template <int A>
__global__ void __launch_bounds__(256, OFHE_NTT_MINB) ntt_inv_tile_stream(RingDev r, const uint64_t* in, uint64_t* out,
                                                                            LimbMap lm, PosMap op, uint32_t n_items) {
    extern __shared__ __align__(1024) uint64_t stg[];  // [2][kTile]
    __shared__ __align__(8) uint64_t full[2];
    const uint32_t t = threadIdx.x;
    auto refill = [&](uint32_t item, uint32_t b) {  // all threads: chunk j = t + 256c of the tile's 2048 chunks
        const uint64_t* src = in + (size_t)lm.pos[item >> A] * r.N + (size_t)(item & ((1u << A) - 1u)) * kTile;
        const uint32_t base = smem_u32(stg + b * kTile);
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) {
            const uint32_t j = t + 256u * c, row = j >> 3, ch = j & 7u;
            cp_async16(base + row * 128u + ((ch ^ (row & 7u)) << 4), src + 2 * j);
        }
        cp_async_arrive(&full[b]);
    };
    if (t == 0) {
        mbar_init(&full[0], 256);
        mbar_init(&full[1], 256);
        fence_mbar_init();
    }
    __syncthreads();
    if (blockIdx.x < n_items)
        refill(blockIdx.x, 0);
    uint32_t it = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t b = it & 1u, limb = item >> A, tile = item & ((1u << A) - 1u);
        uint64_t* sm      = stg + b * kTile;
        const uint32_t mi = lm.idx[limb];
        const Mod m       = make_mod(r.mc[mi]);
        const Twiddle* tw = r.tw_inv + (size_t)mi * r.N;
        TwWin ha;
        tw_preload<true>(ha, tw, ((1u << A) + tile) * 256u + t);
        mbar_wait(&full[b], (it >> 1) & 1u);
        uint64_t x[16];
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sm + 16 * t + 2 * (c ^ (t & 7u)));
            x[2 * c]           = v.x;
            x[2 * c + 1]       = v.y;
        }
        inv_tile_rounds<false>(sm, x, ha, tw, (1u << A) + tile, LastConst{}, m, t, [&] {
            if (item + gridDim.x < n_items)
                refill(item + gridDim.x, b ^ 1u);
        });
        uint64_t* o = out + (size_t)op.pos[limb] * r.N + (size_t)tile * kTile;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            o[t + 256 * k] = x[k];
    }
}

// ------------------------------------------------------------------------------------------------
// Small rings (N < 4096: the reference's unit-test sizes): one CTA per limb, whole limb in shared memory,
// one stage per barrier.  Not a performance path.  Forward points stay below 8q (X reduced below 4q every stage).
template <bool INV>
__global__ void ntt_small(RingDev r, const uint64_t* in, uint64_t* out, LimbMap lm, PosMap op,
                          const uint64_t* __restrict__ fold, PeerOut po) {
    extern __shared__ __align__(16) uint64_t s[];
    const uint32_t limb = lm.pos[blockIdx.x], olimb = op.pos[blockIdx.x], mi = lm.idx[blockIdx.x], N = r.N;
    const ModConst mc = r.mc[mi];
    const Mod m       = make_mod(mc);
    const Twiddle* tw = (INV ? r.tw_inv : r.tw_fwd) + (size_t)mi * N;
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
        s[i] = in[(size_t)limb * N + i];
    __syncthreads();
    if (!INV) {
        for (uint32_t mm = 1, t = N >> 1; mm < N; mm <<= 1, t >>= 1) {
            for (uint32_t b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
                const uint32_t i = b / t, j = 2 * i * t + (b % t);
                const Twiddle w  = ld_tw(tw + mm + i);
                const uint64_t x = csub4i(s[j], m.c4);
                const uint64_t v = shoup_t(s[j + t], w.w, w.ws, m.nq);
                s[j]             = x + v;
                s[j + t]         = x - v + m.c4;
            }
            __syncthreads();
        }
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
            out[(size_t)olimb * N + i] = canon8(s[i], m);
    }
    else {
        for (uint32_t mm = N >> 1, t = 1; mm > 1; mm >>= 1, t <<= 1) {
            for (uint32_t b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
                const uint32_t i = b / t, j = 2 * i * t + (b % t);
                gs_bfly(s[j], s[j + t], ld_tw(tw + mm + i), m);
            }
            __syncthreads();
        }
        const uint32_t h   = N >> 1;
        const LastConst lc = last_const(mc, fold, blockIdx.x);
        for (uint32_t j = threadIdx.x; j < h; j += blockDim.x) {
            const uint64_t a = csub4i(shoup2(s[j] + s[j + h], lc.a, lc.as, m.nq), m.q);
            const uint64_t e = csub4i(shoup2(s[j] - s[j + h] + m.c4, lc.b, lc.bs, m.nq), m.q);
            for_each_dest(out, po, [&](uint64_t* o) {
                o[(size_t)olimb * N + j]     = a;
                o[(size_t)olimb * N + j + h] = e;
            });
        }
    }
}

// ---- launchers ------------------------------------------------------------------------------------------
LimbMap on_output(const LimbMap& lm, const PosMap& op) {
    LimbMap o = lm;
    for (uint32_t i = 0; i < lm.n; ++i)
        o.pos[i] = op.pos[i];
    return o;
}
PosMap default_pos(const LimbMap& lm, const PosMap* out_pos) {
    if (out_pos)
        return *out_pos;
    PosMap p;
    for (uint32_t i = 0; i < lm.n; ++i)
        p.pos[i] = lm.pos[i];
    return p;
}

template <typename K, typename... Args>
void launch_cluster(K kernel, int A, uint32_t limbs, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3(1u << A, limbs, 1);
    cfg.blockDim           = dim3(256, 1, 1);
    cfg.dynamicSmemBytes   = smem;
    cfg.stream             = st;
    cudaLaunchAttribute at[1];
    at[0].id               = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1u << A;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs              = at;
    cfg.numAttrs           = A ? 1 : 0;
    OFHE_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
}
// one-time per-kernel attributes (16-CTA clusters are a non-portable size; the inverse needs 64 KiB of dynamic smem)
template <int A>
void fwd_cluster(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, const PosMap& op, cudaStream_t st) {
    static const bool once = [] {
        if (A > 3)
            OFHE_CUDA(cudaFuncSetAttribute(ntt_fwd_cluster<A>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        return true;
    }();
    (void)once;
    launch_cluster(ntt_fwd_cluster<A>, A, lm.n, 0, st, r.dev(), in, out, lm, op);
}
template <int A>
void inv_cluster(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, const PosMap& op,
                 const uint64_t* fold, const PeerOut& po, cudaStream_t st) {
    constexpr size_t smem = sizeof(uint64_t) * kTile * (A ? 2 : 1);
    static const bool once = [] {
        if (A > 3)
            OFHE_CUDA(cudaFuncSetAttribute(ntt_inv_cluster<A>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        OFHE_CUDA(cudaFuncSetAttribute(ntt_inv_cluster<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return true;
    }();
    (void)once;
    launch_cluster(ntt_inv_cluster<A>, A, lm.n, smem, st, r.dev(), in, out, lm, op, fold, po);
}

// ---- split path: [strided | tile] kernel pairs over chunks of limbs small enough for the lazy intermediate to
// stay in L2 between the two kernels (the pair then costs one HBM read and one HBM write per limb)
// Which decomposition runs for a ring of 2^(12+A) points: 0 = one cluster kernel (one HBM pass, DSMEM exchange),
// 1 = the [strided | tile] kernel pair through L2.  Chosen from measurements on B200 (profiles/r02c_variants.txt):
// the cluster-wide barrier costs the 16-CTA cluster more than the second pass through L2 costs the pair, the smaller
// clusters win or tie in the forward direction.  OFHE_NTT_MODE=cluster|split overrides (measurement switch).
int ntt_mode(bool fwd, int A) {
    static const int forced = [] {
        const char* e = getenv("OFHE_NTT_MODE");
        return !e ? -1 : (e[0] == 's' ? 1 : 0);
    }();
    if (forced >= 0)
        return forced;
    return (fwd && A <= 3) ? 0 : 1;
}
uint32_t ntt_chunk_limbs(const Ring& r) {
    static const size_t bytes = [] {
        const char* e = getenv("OFHE_NTT_CHUNK_MIB");
        return (size_t)(e ? atoi(e) : 1 << 20) << 20;  // default: no chunking (tails cost more than the L2 misses)
    }();
    const uint32_t n = (uint32_t)(bytes / (sizeof(uint64_t) * r.N));
    return n < 1 ? 1 : n;
}
void sub_maps(const LimbMap& lm, const PosMap& op, uint32_t s, uint32_t n, LimbMap& l2, PosMap& o2) {
    l2.n = n;
    for (uint32_t i = 0; i < n; ++i) {
        l2.idx[i] = lm.idx[s + i];
        l2.pos[i] = lm.pos[s + i];
        o2.pos[i] = op.pos[s + i];
    }
}
// grid of a streaming launch: every CTA gets the same number of tiles (+-1) and at most OFHE_NTT_MINB CTAs share an SM
uint32_t stream_grid(const Ring& r, uint32_t n_items) {
    const uint32_t cap    = (uint32_t)r.ctx->sm_count * OFHE_NTT_MINB;
    const uint32_t rounds = (n_items + cap - 1) / cap;
    return (n_items + rounds - 1) / rounds;
}
// OFHE_NTT_STREAM=1 selects the persistent TMA-fed tile kernels.  Measured on B200 (profiles/r02j_ntt_variants.txt):
// bit-identical results, but 0.85 -> 1.02 us per forward limb-NTT at N = 2^16 -- the tile kernel is bound by the
// latency of its per-thread twiddle loads, not by the tile load the copy engine hides, and the two staging buffers
// cost L1 capacity (192 KiB of shared memory per SM leave 28 KiB of L1 for 180 KiB of twiddles in flight).  Default off.
bool ntt_stream() {
    static const bool on = [] {
        const char* e = getenv("OFHE_NTT_STREAM");
        return e && e[0] == '1';
    }();
    return on;
}
template <int A>
void fwd_split(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, const PosMap& op, cudaStream_t st,
               const NttTail* tail = nullptr) {
    static const bool once = [] {
        OFHE_CUDA(cudaFuncSetAttribute(ntt_fwd_tile_stream<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)kTileBytes));
        return true;
    }();
    (void)once;
    const uint32_t chunk = ntt_chunk_limbs(r);
    for (uint32_t s = 0; s < lm.n; s += chunk) {
        LimbMap l2;
        PosMap o2;
        sub_maps(lm, op, s, std::min(chunk, lm.n - s), l2, o2);
        ntt_fwd_strided<A><<<dim3((r.N >> A) / 256, l2.n), 256, 0, st>>>(r.dev(), in, out, l2, o2);
        if (tail) {
            OFHE_REQUIRE(s == 0 && l2.n == lm.n, "fused transform tail: chunked launches are not supported");
            ntt_fwd_tile_tail<A><<<dim3(r.N / kTile, l2.n), 256, 0, st>>>(r.dev(), out, on_output(l2, o2), *tail);
        }
        else if (ntt_stream()) {
            const uint32_t items = l2.n << A;
            ntt_fwd_tile_stream<A><<<stream_grid(r, items), 256, 2 * kTileBytes, st>>>(r.dev(), out, on_output(l2, o2), items);
        }
        else
            ntt_fwd_tile<A><<<dim3(r.N / kTile, l2.n), 256, 0, st>>>(r.dev(), out, on_output(l2, o2));
    }
}
template <int A>
void inv_split(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, const PosMap& op,
               const uint64_t* fold, const PeerOut& po, cudaStream_t st) {
    static const bool once = [] {
        OFHE_CUDA(cudaFuncSetAttribute(ntt_inv_tile_stream<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (int)kTileBytes));
        return true;
    }();
    (void)once;
    const uint32_t chunk = ntt_chunk_limbs(r);
    for (uint32_t s = 0; s < lm.n; s += chunk) {
        LimbMap l2;
        PosMap o2;
        sub_maps(lm, op, s, std::min(chunk, lm.n - s), l2, o2);
        if (ntt_stream()) {
            const uint32_t items = l2.n << A;
            ntt_inv_tile_stream<A><<<stream_grid(r, items), 256, 2 * kTileBytes, st>>>(r.dev(), in, out, l2, o2, items);
        }
        else
            ntt_inv_tile<A><<<dim3(r.N / kTile, l2.n), 256, 0, st>>>(r.dev(), in, out, l2, o2);
        ntt_inv_strided<A><<<dim3((r.N >> A) / 256, l2.n), 256, 0, st>>>(r.dev(), out, on_output(l2, o2),
                                                                          fold ? fold + 4 * (size_t)s : nullptr, po);
    }
}

}  // namespace

// diagnostic: resident clusters the device can hold for the forward / inverse cluster kernel of ring dimension 2^logN
namespace {
template <int A>
void occ(int out[2]) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3(1u << A, 1024, 1);
    cfg.blockDim           = dim3(256, 1, 1);
    cudaLaunchAttribute at[1];
    at[0].id               = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1u << A;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs              = at;
    cfg.numAttrs           = 1;
    if (A > 3) {
        OFHE_CUDA(cudaFuncSetAttribute(ntt_fwd_cluster<A>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        OFHE_CUDA(cudaFuncSetAttribute(ntt_inv_cluster<A>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    const size_t smem = sizeof(uint64_t) * kTile * (A ? 2 : 1);
    OFHE_CUDA(cudaFuncSetAttribute(ntt_inv_cluster<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.dynamicSmemBytes = 0;
    OFHE_CUDA(cudaOccupancyMaxActiveClusters(&out[0], ntt_fwd_cluster<A>, &cfg));
    cfg.dynamicSmemBytes = smem;
    OFHE_CUDA(cudaOccupancyMaxActiveClusters(&out[1], ntt_inv_cluster<A>, &cfg));
}
}  // namespace
void ntt_cluster_occupancy(uint32_t logN, int out[2]) {
    out[0] = out[1] = 0;
    switch ((int)logN - 12) {
        case 1: occ<1>(out); break;
        case 2: occ<2>(out); break;
        case 3: occ<3>(out); break;
        case 4: occ<4>(out); break;
        default: break;
    }
}

bool ntt_fwd_tail_fused(const Ring& r) {
    if (r.logN < 13 || r.logN > 17)
        return false;
    // OFHE_NTT_TAIL=1 fuses the ModDown tail into the tile pass.  Bit-identical, but measured on B200 (N = 2^16, 2 x 45
    // limbs, profiles/r02u_*): the tile pass grows from 60 to 89 us -- each thread owns 16 CONTIGUOUS outputs, so the
    // minuend / addend loads are 16-byte pieces 128 bytes apart across a warp -- against 30 us for the two stand-alone
    // tail launches: 492 vs 497 us for one EvalMult alone, and 2038 vs 2127 mults/s with four in flight, where the
    // HBM-bound tail kernels of one multiplication hide under the integer-bound kernels of another.  Default off.
    static const bool on = [] {
        const char* e = getenv("OFHE_NTT_TAIL");
        return e && e[0] == '1';
    }();
    return on && (r.logN == 17 || ntt_mode(true, (int)r.logN - 12) == 1) && ntt_chunk_limbs(r) >= (uint32_t)kMaxLimbs;
}

void launch_ntt_fwd(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, cudaStream_t st,
                    const PosMap* out_pos, const NttTail* tail) {
    if (lm.n == 0)
        return;
    OFHE_REQUIRE(!tail || ntt_fwd_tail_fused(r), "fused transform tail is not available for this ring");
    const PosMap op = default_pos(lm, out_pos);
    if (r.logN < 12) {
        uint32_t th = r.N / 2 < 32 ? 32 : (r.N / 2 > 512 ? 512 : r.N / 2);
        ntt_small<false><<<lm.n, th, r.N * sizeof(uint64_t), st>>>(r.dev(), in, out, lm, op, nullptr, PeerOut{});
    }
    else {
        switch (r.logN - 12) {
            case 0: fwd_cluster<0>(r, in, out, lm, op, st); break;
            case 1: ntt_mode(true, 1) ? fwd_split<1>(r, in, out, lm, op, st, tail) : fwd_cluster<1>(r, in, out, lm, op, st); break;
            case 2: ntt_mode(true, 2) ? fwd_split<2>(r, in, out, lm, op, st, tail) : fwd_cluster<2>(r, in, out, lm, op, st); break;
            case 3: ntt_mode(true, 3) ? fwd_split<3>(r, in, out, lm, op, st, tail) : fwd_cluster<3>(r, in, out, lm, op, st); break;
            case 4: ntt_mode(true, 4) ? fwd_split<4>(r, in, out, lm, op, st, tail) : fwd_cluster<4>(r, in, out, lm, op, st); break;
            case 5: fwd_split<5>(r, in, out, lm, op, st, tail); break;
            default: throw Error("ring dimension above 2^17 is not supported");
        }
    }
    OFHE_CUDA(cudaGetLastError());
}

void launch_ntt_inv(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, cudaStream_t st,
                    const PosMap* out_pos, const uint64_t* fold, const PeerOut* peers) {
    if (lm.n == 0)
        return;
    const PosMap op  = default_pos(lm, out_pos);
    const PeerOut po = peers ? *peers : PeerOut{};
    if (r.logN < 12) {
        uint32_t th = r.N / 2 < 32 ? 32 : (r.N / 2 > 512 ? 512 : r.N / 2);
        ntt_small<true><<<lm.n, th, r.N * sizeof(uint64_t), st>>>(r.dev(), in, out, lm, op, fold, po);
    }
    else {
        switch (r.logN - 12) {
            case 0: inv_cluster<0>(r, in, out, lm, op, fold, po, st); break;
            case 1: ntt_mode(false, 1) ? inv_split<1>(r, in, out, lm, op, fold, po, st) : inv_cluster<1>(r, in, out, lm, op, fold, po, st); break;
            case 2: ntt_mode(false, 2) ? inv_split<2>(r, in, out, lm, op, fold, po, st) : inv_cluster<2>(r, in, out, lm, op, fold, po, st); break;
            case 3: ntt_mode(false, 3) ? inv_split<3>(r, in, out, lm, op, fold, po, st) : inv_cluster<3>(r, in, out, lm, op, fold, po, st); break;
            case 4: ntt_mode(false, 4) ? inv_split<4>(r, in, out, lm, op, fold, po, st) : inv_cluster<4>(r, in, out, lm, op, fold, po, st); break;
            case 5: inv_split<5>(r, in, out, lm, op, fold, po, st); break;
            default: throw Error("ring dimension above 2^17 is not supported");
        }
    }
    OFHE_CUDA(cudaGetLastError());
}

}  // namespace ofhe
