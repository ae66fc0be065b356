// ntt.cu -- negacyclic NTT / inverse NTT over [limbs x N] slabs of 64-bit residues, sm_100a.
//
// Replaces ChineseRemainderTransformFTTNat::{ForwardTransformToBitReverseInPlace,
// InverseTransformFromBitReverseInPlace} (src/core/include/math/hal/intnat/transformnat-impl.h:303-374,
// 512-625, 648-691) and the limb loop of DCRTPolyImpl::SwitchFormat (lattice/hal/default/dcrtpoly-impl.h:1932).
// Same contract: forward = natural order in, bit-reversed order out; inverse = the opposite; canonical
// residues in [0,q) on exit; twiddle tables T[brev(i)] = psi^i with Shoup constants.
//
// B200 decomposition (not the reference's stage-by-stage loop):
//   forward :  [strided kernel: first logN-12 stages, 2^(logN-12) points per thread, no shared memory]
//              [tile kernel   : last 12 stages on contiguous 4096-point tiles (32 KiB of shared memory),
//                               three radix-16 register rounds with two swizzled smem exchanges]
//   inverse :  the mirror image (tile kernel first, strided kernel last with N^-1 folded in).
// Butterflies are Harvey-lazy: forward values live in [0,4q), inverse values in [0,2q); one final
// conditional subtraction makes the output canonical, hence bit-identical to the reference.
// The intermediate between the two kernels stays L2-resident for slabs that fit the 126 MB L2.
#include "internal.h"

namespace ofhe {

namespace {
// NOTE: `in` and `out` may be the same slab (in-place transforms), so data pointers carry no __restrict__.

__device__ __forceinline__ Twiddle ld_tw(const Twiddle* p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    return Twiddle{v.x, v.y};
}

// Lazy-range discipline (Harvey).  On sm_100 the integer multiplier (IMAD.WIDE: 4 issue cycles per warp on the
// fmaheavy pipe, 2 for a 32-bit IMAD) is the bottleneck of a 60-bit butterfly: a Shoup product costs 6 IMAD.WIDE +
// 4 IMAD = 32 cycles, while the ALU pipe (IADD3 / ISETP / SEL, 2 cycles each) has ~16 instruction slots to spare in
// that time.  So range control is done with conditional subtractions on the ALU pipe (not with multiply-based
// reductions), one per butterfly:  forward values live in [0,4q), inverse values in [0,2q).
struct Mod {
    uint64_t q, nq, twoq, one_s;  // nq = 2^64 - q
};
__device__ __forceinline__ Mod make_mod(const ModConst& mc) {
    return Mod{mc.q, 0 - mc.q, 2 * mc.q, mc.one_s};
}
// Cooley-Tukey butterfly, inputs/outputs in [0,4q)
__device__ __forceinline__ void ct_bfly(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m) {
    const uint64_t x = X >= m.twoq ? X - m.twoq : X;        // [0,2q)
    const uint64_t v = shoup_lazy_n(Y, t.w, t.ws, m.nq);   // [0,2q)
    X                = x + v;
    Y                = x - v + m.twoq;
}
// Gentleman-Sande butterfly, inputs/outputs in [0,2q)
__device__ __forceinline__ void gs_bfly(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m) {
    const uint64_t s = X + Y;
    const uint64_t d = X - Y + m.twoq;
    X                = s >= m.twoq ? s - m.twoq : s;
    Y                = shoup_lazy_n(d, t.w, t.ws, m.nq);
}

// One stage (local depth D of R) on 2^R register-resident points; all indices are compile-time constants so the
// point array stays in registers.  x[k]: local index k, MSB = the round's first stage bit.
// tw_base = 2^s0 + u (s0 = the round's first global stage, u = the s0 index bits above this group).
template <int R, int D>
__device__ __forceinline__ void ct_stage(uint64_t (&x)[1 << R], const Twiddle* __restrict__ tw, uint32_t tw_base,
                                         const Mod& m) {
    constexpr int half = 1 << (R - 1 - D);
#pragma unroll
    for (int g = 0; g < (1 << D); ++g) {
        const Twiddle t = ld_tw(tw + ((tw_base << D) + g));
#pragma unroll
        for (int l = 0; l < half; ++l)
            ct_bfly(x[(g << (R - D)) + l], x[(g << (R - D)) + l + half], t, m);
    }
}
template <int R, int D>
__device__ __forceinline__ void gs_stage(uint64_t (&x)[1 << R], const Twiddle* __restrict__ tw, uint32_t tw_base,
                                         const Mod& m) {
    constexpr int half = 1 << (R - 1 - D);
#pragma unroll
    for (int g = 0; g < (1 << D); ++g) {
        const Twiddle t = ld_tw(tw + ((tw_base << D) + g));
#pragma unroll
        for (int l = 0; l < half; ++l)
            gs_bfly(x[(g << (R - D)) + l], x[(g << (R - D)) + l + half], t, m);
    }
}
// The four constants (nInv, Shoup(nInv), Tinv[1]*nInv, Shoup(..)) may carry an extra per-limb scalar folded in by
// the caller (e.g. QHatInvModq ahead of a base conversion): INTT(x) * s costs nothing extra.
struct LastConst {
    uint64_t a, as, b, bs;
};
__device__ __forceinline__ LastConst last_const(const ModConst& mc, const uint64_t* __restrict__ fold, uint32_t limb) {
    if (fold)
        return LastConst{fold[4 * limb], fold[4 * limb + 1], fold[4 * limb + 2], fold[4 * limb + 3]};
    return LastConst{mc.nInv, mc.nInvS, mc.w1nInv, mc.w1nInvS};
}
template <int R>
__device__ __forceinline__ void gs_last_stage(uint64_t (&x)[1 << R], const LastConst& lc, const Mod& m) {
    constexpr int half = 1 << (R - 1);
#pragma unroll
    for (int l = 0; l < half; ++l) {
        const uint64_t s = x[l] + x[l + half];
        const uint64_t e = x[l] - x[l + half] + m.twoq;
        x[l]             = shoup(s, lc.a, lc.as, m.q);
        x[l + half]      = shoup(e, lc.b, lc.bs, m.q);
    }
}

// R consecutive CT stages, [0,4q) -> [0,4q)
template <int R>
__device__ __forceinline__ void ct_round(uint64_t (&x)[1 << R], const Twiddle* __restrict__ tw, uint32_t tw_base,
                                         const Mod& m) {
    ct_stage<R, 0>(x, tw, tw_base, m);
    if constexpr (R > 1) ct_stage<R, 1>(x, tw, tw_base, m);
    if constexpr (R > 2) ct_stage<R, 2>(x, tw, tw_base, m);
    if constexpr (R > 3) ct_stage<R, 3>(x, tw, tw_base, m);
    if constexpr (R > 4) ct_stage<R, 4>(x, tw, tw_base, m);
}
// R consecutive GS stages (mirror order), [0,2q) -> [0,2q).  If LAST, depth 0 is the transform's final stage
// (canonical outputs).
template <int R, bool LAST>
__device__ __forceinline__ void gs_round(uint64_t (&x)[1 << R], const Twiddle* __restrict__ tw, uint32_t tw_base,
                                         const LastConst& lc, const Mod& m) {
    if constexpr (R > 4) gs_stage<R, 4>(x, tw, tw_base, m);
    if constexpr (R > 3) gs_stage<R, 3>(x, tw, tw_base, m);
    if constexpr (R > 2) gs_stage<R, 2>(x, tw, tw_base, m);
    if constexpr (R > 1) gs_stage<R, 1>(x, tw, tw_base, m);
    if constexpr (LAST)
        gs_last_stage<R>(x, lc, m);
    else
        gs_stage<R, 0>(x, tw, tw_base, m);
}

__device__ __forceinline__ uint64_t canon(uint64_t v, const Mod& m) {  // [0,4q) -> [0,q)
    v = v >= m.twoq ? v - m.twoq : v;
    return v >= m.q ? v - m.q : v;
}

// ------------------------------------------------------------------------------------------------
// Strided kernels: the A = logN - 12 stages whose butterfly span is >= 4096.  One thread owns the 2^A
// points {c + k * N/2^A}; a warp's loads are 256 contiguous bytes per k.  Twiddles T[1 .. 2^A) are
// warp-uniform.  No shared memory.
template <int A>
__global__ void __launch_bounds__(256) ntt_fwd_strided(RingDev r, const uint64_t* in,
                                                        uint64_t* out, LimbMap lm, PosMap op) {
    const uint32_t limb   = blockIdx.y;
    const uint32_t mi     = lm.idx[limb];
    const uint32_t stride = r.N >> A;
    const uint32_t c      = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= stride)
        return;
    const Mod m = make_mod(r.mc[mi]);
    const uint64_t* src = in + (size_t)lm.pos[limb] * r.N + c;
    uint64_t* dst       = out + (size_t)op.pos[limb] * r.N + c;
    uint64_t x[1 << A];
#pragma unroll
    for (int k = 0; k < (1 << A); ++k)
        x[k] = src[(size_t)k * stride];
    ct_round<A>(x, r.tw_fwd + (size_t)mi * r.N, 1u, m);
#pragma unroll
    for (int k = 0; k < (1 << A); ++k)
        dst[(size_t)k * stride] = x[k];  // lazy [0,4q): the tile kernel canonicalises
}

template <int A>
__global__ void __launch_bounds__(256) ntt_inv_strided(RingDev r, uint64_t* data, LimbMap lm,
                                                        const uint64_t* __restrict__ fold, PeerOut po) {
    const uint32_t limb   = blockIdx.y;
    const uint32_t mi     = lm.idx[limb];
    const uint32_t stride = r.N >> A;
    const uint32_t c      = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= stride)
        return;
    const ModConst mc = r.mc[mi];
    const Mod m       = make_mod(mc);
    uint64_t* p = data + (size_t)lm.pos[limb] * r.N + c;
    uint64_t x[1 << A];
#pragma unroll
    for (int k = 0; k < (1 << A); ++k)
        x[k] = p[(size_t)k * stride];
    gs_round<A, true>(x, r.tw_inv + (size_t)mi * r.N, 1u, last_const(mc, fold, limb), m);
    // every point passed through the folded last stage => already canonical
    if (po.n == 0) {
#pragma unroll
        for (int k = 0; k < (1 << A); ++k)
            p[(size_t)k * stride] = x[k];
    }
    else {  // push the finished rows into every destination buffer (local one included)
        const size_t off = (size_t)lm.pos[limb] * r.N + c;
        for (uint32_t d = 0; d < po.n; ++d) {
            uint64_t* q = po.base[d] + off;
#pragma unroll
            for (int k = 0; k < (1 << A); ++k)
                q[(size_t)k * stride] = x[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tile kernels: 12 stages on a contiguous tile of 4096 points, 256 threads x 16 points.
//   round A: tile bits 11..8   thread t holds c = t + 256k              (coalesced global access)
//   round B: tile bits  7..4   thread t holds c = (t>>4)*256 + (t&15) + 16k
//   round C: tile bits  3..0   thread t holds c = 16t + k               (128 contiguous bytes / thread)
// Exchange A<->B is conflict-free in linear layout; B<->C uses the swizzle c ^ ((c>>4)&15).
constexpr int kTile = 4096;
#ifndef OFHE_TILE_MINB
#define OFHE_TILE_MINB 3  // resident CTAs per SM the tile kernels are compiled for (register cap 65536/(256*MINB))
#endif
__device__ __forceinline__ uint32_t swz(uint32_t c) {
    return c ^ ((c >> 4) & 15u);
}

template <bool SINGLE>  // SINGLE: logN == 12, input is canonical and no strided kernel ran before
__global__ void __launch_bounds__(256, OFHE_TILE_MINB) ntt_fwd_tile(RingDev r, const uint64_t* in,
                                                     uint64_t* out, LimbMap lm, PosMap op) {
    __shared__ uint64_t sm[kTile];
    const uint32_t limb = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
    const uint32_t mi = lm.idx[limb];
    const Mod m         = make_mod(r.mc[mi]);
    const Twiddle* tw   = r.tw_fwd + (size_t)mi * r.N;
    const uint32_t s0   = r.logN - 12;  // first stage handled here
    const size_t base   = (size_t)lm.pos[limb] * r.N + (size_t)tile * kTile;
    const size_t obase  = (size_t)op.pos[limb] * r.N + (size_t)tile * kTile;
    uint64_t x[16];
    // ---- round A
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = in[base + t + 256 * k];
    ct_round<4>(x, tw, (1u << s0) + tile, m);
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sm[t + 256 * k] = x[k];
    __syncthreads();
    // ---- round B
    const uint32_t bh = t >> 4, bl = t & 15u;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = sm[bh * 256 + bl + 16 * k];
    ct_round<4>(x, tw, ((1u << s0) + tile) * 16u + bh, m);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sm[swz(bh * 256 + bl + 16 * k)] = x[k];
    __syncthreads();
    // ---- round C
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = sm[swz(16 * t + k)];
    ct_round<4>(x, tw, ((1u << s0) + tile) * 256u + t, m);
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + obase + 16 * t);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        dst[k] = make_ulonglong2(canon(x[2 * k], m), canon(x[2 * k + 1], m));
}

template <bool SINGLE>  // SINGLE: logN == 12, the folded last stage happens in round A
__global__ void __launch_bounds__(256, OFHE_TILE_MINB) ntt_inv_tile(RingDev r, const uint64_t* in,
                                                     uint64_t* out, LimbMap lm, PosMap op,
                                                     const uint64_t* __restrict__ fold, PeerOut po) {
    __shared__ uint64_t sm[kTile];
    const uint32_t limb = blockIdx.y, tile = blockIdx.x, t = threadIdx.x;
    const uint32_t mi = lm.idx[limb];
    const ModConst mc = r.mc[mi];
    const Mod m       = make_mod(mc);
    const Twiddle* tw = r.tw_inv + (size_t)mi * r.N;
    const uint32_t s0 = r.logN - 12;
    const size_t base = (size_t)lm.pos[limb] * r.N + (size_t)tile * kTile;
    const size_t obase = (size_t)op.pos[limb] * r.N + (size_t)tile * kTile;
    uint64_t x[16];
    // ---- round C (stages logN-1 .. logN-4)
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(in + base + 16 * t);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        ulonglong2 v = src[k];
        x[2 * k]     = v.x;
        x[2 * k + 1] = v.y;
    }
    const LastConst lc = last_const(mc, SINGLE ? fold : nullptr, limb);
    gs_round<4, false>(x, tw, ((1u << s0) + tile) * 256u + t, lc, m);  // canonical input
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sm[swz(16 * t + k)] = x[k];
    __syncthreads();
    // ---- round B
    const uint32_t bh = t >> 4, bl = t & 15u;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = sm[swz(bh * 256 + bl + 16 * k)];
    gs_round<4, false>(x, tw, ((1u << s0) + tile) * 16u + bh, lc, m);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k)
        sm[bh * 256 + bl + 16 * k] = x[k];
    __syncthreads();
    // ---- round A
#pragma unroll
    for (int k = 0; k < 16; ++k)
        x[k] = sm[t + 256 * k];
    gs_round<4, SINGLE>(x, tw, (1u << s0) + tile, lc, m);
    if (!SINGLE || po.n == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            out[obase + t + 256 * k] = x[k];  // SINGLE: canonical; else lazy [0,2q), finished by the strided kernel
    }
    else {
        for (uint32_t d = 0; d < po.n; ++d)
#pragma unroll
            for (int k = 0; k < 16; ++k)
                po.base[d][obase + t + 256 * k] = x[k];
    }
}

// ------------------------------------------------------------------------------------------------
// Small rings (N < 4096: the reference's unit-test sizes): one CTA per limb, whole limb in shared memory,
// one stage per barrier.  Not a performance path.
template <bool INV>
__global__ void ntt_small(RingDev r, const uint64_t* in, uint64_t* out, LimbMap lm, PosMap op,
                          const uint64_t* __restrict__ fold, PeerOut po) {
    extern __shared__ uint64_t s[];
    const uint32_t limb = lm.pos[blockIdx.x], olimb = op.pos[blockIdx.x], mi = lm.idx[blockIdx.x], N = r.N;
    const ModConst mc = r.mc[mi];
    const Mod m       = make_mod(mc);
    const uint64_t q = mc.q, twoq = 2 * q;
    const Twiddle* tw = (INV ? r.tw_inv : r.tw_fwd) + (size_t)mi * N;
    for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
        s[i] = in[(size_t)limb * N + i];
    __syncthreads();
    if (!INV) {
        for (uint32_t mm = 1, t = N >> 1; mm < N; mm <<= 1, t >>= 1) {
            for (uint32_t b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
                uint32_t i = b / t, j = 2 * i * t + (b % t);
                ct_bfly(s[j], s[j + t], ld_tw(tw + mm + i), m);
            }
            __syncthreads();
        }
        for (uint32_t i = threadIdx.x; i < N; i += blockDim.x)
            out[(size_t)olimb * N + i] = canon(s[i], m);
    }
    else {
        for (uint32_t mm = N >> 1, t = 1; mm > 1; mm >>= 1, t <<= 1) {
            for (uint32_t b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
                uint32_t i = b / t, j = 2 * i * t + (b % t);
                gs_bfly(s[j], s[j + t], ld_tw(tw + mm + i), m);
            }
            __syncthreads();
        }
        const uint32_t h  = N >> 1;
        const LastConst lc = last_const(mc, fold, blockIdx.x);
        for (uint32_t j = threadIdx.x; j < h; j += blockDim.x) {
            uint64_t a = s[j] + s[j + h], e = s[j] - s[j + h] + twoq;
            a = shoup(a, lc.a, lc.as, q);
            e = shoup(e, lc.b, lc.bs, q);
            for (uint32_t d = 0; d < (po.n ? po.n : 1u); ++d) {
                uint64_t* o = po.n ? po.base[d] : out;
                o[(size_t)olimb * N + j]     = a;
                o[(size_t)olimb * N + j + h] = e;
            }
        }
    }
}

// the second kernel of a pair works in place on the OUTPUT slab: same moduli, output rows
LimbMap on_output(const LimbMap& lm, const PosMap& op) {
    LimbMap o = lm;
    for (uint32_t i = 0; i < lm.n; ++i)
        o.pos[i] = op.pos[i];
    return o;
}
template <int A>
void fwd_big(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, const PosMap& op, cudaStream_t st) {
    const uint32_t stride = r.N >> A;
    dim3 g1((stride + 255) / 256, lm.n);
    ntt_fwd_strided<A><<<g1, 256, 0, st>>>(r.dev(), in, out, lm, op);
    dim3 g2(r.N / kTile, lm.n);
    ntt_fwd_tile<false><<<g2, 256, 0, st>>>(r.dev(), out, out, on_output(lm, op), op);
}
template <int A>
void inv_big(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, const PosMap& op,
             const uint64_t* fold, const PeerOut& po, cudaStream_t st) {
    dim3 g2(r.N / kTile, lm.n);
    ntt_inv_tile<false><<<g2, 256, 0, st>>>(r.dev(), in, out, lm, op, nullptr, PeerOut{});
    const uint32_t stride = r.N >> A;
    dim3 g1((stride + 255) / 256, lm.n);
    ntt_inv_strided<A><<<g1, 256, 0, st>>>(r.dev(), out, on_output(lm, op), fold, po);
}
PosMap default_pos(const LimbMap& lm, const PosMap* out_pos) {
    if (out_pos)
        return *out_pos;
    PosMap p;
    for (uint32_t i = 0; i < lm.n; ++i)
        p.pos[i] = lm.pos[i];
    return p;
}

}  // namespace

void launch_ntt_fwd(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, cudaStream_t st,
                    const PosMap* out_pos) {
    if (lm.n == 0)
        return;
    const PosMap op = default_pos(lm, out_pos);
    if (r.logN < 12) {
        uint32_t th = r.N / 2 < 32 ? 32 : (r.N / 2 > 512 ? 512 : r.N / 2);
        ntt_small<false><<<lm.n, th, r.N * sizeof(uint64_t), st>>>(r.dev(), in, out, lm, op, nullptr, PeerOut{});
    }
    else if (r.logN == 12) {
        ntt_fwd_tile<true><<<dim3(1, lm.n), 256, 0, st>>>(r.dev(), in, out, lm, op);
    }
    else {
        switch (r.logN - 12) {
            case 1: fwd_big<1>(r, in, out, lm, op, st); break;
            case 2: fwd_big<2>(r, in, out, lm, op, st); break;
            case 3: fwd_big<3>(r, in, out, lm, op, st); break;
            case 4: fwd_big<4>(r, in, out, lm, op, st); break;
            case 5: fwd_big<5>(r, in, out, lm, op, st); break;
            default: throw Error("ring dimension above 2^17 is not supported");
        }
    }
    OFHE_CUDA(cudaGetLastError());
}

void launch_ntt_inv(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, cudaStream_t st,
                    const PosMap* out_pos, const uint64_t* fold, const PeerOut* peers) {
    if (lm.n == 0)
        return;
    const PosMap op = default_pos(lm, out_pos);
    const PeerOut po = peers ? *peers : PeerOut{};
    if (r.logN < 12) {
        uint32_t th = r.N / 2 < 32 ? 32 : (r.N / 2 > 512 ? 512 : r.N / 2);
        ntt_small<true><<<lm.n, th, r.N * sizeof(uint64_t), st>>>(r.dev(), in, out, lm, op, fold, po);
    }
    else if (r.logN == 12) {
        ntt_inv_tile<true><<<dim3(1, lm.n), 256, 0, st>>>(r.dev(), in, out, lm, op, fold, po);
    }
    else {
        switch (r.logN - 12) {
            case 1: inv_big<1>(r, in, out, lm, op, fold, po, st); break;
            case 2: inv_big<2>(r, in, out, lm, op, fold, po, st); break;
            case 3: inv_big<3>(r, in, out, lm, op, fold, po, st); break;
            case 4: inv_big<4>(r, in, out, lm, op, fold, po, st); break;
            case 5: inv_big<5>(r, in, out, lm, op, fold, po, st); break;
            default: throw Error("ring dimension above 2^17 is not supported");
        }
    }
    OFHE_CUDA(cudaGetLastError());
}

}  // namespace ofhe
