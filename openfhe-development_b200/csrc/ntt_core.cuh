// ntt_core.cuh -- register-level building blocks of the negacyclic NTT kernels (ntt.cu) for sm_100a.
//
// Contract of the transforms built from these blocks: ChineseRemainderTransformFTTNat::
// ForwardTransformToBitReverseInPlace / InverseTransformFromBitReverseInPlace
// (src/core/include/math/hal/intnat/transformnat-impl.h:303-374, 512-625): canonical residues in, canonical
// residues out, tables T[brev(i)] = psi^i with Shoup constants.  Everything in between is this file's business.
//
// Instruction diet (the 60-bit butterfly is bound by the integer multiplier pipe):
//   * the Shoup quotient is taken from the three upper 32x32 partial products only (under-estimates the exact
//     floor by at most 2), so a twiddle product costs 5 IMAD.WIDE + 4 IMAD instead of 6 + 4 (+ carry chain);
//     the result is then in [0, 4q) instead of [0, 2q);
//   * forward (Cooley-Tukey) values live in [0, 16q) (q < 2^60): a conditional subtraction of 8q is inserted only
//     on the stages where the compile-time range bound would overflow (7 of 16 stages at N = 2^16);
//   * inverse (Gentleman-Sande) values live in [0, 4q);
//   * conditional subtractions are 4 instructions (sub.cc / subc.cc, borrow straight into the select predicate).
// Exactness: all quantities are integers below 2^64 by the bounds proved next to each routine, the last stage
// canonicalises, hence outputs are bit-identical to the reference's.
#pragma once
#include "internal.h"

namespace ofhe {
namespace nttc {

__device__ __forceinline__ Twiddle ld_tw(const Twiddle* p) {
    ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    return Twiddle{v.x, v.y};
}

// Twiddles of a radix-16 round (tw_base as in ct_stage / gs_stage) into L1, one round ahead of their use: every
// stage otherwise starts with a load that misses L1 (each tile touches its own table rows) and the round becomes a
// chain of exposed L2 latencies (profiles/r02e_split_summary.csv: long-scoreboard stalls 4 per issued instruction).
// Stage D reads the 2^D consecutive entries from (tw_base << D): 16, 32, 64, 128 contiguous bytes.
#ifndef OFHE_NTT_PREFETCH
#define OFHE_NTT_PREFETCH 1
#endif
__device__ __forceinline__ void prefetch_round(const Twiddle* tw, uint32_t tw_base) {
#if OFHE_NTT_PREFETCH
#pragma unroll
    for (int D = 0; D < 4; ++D)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tw + ((size_t)tw_base << D)));
#endif
}

struct Mod {
    uint64_t q, nq, c2, c4, c8;  // nq = 2^64 - q, ck = k*q
};
__device__ __forceinline__ Mod make_mod(const ModConst& mc) {
    return Mod{mc.q, 0 - mc.q, 2 * mc.q, 4 * mc.q, 8 * mc.q};
}

// x >= c ? x - c : x in four instructions (IADD3, IADD3.X with the borrow as predicate, 2 SEL)
__device__ __forceinline__ uint64_t csub4i(uint64_t X, uint64_t c) {
    uint32_t xl = (uint32_t)X, xh = (uint32_t)(X >> 32), cl = (uint32_t)c, ch = (uint32_t)(c >> 32), tl, th, b;
    asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;"
        : "=r"(tl), "=r"(th), "=r"(b)
        : "r"(xl), "r"(xh), "r"(cl), "r"(ch));
    tl = b ? xl : tl;
    th = b ? xh : th;
    return ((uint64_t)th << 32) | tl;
}

// floor(y * s / 2^64) - e with e in {0,1,2}: the lowest partial product and the carry of the middle column dropped
__device__ __forceinline__ uint64_t mulhi_t(uint64_t y, uint64_t s) {
    const uint32_t y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32), s0 = (uint32_t)s, s1 = (uint32_t)(s >> 32);
    const uint64_t p01 = (uint64_t)y0 * s1, p10 = (uint64_t)y1 * s0, p11 = (uint64_t)y1 * s1;
    return p11 + (p01 >> 32) + (p10 >> 32);
}
// y * w mod q up to a multiple of q: result in [0, q * (3 + y / 2^64)) for ANY 64-bit y  (exact quotient h gives
// y*w - h*q in [0, q * (1 + y / 2^64)); the truncated one adds at most 2q)
#ifndef OFHE_NTT_PTX
#define OFHE_NTT_PTX 1
#endif
__device__ __forceinline__ uint64_t shoup_t(uint64_t y, uint64_t w, uint64_t ws, uint64_t nq) {
#if OFHE_NTT_PTX
    // the same arithmetic spelled out: ptxas then keeps the quotient columns in one 3-input carry chain instead of
    // materialising zero-extended 64-bit halves (27.5 vs 29.5 cycles per warp-butterfly, profiles/r02a_ubench_bfly.txt)
    uint64_t r;
    asm("{\n\t"
        ".reg .u64 p01, p10, p11, a;\n\t"
        ".reg .u32 y0, y1, w0, w1, s0, s1, n0, n1, p01l, p01h, p10l, p10h, p11l, p11h, h0, h1, al, ah;\n\t"
        "mov.b64 {y0, y1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {s0, s1}, %3;\n\t"
        "mov.b64 {n0, n1}, %4;\n\t"
        "mul.wide.u32 p01, y0, s1;\n\t"
        "mul.wide.u32 p10, y1, s0;\n\t"
        "mul.wide.u32 p11, y1, s1;\n\t"
        "mov.b64 {p01l, p01h}, p01;\n\t"
        "mov.b64 {p10l, p10h}, p10;\n\t"
        "mov.b64 {p11l, p11h}, p11;\n\t"
        "add.cc.u32 h0, p11l, p01h;\n\t"
        "addc.u32 h1, p11h, 0;\n\t"
        "add.cc.u32 h0, h0, p10h;\n\t"
        "addc.u32 h1, h1, 0;\n\t"
        "mul.wide.u32 a, y0, w0;\n\t"
        "mov.b64 {al, ah}, a;\n\t"
        "mad.lo.u32 ah, y0, w1, ah;\n\t"
        "mad.lo.u32 ah, y1, w0, ah;\n\t"
        "mov.b64 a, {al, ah};\n\t"
        "mad.wide.u32 a, h0, n0, a;\n\t"
        "mov.b64 {al, ah}, a;\n\t"
        "mad.lo.u32 ah, h0, n1, ah;\n\t"
        "mad.lo.u32 ah, h1, n0, ah;\n\t"
        "mov.b64 %0, {al, ah};\n\t"
        "}"
        : "=l"(r)
        : "l"(y), "l"(w), "l"(ws), "l"(nq));
    return r;
#else
    return y * w + mulhi_t(y, ws) * nq;
#endif
}
// c + (y * w mod q up to a multiple of q): the addend rides on the first wide product's 64-bit accumulator, so the
// butterfly's X + v costs no instruction of its own (the sum must stay below 2^64: callers' range bounds)
#ifndef OFHE_NTT_ACC
#define OFHE_NTT_ACC 0
#endif
__device__ __forceinline__ uint64_t shoup_t_acc(uint64_t y, uint64_t w, uint64_t ws, uint64_t nq, uint64_t c) {
    uint64_t r;
    asm("{\n\t"
        ".reg .u64 p01, p10, p11, a;\n\t"
        ".reg .u32 y0, y1, w0, w1, s0, s1, n0, n1, p01l, p01h, p10l, p10h, p11l, p11h, h0, h1, al, ah;\n\t"
        "mov.b64 {y0, y1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {s0, s1}, %3;\n\t"
        "mov.b64 {n0, n1}, %4;\n\t"
        "mul.wide.u32 p01, y0, s1;\n\t"
        "mul.wide.u32 p10, y1, s0;\n\t"
        "mul.wide.u32 p11, y1, s1;\n\t"
        "mov.b64 {p01l, p01h}, p01;\n\t"
        "mov.b64 {p10l, p10h}, p10;\n\t"
        "mov.b64 {p11l, p11h}, p11;\n\t"
        "add.cc.u32 h0, p11l, p01h;\n\t"
        "addc.u32 h1, p11h, 0;\n\t"
        "add.cc.u32 h0, h0, p10h;\n\t"
        "addc.u32 h1, h1, 0;\n\t"
        "mad.wide.u32 a, y0, w0, %5;\n\t"
        "mov.b64 {al, ah}, a;\n\t"
        "mad.lo.u32 ah, y0, w1, ah;\n\t"
        "mad.lo.u32 ah, y1, w0, ah;\n\t"
        "mov.b64 a, {al, ah};\n\t"
        "mad.wide.u32 a, h0, n0, a;\n\t"
        "mov.b64 {al, ah}, a;\n\t"
        "mad.lo.u32 ah, h0, n1, ah;\n\t"
        "mad.lo.u32 ah, h1, n0, ah;\n\t"
        "mov.b64 %0, {al, ah};\n\t"
        "}"
        : "=l"(r)
        : "l"(y), "l"(w), "l"(ws), "l"(nq), "l"(c));
    return r;
}
// exact Harvey product, result in [0, 2q) for any y
__device__ __forceinline__ uint64_t shoup2(uint64_t y, uint64_t w, uint64_t ws, uint64_t nq) {
    return y * w + __umul64hi(y, ws) * nq;
}

// ---- forward (Cooley-Tukey) -------------------------------------------------------------------------------
// Range bound (in units of q, exclusive) of every point before stage s of a 2^logn transform.  A stage maps
// inputs below B*q (B <= 12) to outputs below (B+4)*q:  v = shoup_t(Y) < 4q,  X' = X + v,  Y' = X - v + 4q > 0.
// Before a stage whose input bound exceeds its limit the X operand is reduced (8q, then 4q if still needed); the
// last stage takes X < 4q so that its outputs are below 8q and three subtractions canonicalise.
__host__ __device__ constexpr int fwd_lim(int logn, int s) {
    return s == logn - 1 ? 4 : 12;
}
__host__ __device__ constexpr int fwd_bound(int logn, int s) {  // bound of the inputs of stage s (before reduction)
    int b = 1;
    for (int i = 0; i < s; ++i) {
        const int l = fwd_lim(logn, i);
        if (b > l && b > 8)
            b = 8;
        if (b > l)
            b = 4;
        b += 4;
    }
    return b;
}
template <int LOGN, int S>
__device__ __forceinline__ void ct_bfly(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m) {
    constexpr int b = fwd_bound(LOGN, S), lim = fwd_lim(LOGN, S);
    // NOTE: only X needs the reduction -- Y enters the product, which accepts any 64-bit value
    uint64_t x = X;
    if constexpr (b > lim && b > 8)
        x = csub4i(x, m.c8);
    if constexpr ((b > 8 ? 8 : b) > lim)
        x = csub4i(x, m.c4);
#if OFHE_NTT_ACC
    const uint64_t xv = shoup_t_acc(Y, t.w, t.ws, m.nq, x);  // x + v
    Y                 = (x + x + m.c4) - xv;                  // x - v + 4q
    X                 = xv;
#else
    const uint64_t v = shoup_t(Y, t.w, t.ws, m.nq);
    X                = x + v;
    Y                = x - v + m.c4;
#endif
}
// outputs of the last stage (below 8q) -> [0,q)
__device__ __forceinline__ uint64_t canon8(uint64_t v, const Mod& m) {
    return csub4i(csub4i(csub4i(v, m.c4), m.c2), m.q);
}

// One stage (local depth D of R, global stage S0 + D) on the 2^R register-resident points x[OFF .. OFF + 2^R);
// all indices are compile-time constants.  Local index MSB = the round's first stage bit.
// tw_base = 2^S0 + u (u = the S0 index bits above this group).
template <int LOGN, int S0, int R, int D, int OFF>
__device__ __forceinline__ void ct_stage(uint64_t* x, const Twiddle* __restrict__ tw, uint32_t tw_base, const Mod& m) {
    constexpr int half = 1 << (R - 1 - D);
#pragma unroll
    for (int g = 0; g < (1 << D); ++g) {
        const Twiddle t = ld_tw(tw + ((tw_base << D) + g));
#pragma unroll
        for (int l = 0; l < half; ++l)
            ct_bfly<LOGN, S0 + D>(x[OFF + (g << (R - D)) + l], x[OFF + (g << (R - D)) + l + half], t, m);
    }
}
template <int LOGN, int S0, int R, int OFF = 0>
__device__ __forceinline__ void ct_round(uint64_t* x, const Twiddle* __restrict__ tw, uint32_t tw_base, const Mod& m) {
    if constexpr (R > 0) ct_stage<LOGN, S0, R, 0, OFF>(x, tw, tw_base, m);
    if constexpr (R > 1) ct_stage<LOGN, S0, R, 1, OFF>(x, tw, tw_base, m);
    if constexpr (R > 2) ct_stage<LOGN, S0, R, 2, OFF>(x, tw, tw_base, m);
    if constexpr (R > 3) ct_stage<LOGN, S0, R, 3, OFF>(x, tw, tw_base, m);
    if constexpr (R > 4) ct_stage<LOGN, S0, R, 4, OFF>(x, tw, tw_base, m);
}

// Radix-16 rounds in DEPTH-FIRST order with a rolling twiddle window.
// A radix-16 round is a binary tree of 15 nodes (depth D = 0..3, 2^D nodes per depth, node = the butterflies that
// share one twiddle: 8, 4, 2, 1 of them).  Walking the tree depth-first instead of stage by stage keeps the set of
// twiddles needed next small, so the loads can run kWin nodes (>= 100 instructions) AHEAD of their use inside the
// same warp without exceeding the 80-register budget: the first kWin are issued before the block barrier that
// precedes the round (the 16 points are in shared memory then, their registers free), the others as the window
// rolls.  Stage-by-stage order made every stage start with a load issued a few instructions before its first use:
// the ncu source page of the round-2 kernel has 70 % of its long-scoreboard stalls on the first product behind a
// twiddle load (profiles/r02h_tile_stalls.txt).  asm volatile keeps the loads in program order.
__device__ __forceinline__ Twiddle ld_tw_pinned(const Twiddle* p) {
    Twiddle t;
#if defined(OFHE_NTT_EXP) && OFHE_NTT_EXP == 4  // TIMING EXPERIMENT ONLY (wrong results): every twiddle from one cache line
    p = reinterpret_cast<const Twiddle*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)0xFFFFF) + ((reinterpret_cast<uintptr_t>(p) >> 4) & 7);
#endif
    asm volatile("ld.global.nc.v2.u64 {%0, %1}, [%2];" : "=l"(t.w), "=l"(t.ws) : "l"(p) : "memory");
    return t;
}
#ifndef OFHE_NTT_DF_FWD
#define OFHE_NTT_DF_FWD 1  // 0: stage-by-stage order with just-in-time twiddle loads (measurement switch)
#endif
#ifndef OFHE_NTT_DF_INV
#define OFHE_NTT_DF_INV 1
#endif
#ifndef OFHE_NTT_WIN
#define OFHE_NTT_WIN 4
#endif
constexpr int kWin = OFHE_NTT_WIN;
struct TwWin {
    Twiddle t[kWin];
};
// forward (pre-order) and inverse (post-order) walks: depth and node-within-depth of step i
__device__ constexpr int kCtD[15] = {0, 1, 2, 3, 3, 2, 3, 3, 1, 2, 3, 3, 2, 3, 3};
__device__ constexpr int kCtG[15] = {0, 0, 0, 0, 1, 1, 2, 3, 1, 2, 4, 5, 3, 6, 7};
__device__ constexpr int kGsD[15] = {3, 3, 2, 3, 3, 2, 1, 3, 3, 2, 3, 3, 2, 1, 0};
__device__ constexpr int kGsG[15] = {0, 1, 0, 2, 3, 1, 0, 4, 5, 2, 6, 7, 3, 1, 0};
template <bool INV>
__device__ __forceinline__ void tw_preload(TwWin& w, const Twiddle* tw, uint32_t tw_base) {
    if (INV ? !OFHE_NTT_DF_INV : !OFHE_NTT_DF_FWD)
        return;
#pragma unroll
    for (int i = 0; i < kWin; ++i)
        w.t[i] = ld_tw_pinned(tw + ((tw_base << (INV ? kGsD[i] : kCtD[i])) + (INV ? kGsG[i] : kCtG[i])));
}
template <int LOGN, int S0, int OFF, int I>
__device__ __forceinline__ void ct_df_step(uint64_t* x, TwWin& w, const Twiddle* __restrict__ tw, uint32_t tw_base,
                                           const Mod& m) {
    if constexpr (I < 15) {
        constexpr int D = kCtD[I], g = kCtG[I], half = 1 << (3 - D);
        const Twiddle t = w.t[I % kWin];
        if constexpr (I + kWin < 15)
            w.t[I % kWin] = ld_tw_pinned(tw + ((tw_base << kCtD[I + kWin]) + kCtG[I + kWin]));
#pragma unroll
        for (int l = 0; l < half; ++l)
            ct_bfly<LOGN, S0 + D>(x[OFF + (g << (4 - D)) + l], x[OFF + (g << (4 - D)) + l + half], t, m);
        ct_df_step<LOGN, S0, OFF, I + 1>(x, w, tw, tw_base, m);
    }
}
// the window must hold the first kWin twiddles of the walk (tw_preload<false> with the same tw_base)
template <int LOGN, int S0, int OFF = 0>
__device__ __forceinline__ void ct_round16_df(uint64_t* x, TwWin& w, const Twiddle* __restrict__ tw, uint32_t tw_base,
                                              const Mod& m) {
#if OFHE_NTT_DF_FWD
    ct_df_step<LOGN, S0, OFF, 0>(x, w, tw, tw_base, m);
#else
    ct_round<LOGN, S0, 4, OFF>(x, tw, tw_base, m);
#endif
}

// ---- inverse (Gentleman-Sande) ----------------------------------------------------------------------------
// inputs/outputs in [0,4q):  s = X + Y < 8q -> one conditional subtraction;  d = X - Y + 4q in (0, 8q) and
// shoup_t(d) < q * (3 + 8q / 2^64) < 4q.
__device__ __forceinline__ void gs_bfly(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m) {
    const uint64_t s = X + Y;
    const uint64_t d = X - Y + m.c4;
    X                = csub4i(s, m.c4);
    Y                = shoup_t(d, t.w, t.ws, m.nq);
}
template <int R, int D, int OFF>
__device__ __forceinline__ void gs_stage(uint64_t* x, const Twiddle* __restrict__ tw, uint32_t tw_base, const Mod& m) {
    constexpr int half = 1 << (R - 1 - D);
#pragma unroll
    for (int g = 0; g < (1 << D); ++g) {
        const Twiddle t = ld_tw(tw + ((tw_base << D) + g));
#pragma unroll
        for (int l = 0; l < half; ++l)
            gs_bfly(x[OFF + (g << (R - D)) + l], x[OFF + (g << (R - D)) + l + half], t, m);
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
// final stage with N^-1 folded in (transformnat-impl.h:538-539, 623-624): exact products, canonical outputs
template <int R, int OFF>
__device__ __forceinline__ void gs_last_stage(uint64_t* x, const LastConst& lc, const Mod& m) {
    constexpr int half = 1 << (R - 1);
#pragma unroll
    for (int l = 0; l < half; ++l) {
        const uint64_t s = x[OFF + l] + x[OFF + l + half];           // < 8q
        const uint64_t e = x[OFF + l] - x[OFF + l + half] + m.c4;    // (0, 8q)
        x[OFF + l]        = csub4i(shoup2(s, lc.a, lc.as, m.nq), m.q);
        x[OFF + l + half] = csub4i(shoup2(e, lc.b, lc.bs, m.nq), m.q);
    }
}
// R consecutive GS stages (mirror order).  If LAST, depth 0 is the transform's final stage (canonical outputs).
template <int R, bool LAST, int OFF = 0>
__device__ __forceinline__ void gs_round(uint64_t* x, const Twiddle* __restrict__ tw, uint32_t tw_base,
                                         const LastConst& lc, const Mod& m) {
    if constexpr (R > 4) gs_stage<R, 4, OFF>(x, tw, tw_base, m);
    if constexpr (R > 3) gs_stage<R, 3, OFF>(x, tw, tw_base, m);
    if constexpr (R > 2) gs_stage<R, 2, OFF>(x, tw, tw_base, m);
    if constexpr (R > 1) gs_stage<R, 1, OFF>(x, tw, tw_base, m);
    if constexpr (R > 0) {
        if constexpr (LAST)
            gs_last_stage<R, OFF>(x, lc, m);
        else
            gs_stage<R, 0, OFF>(x, tw, tw_base, m);
    }
}

// inverse radix-16 round, depth-first (post-order) with the rolling twiddle window (see ct_round16_df)
template <bool LAST, int OFF, int I>
__device__ __forceinline__ void gs_df_step(uint64_t* x, TwWin& w, const Twiddle* __restrict__ tw, uint32_t tw_base,
                                           const LastConst& lc, const Mod& m) {
    if constexpr (I < 15) {
        constexpr int D = kGsD[I], g = kGsG[I], half = 1 << (3 - D);
        if constexpr (LAST && I == 14) {
            gs_last_stage<4, OFF>(x, lc, m);
        }
        else {
            const Twiddle t = w.t[I % kWin];
            if constexpr (I + kWin < (LAST ? 14 : 15))
                w.t[I % kWin] = ld_tw_pinned(tw + ((tw_base << kGsD[I + kWin]) + kGsG[I + kWin]));
#pragma unroll
            for (int l = 0; l < half; ++l)
                gs_bfly(x[OFF + (g << (4 - D)) + l], x[OFF + (g << (4 - D)) + l + half], t, m);
        }
        gs_df_step<LAST, OFF, I + 1>(x, w, tw, tw_base, lc, m);
    }
}
template <bool LAST, int OFF = 0>
__device__ __forceinline__ void gs_round16_df(uint64_t* x, TwWin& w, const Twiddle* __restrict__ tw, uint32_t tw_base,
                                              const LastConst& lc, const Mod& m) {
#if OFHE_NTT_DF_INV
    gs_df_step<LAST, OFF, 0>(x, w, tw, tw_base, lc, m);
#else
    gs_round<4, LAST, OFF>(x, tw, tw_base, lc, m);
#endif
}

// ---- cluster / distributed shared memory primitives (sm_90+) ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {  // same offset in CTA `rank`'s window
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, uint64_t v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void cluster_arrive_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- asynchronous staging: mbarrier + 1-D TMA bulk copies (UBLKCP) and 16-byte cp.async (LDGSTS) ------------
// The streaming tile kernels keep two 32 KiB tiles per CTA in flight: the copy engine fills one while the warps
// transform the other, so no warp ever waits on a global load with its registers allocated.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// spins (hardware-suspended try_wait) until the phase with the given parity has completed; a copy that never lands
// (a bug) traps after ~2 s instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done    = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (!done && spins > 400)
            __trap();
    }
}
// global -> shared bulk copy of `bytes` (multiple of 16, both sides 16-byte aligned), completion on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
// this thread's earlier cp.async copies arrive on `bar` when they have landed (counts towards the expected arrivals)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace nttc
}  // namespace ofhe
