// baseconv.cu -- RNS fast base conversion (ApproxSwitchCRTBasis) as one fused per-coefficient kernel.
//
// Replaces DCRTPolyImpl::ApproxSwitchCRTBasis (lattice/hal/default/dcrtpoly-impl.h:888-932, int128 branch
// :895-918): for every coefficient r
//     y_i   = x_i[r] * QHatInvModq_i  mod q_i                       (Shoup)
//     out_j = ( sum_i y_i * QHatModp[i][j] )  mod p_j               (128-bit accumulate + Barrett)
// Results are the same canonical residues as the reference's Mul128 / BarrettUint128ModUint64 path.
//
// GPU formulation.  grid = (coefficient blocks, output groups, batch): a thread owns one coefficient and a group of
// up to JPT = 16 target limbs (a 15 -> 45 conversion of 65536 coefficients runs on 3 x 65536 threads); the group loop
// stays rolled, two outputs per iteration, and its trip count follows the real number of outputs, so the prologue
// (16 source loads, the 30-bit split, the shared-memory fills: a third of the run time at JPT = 8 by ncu's sampling)
// is amortised without padding waste.  Every 128-bit shared-memory read of the conversion matrix is a warp-wide
// broadcast and feeds 8 IMAD.WIDE.  The source residues arrive already multiplied by QHatInvModq when the caller
// folded that constant into the last stage of the inverse NTT (`premultiplied`), else the kernel does the Shoup
// multiply itself.  The 60x60-bit multiply-accumulate works on 30-bit halves so that each MAC is exactly four
// IMAD.WIDE.U32 with NO carry chain:
//     y = y1*2^30 + y0,  m = m1*2^30 + m0   ->   c0 += y0*m0; c1a += y0*m1; c1b += y1*m0; c2 += y1*m1
// Each partial sum stays below 2^64 for 16 terms.  For up to 16 source limbs (every conversion of config 2) the MAC is
// three products instead (Karatsuba on the halves: the middle column is c0 + c2 - sum (y1-y0)(m1-m0), a signed
// IMAD.WIDE whose 16 terms are split over two int64 accumulators) -- the multiplier pipe is the kernel's bound.
// The four column sums are then assembled into one 128-bit integer
// S = c0 + (c1a + c1b) 2^30 + c2 2^60 (shifts and adds only) and reduced by a single 128 -> 64 Barrett step per output
// (~14 multiplier instructions; reducing the four columns separately with Shoup-lazy products took ~32).  The matrix
// slice of the CTA's output group is zero-padded in shared memory to 16*CH x JPT, so the MAC loops have compile-time
// bounds and no predication.  Requires q_i, p_j < 2^60 (MAX_MODULUS_SIZE, core/include/math/hal/basicint.h:48).
#include <cstdlib>

#include "internal.h"

namespace ofhe {
namespace {

constexpr uint32_t kMask30 = (1u << 30) - 1;
#ifndef OFHE_BCONV_JPT
#define OFHE_BCONV_JPT 16
#endif
#ifndef OFHE_BCONV_KARA
#define OFHE_BCONV_KARA 1  // 1: three products per 60x60 MAC (Karatsuba on the 30-bit halves) for <= 16 source limbs, mode 0
#endif
#ifndef OFHE_BCONV_MINB
#define OFHE_BCONV_MINB 1
#endif
constexpr int kJPT         = OFHE_BCONV_JPT;  // target limbs per thread
constexpr int kThreads     = 128;  // coefficients per CTA

struct BConvDev {
    uint32_t sizeQ, sizeP;
    const uint64_t* qhatinv;  // [sizeQ][2]  (w, Shoup(w))
    const uint2* mat;         // [sizeQ][sizeP]  (lo30, hi30)
    int mode;                 // 0 approx (a10), 1 exact (a12), 2 scale-and-round (a13)
    const double* dbl;        // [sizeQ]  qInv (mode 1) / tOSHatInvModsDivsFrac (mode 2)
    const uint64_t* alpha;    // [sizeQ+1][sizeP]  alphaQModp (mode 1)
    const uint64_t* extra;    // [sizeP]  tOSHatInvModsDivsModo[j][sizeI] (mode 2)
};

// (unsigned __int128) of a double above 2^63 (the reference's static_cast<DoubleNativeInt>(nu), dcrtpoly-impl.h:1569)
__device__ __forceinline__ void double_to_u128(double d, uint64_t& hi, uint64_t& lo) {
    const uint64_t bits = (uint64_t)__double_as_longlong(d);
    const uint64_t mant = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    const int sh        = (int)((bits >> 52) & 0x7ff) - 1075;  // value = mant * 2^sh, sh in [11, 75) here
    if (sh >= 64) {
        hi = mant << (sh - 64);
        lo = 0;
    }
    else {
        hi = sh ? mant >> (64 - sh) : 0;
        lo = mant << sh;
    }
}

__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}


__device__ __forceinline__ int64_t mad_wide_s(int32_t a, int32_t b, int64_t c) {
    int64_t d;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}

// CH chunks of 16 source limbs; kCPT coefficients per thread (2 when the register budget allows: every 128-bit
// matrix read from shared memory then feeds 16 IMADs)
template <int CH, int kCPT, int MODE>  // MODE is a template parameter so that the a10 instance carries no floating-point code
__global__ void __launch_bounds__(kThreads, OFHE_BCONV_MINB) bconv_kernel(RingDev r, BConvDev t, LimbMap qmap, LimbMap pmap,
                                                          PosMap xomap, const uint64_t* __restrict__ in,
                                                          uint64_t* __restrict__ out, size_t in_stride,
                                                          size_t out_stride, int premultiplied) {
    constexpr int MAXQ = 16 * CH;
    __shared__ __align__(16) uint2 s_mat[MAXQ * kJPT];
    constexpr bool KARA = (OFHE_BCONV_KARA != 0) && CH == 1 && kCPT == 1 && MODE == 0;
    __shared__ __align__(8) int32_t s_md[KARA ? MAXQ * kJPT : 2];  // m1 - m0 of every matrix entry (Karatsuba)
    __shared__ uint64_t s_src[MAXQ * 3];  // q, w, ws
    __shared__ uint64_t s_dst[kJPT * 4];  // p, floor(2^128/p) hi, lo, (pad)
    in += blockIdx.z * in_stride;
    out += blockIdx.z * out_stride;
    const uint32_t j0 = blockIdx.y * kJPT;
    for (uint32_t e = threadIdx.x; e < MAXQ * kJPT; e += kThreads) {
        const uint32_t i = e / kJPT, j = j0 + e % kJPT;
        s_mat[e] = (i < t.sizeQ && j < t.sizeP) ? t.mat[(size_t)i * t.sizeP + j] : make_uint2(0u, 0u);
        if constexpr (KARA)
            s_md[e] = (int32_t)s_mat[e].y - (int32_t)s_mat[e].x;
    }
    for (uint32_t i = threadIdx.x; i < MAXQ; i += kThreads) {
        const bool ok    = i < t.sizeQ;
        s_src[3 * i]     = ok ? r.mc[qmap.idx[i]].q : 1;
        s_src[3 * i + 1] = ok ? t.qhatinv[2 * i] : 0;
        s_src[3 * i + 2] = ok ? t.qhatinv[2 * i + 1] : 0;
    }
    if (threadIdx.x < kJPT) {
        static_assert(kJPT <= kThreads, "one thread per target limb fills s_dst");
        const uint32_t j   = j0 + threadIdx.x < t.sizeP ? j0 + threadIdx.x : t.sizeP - 1;
        const ModConst& mc = r.mc[pmap.idx[j]];
        uint64_t* d        = s_dst + 4 * threadIdx.x;
        d[0] = mc.q, d[1] = mc.mu128_hi, d[2] = mc.mu128_lo, d[3] = mc.one_s;
    }
    __syncthreads();

    const uint32_t N   = r.N;
    const uint32_t rr0 = blockIdx.x * (kThreads * kCPT) + threadIdx.x;  // coefficient c of this thread: rr0 + c*kThreads
    uint32_t y0[kCPT][MAXQ], y1[kCPT][MAXQ];
    int32_t yd[KARA ? MAXQ : 1];  // y1 - y0 (Karatsuba middle factor)
    double nu[kCPT];  // modes 1, 2: the floating-point overflow / rounding estimate, accumulated in index order with
                      // separately rounded multiply and add (no FMA contraction), exactly as the reference's loop
#pragma unroll
    for (int c = 0; c < kCPT; ++c) {
        const uint32_t rr = min(rr0 + c * kThreads, N - 1);  // clamped (branch-free); stores are predicated below
        nu[c]             = 0.5;
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) {
            const int ii = i < (int)t.sizeQ ? i : 0;  // padded rows re-read row 0 and are zeroed
            uint64_t y   = in[(size_t)qmap.pos[ii] * N + rr];
            if (!premultiplied && MODE != 2)
                y = shoup(y, s_src[3 * ii + 1], s_src[3 * ii + 2], s_src[3 * ii]);
            y        = i < (int)t.sizeQ ? y : 0;
            if (MODE != 0 && i < (int)t.sizeQ)
                nu[c] = __dadd_rn(nu[c], __dmul_rn(__ull2double_rn(y), t.dbl[i]));
            y0[c][i] = (uint32_t)y & kMask30;
            y1[c][i] = (uint32_t)(y >> 30);
            if constexpr (KARA)
                yd[i] = (int32_t)y1[c][i] - (int32_t)y0[c][i];
        }
    }
    // the output-pair loop stays rolled: the fully unrolled body (64 KB of code) thrashed the instruction cache
    const int nj = min(kJPT, (int)((t.sizeP - j0 + 1u) & ~1u));  // outputs of this group, rounded up to a pair
#pragma unroll 1
    for (int j = 0; j < nj; j += 2) {
        unsigned __int128 total[kCPT][2];
#pragma unroll
        for (int c = 0; c < kCPT; ++c)
            total[c][0] = total[c][1] = 0;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
            uint64_t a[kCPT][2][4];  // [coefficient][output j, j+1][c0, c1a, c1b, c2]
#pragma unroll
            for (int c = 0; c < kCPT; ++c)
#pragma unroll
                for (int o = 0; o < 2; ++o)
                    a[c][o][0] = a[c][o][1] = a[c][o][2] = a[c][o][3] = 0;
            if constexpr (KARA) {
                // three products per MAC:  y*m = y0 m0 + y1 m1 2^60 + (y0 m0 + y1 m1 - (y1-y0)(m1-m0)) 2^30.
                // a[.][0] = sum y0 m0, a[.][3] = sum y1 m1 (unsigned, < 2^64 for 16 terms); the signed middle products
                // (|.| < 2^60) go to a[.][1] for i < 8 and a[.][2] for i >= 8, each within int64.
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint4 m  = *reinterpret_cast<const uint4*>(&s_mat[i * kJPT + j]);
                    const int2 md  = *reinterpret_cast<const int2*>(&s_md[i * kJPT + j]);
                    const int mid  = i < 8 ? 1 : 2;
                    a[0][0][0]     = mad_wide(y0[0][i], m.x, a[0][0][0]);
                    a[0][0][3]     = mad_wide(y1[0][i], m.y, a[0][0][3]);
                    a[0][0][mid]   = (uint64_t)mad_wide_s(yd[i], md.x, (int64_t)a[0][0][mid]);
                    a[0][1][0]     = mad_wide(y0[0][i], m.z, a[0][1][0]);
                    a[0][1][3]     = mad_wide(y1[0][i], m.w, a[0][1][3]);
                    a[0][1][mid]   = (uint64_t)mad_wide_s(yd[i], md.y, (int64_t)a[0][1][mid]);
                }
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const uint64_t* a4 = a[0][o];
                    // middle column = A + C - D (non-negative); D = D_a + D_b signed
                    const unsigned __int128 mid = (unsigned __int128)a4[0] + a4[3] -
                                                  (unsigned __int128)((__int128)(int64_t)a4[1] + (int64_t)a4[2]);
                    total[0][o] += (unsigned __int128)a4[0] + (mid << 30) + (((unsigned __int128)a4[3]) << 60);
                }
                continue;
            }
#pragma unroll
            for (int i = 16 * ch; i < 16 * ch + 16; ++i) {
                const uint4 m = *reinterpret_cast<const uint4*>(&s_mat[i * kJPT + j]);  // (lo,hi) of j and of j+1
#pragma unroll
                for (int c = 0; c < kCPT; ++c) {
                    a[c][0][0] = mad_wide(y0[c][i], m.x, a[c][0][0]);
                    a[c][0][1] = mad_wide(y0[c][i], m.y, a[c][0][1]);
                    a[c][0][2] = mad_wide(y1[c][i], m.x, a[c][0][2]);
                    a[c][0][3] = mad_wide(y1[c][i], m.y, a[c][0][3]);
                    a[c][1][0] = mad_wide(y0[c][i], m.z, a[c][1][0]);
                    a[c][1][1] = mad_wide(y0[c][i], m.w, a[c][1][1]);
                    a[c][1][2] = mad_wide(y1[c][i], m.z, a[c][1][2]);
                    a[c][1][3] = mad_wide(y1[c][i], m.w, a[c][1][3]);
                }
            }
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int c = 0; c < kCPT; ++c) {
                    // the four column sums go into ONE 128-bit integer (shifts and adds, no multiplier):
                    // S = c0 + (c1a + c1b) 2^30 + c2 2^60 < 2^125 per chunk of 16 source limbs, CH <= 4 chunks
                    const uint64_t* a4 = a[c][o];
                    total[c][o] += (unsigned __int128)a4[0] + ((((unsigned __int128)a4[1]) + a4[2]) << 30) +
                                   (((unsigned __int128)a4[3]) << 60);
                }
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const uint64_t* d = s_dst + 4 * (j + o);
#pragma unroll
            for (int c = 0; c < kCPT; ++c) {
                const uint32_t rr = rr0 + c * kThreads;
                if (j0 + j + o < t.sizeP && rr < N) {
                    const uint32_t jj = j0 + j + o;
                    const uint64_t p  = d[0];
                    // one 128 -> 64 Barrett step per output (the reference's BarrettUint128ModUint64 value)
                    uint64_t v = barrett128((uint64_t)(total[c][o] >> 64), (uint64_t)total[c][o], p, d[1], d[2]);
                    if (MODE == 1) {  // dcrtpoly-impl.h:1060,1073: minus alphaQModp[(size_t)nu][j]
                        const uint64_t a = t.alpha[(size_t)__double2ull_rz(nu[c]) * t.sizeP + jj];
                        v                = submod(v, a, p);
                    }
                    else if (MODE == 2) {  // dcrtpoly-impl.h:1550-1586
                        const ModConst mc = r.mc[pmap.idx[jj]];
                        const uint64_t xo = in[(size_t)xomap.pos[jj] * N + rr];
                        v                 = addmod(v, mulmod(xo, t.extra[jj], mc), p);
                        uint64_t a;
                        if (fabs(nu[c]) <= 9223372036854774784.0)  // isConvertableToNativeInt (utils/utilities.h:100-126)
                            a = csub(lazy_reduce_n(__double2ull_rz(nu[c]), d[3], 0 - p), p);
                        else {
                            uint64_t hi, lo;
                            double_to_u128(nu[c], hi, lo);
                            a = barrett128(hi, lo, p, mc.mu128_hi, mc.mu128_lo);
                        }
                        v = addmod(v, a, p);
                    }
                    out[(size_t)pmap.pos[jj] * N + rr] = v;
                }
            }
        }
    }
}

// ---- more than 64 source limbs (the reference's depth-limit unit tests reach ~140: UnitTestBFVrnsCRTOperations.cpp:
// 142-193).  Not a performance path: one thread per (coefficient, target limb), run-time loop over the sources, 64x64
// products accumulated in 128 bits with a Barrett fold every 64 terms (64 * 2^120 < 2^127), the floating-point estimate
// in the same index order as the fast kernel.  Same epilogues, same canonical results.
template <int MODE>
__global__ void __launch_bounds__(kThreads) bconv_wide_kernel(RingDev r, BConvDev t, LimbMap qmap, LimbMap pmap, PosMap xomap,
                                                               const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                               size_t in_stride, size_t out_stride, int premultiplied) {
    in += blockIdx.z * in_stride;
    out += blockIdx.z * out_stride;
    const uint32_t N = r.N, rr = blockIdx.x * kThreads + threadIdx.x, jj = blockIdx.y;
    if (rr >= N)
        return;
    const ModConst mc = r.mc[pmap.idx[jj]];
    const uint64_t p  = mc.q;
    unsigned __int128 acc = 0;
    double nu             = 0.5;
    for (uint32_t i = 0; i < t.sizeQ; ++i) {
        uint64_t y = in[(size_t)qmap.pos[i] * N + rr];
        if (!premultiplied && MODE != 2)
            y = shoup(y, t.qhatinv[2 * i], t.qhatinv[2 * i + 1], r.mc[qmap.idx[i]].q);
        if (MODE != 0)
            nu = __dadd_rn(nu, __dmul_rn(__ull2double_rn(y), t.dbl[i]));
        const uint2 mm   = t.mat[(size_t)i * t.sizeP + jj];
        const uint64_t m = (uint64_t)mm.x | ((uint64_t)mm.y << 30);
        acc += (unsigned __int128)y * m;
        if ((i & 63u) == 63u)
            acc = barrett128((uint64_t)(acc >> 64), (uint64_t)acc, p, mc.mu128_hi, mc.mu128_lo);
    }
    uint64_t v = barrett128((uint64_t)(acc >> 64), (uint64_t)acc, p, mc.mu128_hi, mc.mu128_lo);
    if (MODE == 1)
        v = submod(v, t.alpha[(size_t)__double2ull_rz(nu) * t.sizeP + jj], p);
    else if (MODE == 2) {
        const uint64_t xo = in[(size_t)xomap.pos[jj] * N + rr];
        v                 = addmod(v, mulmod(xo, t.extra[jj], mc), p);
        uint64_t a;
        if (fabs(nu) <= 9223372036854774784.0)
            a = csub(lazy_reduce_n(__double2ull_rz(nu), mc.one_s, 0 - p), p);
        else {
            uint64_t hi, lo;
            double_to_u128(nu, hi, lo);
            a = barrett128(hi, lo, p, mc.mu128_hi, mc.mu128_lo);
        }
        v = addmod(v, a, p);
    }
    out[(size_t)pmap.pos[jj] * N + rr] = v;
}

}  // namespace

BConv::~BConv() {
    for (void* p : {(void*)qhatinv_d, (void*)mat_d, (void*)dbl_d, (void*)alpha_d, (void*)extra_d, (void*)tc_bmat_d, (void*)tc_dst_tab_d, (void*)tc_desc_d})
        if (p)
            cudaFree(p);
}

void launch_bconv(const Ring& r, const BConv& t, const uint64_t* in, uint64_t* out, uint32_t batch, size_t in_stride,
                  size_t out_stride, bool premultiplied, cudaStream_t st) {
    if (t.sizeP == 0 || t.sizeQ == 0 || batch == 0)
        return;
    if (bconv_tc_eligible(r, t)) {  // dense integer contraction on tcgen05 (baseconv_tc.cu)
        const BConvJob job{&t, in, out, batch, in_stride, out_stride};
        launch_bconv_tc(r, &job, 1, premultiplied, st);
        return;
    }
    BConvDev d{t.sizeQ, t.sizeP, t.qhatinv_d, reinterpret_cast<const uint2*>(t.mat_d), t.mode, t.dbl_d, t.alpha_d, t.extra_d};
    if (t.sizeQ > 64) {
        dim3 g((r.N + kThreads - 1) / kThreads, t.sizeP, batch);
        const int pmw = premultiplied ? 1 : 0;
        if (t.mode == 0)
            bconv_wide_kernel<0><<<g, kThreads, 0, st>>>(r.dev(), d, t.qmap, t.pmap, t.xomap, in, out, in_stride, out_stride, pmw);
        else if (t.mode == 1)
            bconv_wide_kernel<1><<<g, kThreads, 0, st>>>(r.dev(), d, t.qmap, t.pmap, t.xomap, in, out, in_stride, out_stride, pmw);
        else
            bconv_wide_kernel<2><<<g, kThreads, 0, st>>>(r.dev(), d, t.qmap, t.pmap, t.xomap, in, out, in_stride, out_stride, pmw);
        OFHE_CUDA(cudaGetLastError());
        return;
    }
    static const int cpt1 = [] {
        const char* e = getenv("OFHE_BCONV_CPT");  // tuning knob: coefficients per thread for <= 16 source limbs
        return e ? atoi(e) : 1;
    }();
    const int cpt = (t.sizeQ <= 16 && t.mode == 0) ? (cpt1 == 1 ? 1 : 2) : 1;
    dim3 grid((r.N + kThreads * cpt - 1) / (kThreads * cpt), (t.sizeP + kJPT - 1) / kJPT, batch);
    const int pm = premultiplied ? 1 : 0;
#define OFHE_BCONV_LAUNCH(CH, CPT, MODE) \
    bconv_kernel<CH, CPT, MODE><<<grid, kThreads, 0, st>>>(r.dev(), d, t.qmap, t.pmap, t.xomap, in, out, in_stride, out_stride, pm)
#define OFHE_BCONV_BY_SIZE(MODE)            \
    do {                                    \
        if (t.sizeQ <= 16)                  \
            OFHE_BCONV_LAUNCH(1, 1, MODE);  \
        else if (t.sizeQ <= 32)             \
            OFHE_BCONV_LAUNCH(2, 1, MODE);  \
        else if (t.sizeQ <= 48)             \
            OFHE_BCONV_LAUNCH(3, 1, MODE);  \
        else                                \
            OFHE_BCONV_LAUNCH(4, 1, MODE);  \
    } while (0)
    if (t.mode == 0 && t.sizeQ <= 16 && cpt == 2)
        OFHE_BCONV_LAUNCH(1, 2, 0);
    else if (t.mode == 0)
        OFHE_BCONV_BY_SIZE(0);
    else if (t.mode == 1)
        OFHE_BCONV_BY_SIZE(1);
    else
        OFHE_BCONV_BY_SIZE(2);
    OFHE_CUDA(cudaGetLastError());
}

}  // namespace ofhe
