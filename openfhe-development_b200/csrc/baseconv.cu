// baseconv.cu -- RNS fast base conversion (ApproxSwitchCRTBasis) as one fused per-coefficient kernel.
//
// Replaces DCRTPolyImpl::ApproxSwitchCRTBasis (lattice/hal/default/dcrtpoly-impl.h:888-932, int128 branch
// :895-918): for every coefficient r
//     y_i   = x_i[r] * QHatInvModq_i  mod q_i                       (Shoup)
//     out_j = ( sum_i y_i * QHatModp[i][j] )  mod p_j               (128-bit accumulate + Barrett)
// Results are the same canonical residues as the reference's Mul128 / BarrettUint128ModUint64 path.
//
// GPU formulation: one thread per coefficient, all source residues premultiplied into registers, the
// conversion matrix broadcast from shared memory.  The 60x60-bit multiply-accumulate is done on 30-bit
// halves so that each MAC is exactly four IMAD.WIDE.U32 with NO carry chain:
//     y = y1*2^30 + y0,  m = m1*2^30 + m0   ->   c0 += y0*m0; c1a += y0*m1; c1b += y1*m0; c2 += y1*m1
// Each partial sum stays below 2^64 for up to 16 terms; every 16 source limbs the four sums are folded into
// a 128-bit accumulator.  Requires q_i, p_j < 2^60 (MAX_MODULUS_SIZE, core/include/math/hal/basicint.h:48).
#include "internal.h"

namespace ofhe {
namespace {

constexpr uint32_t kMask30 = (1u << 30) - 1;

struct BConvDev {
    uint32_t sizeQ, sizeP;
    const uint64_t* qhatinv;  // [sizeQ][2]
    const uint2* mat;         // [sizeQ][sizeP]  (lo30, hi30)
};

template <int MAXQ>
__global__ void __launch_bounds__(128) bconv_kernel(RingDev r, BConvDev t, LimbMap qmap, LimbMap pmap,
                                                     const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                     size_t in_stride, size_t out_stride) {
    in += blockIdx.y * in_stride;
    out += blockIdx.y * out_stride;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: mat[sizeQ*sizeP] (uint2) | src consts [sizeQ][3] (q, w, ws) | dst consts [sizeP][3] (p, mu_hi, mu_lo)
    uint2* s_mat    = reinterpret_cast<uint2*>(smem_raw);
    uint64_t* s_src = reinterpret_cast<uint64_t*>(s_mat + (size_t)t.sizeQ * t.sizeP);
    uint64_t* s_dst = s_src + 3 * t.sizeQ;
    for (uint32_t i = threadIdx.x; i < t.sizeQ * t.sizeP; i += blockDim.x)
        s_mat[i] = t.mat[i];
    for (uint32_t i = threadIdx.x; i < t.sizeQ; i += blockDim.x) {
        s_src[3 * i]     = r.mc[qmap.idx[i]].q;
        s_src[3 * i + 1] = t.qhatinv[2 * i];
        s_src[3 * i + 2] = t.qhatinv[2 * i + 1];
    }
    for (uint32_t j = threadIdx.x; j < t.sizeP; j += blockDim.x) {
        const ModConst& mc = r.mc[pmap.idx[j]];
        s_dst[3 * j]       = mc.q;
        s_dst[3 * j + 1]   = mc.mu128_hi;
        s_dst[3 * j + 2]   = mc.mu128_lo;
    }
    __syncthreads();

    const uint32_t N  = r.N;
    const uint32_t rr = blockIdx.x * blockDim.x + threadIdx.x;
    if (rr >= N)
        return;
    uint32_t y0[MAXQ], y1[MAXQ];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) {
        if (i < (int)t.sizeQ) {
            uint64_t x = in[(size_t)qmap.pos[i] * N + rr];
            uint64_t y = shoup(x, s_src[3 * i + 1], s_src[3 * i + 2], s_src[3 * i]);
            y0[i]      = (uint32_t)y & kMask30;
            y1[i]      = (uint32_t)(y >> 30);
        }
        else {
            y0[i] = 0;
            y1[i] = 0;
        }
    }
    for (uint32_t j = 0; j < t.sizeP; ++j) {
        uint64_t hi = 0, lo = 0;
#pragma unroll
        for (int c = 0; c < MAXQ; c += 16) {
            if (c < (int)t.sizeQ) {
                uint64_t c0 = 0, c1a = 0, c1b = 0, c2 = 0;
#pragma unroll
                for (int i = c; i < c + 16 && i < MAXQ; ++i) {
                    if (i < (int)t.sizeQ) {
                        uint2 m = s_mat[(size_t)i * t.sizeP + j];
                        c0 += (uint64_t)y0[i] * m.x;
                        c1a += (uint64_t)y0[i] * m.y;
                        c1b += (uint64_t)y1[i] * m.x;
                        c2 += (uint64_t)y1[i] * m.y;
                    }
                }
                // fold: S += c0 + (c1a + c1b) * 2^30 + c2 * 2^60
                unsigned __int128 s = ((unsigned __int128)hi << 64) | lo;
                s += c0;
                s += (unsigned __int128)c1a << 30;
                s += (unsigned __int128)c1b << 30;
                s += (unsigned __int128)c2 << 60;
                hi = (uint64_t)(s >> 64);
                lo = (uint64_t)s;
            }
        }
        out[(size_t)pmap.pos[j] * N + rr] = barrett128(hi, lo, s_dst[3 * j], s_dst[3 * j + 1], s_dst[3 * j + 2]);
    }
}

}  // namespace

BConv::~BConv() {
    if (qhatinv_d)
        cudaFree(qhatinv_d);
    if (mat_d)
        cudaFree(mat_d);
}

void launch_bconv(const Ring& r, const BConv& t, const uint64_t* in, uint64_t* out, uint32_t batch, size_t in_stride,
                  size_t out_stride, cudaStream_t st) {
    if (t.sizeP == 0 || t.sizeQ == 0)
        return;
    BConvDev d{t.sizeQ, t.sizeP, t.qhatinv_d, reinterpret_cast<const uint2*>(t.mat_d)};
    size_t smem = (size_t)t.sizeQ * t.sizeP * sizeof(uint2) + 3 * sizeof(uint64_t) * (t.sizeQ + t.sizeP);
    const int threads = 128;
    dim3 grid((r.N + threads - 1) / threads, batch);
    OFHE_REQUIRE(t.sizeQ <= 64, "base conversion supports at most 64 source limbs");
    if (t.sizeQ <= 16) {
        if (smem > 48 * 1024)
            OFHE_CUDA(cudaFuncSetAttribute(bconv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bconv_kernel<16><<<grid, threads, smem, st>>>(r.dev(), d, t.qmap, t.pmap, in, out, in_stride, out_stride);
    }
    else if (t.sizeQ <= 32) {
        if (smem > 48 * 1024)
            OFHE_CUDA(cudaFuncSetAttribute(bconv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bconv_kernel<32><<<grid, threads, smem, st>>>(r.dev(), d, t.qmap, t.pmap, in, out, in_stride, out_stride);
    }
    else {
        if (smem > 48 * 1024)
            OFHE_CUDA(cudaFuncSetAttribute(bconv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bconv_kernel<64><<<grid, threads, smem, st>>>(r.dev(), d, t.qmap, t.pmap, in, out, in_stride, out_stride);
    }
    OFHE_CUDA(cudaGetLastError());
}

}  // namespace ofhe
