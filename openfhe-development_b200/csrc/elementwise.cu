// elementwise.cu -- limb-wise (Hadamard) arithmetic on [limbs x N] slabs: HBM-streaming kernels with
// 128-bit vectorised, fully coalesced loads/stores; the grid's y dimension is the RNS limb.
//
// Replaces, with identical canonical results:
//   DCRTPolyImpl::Plus/Minus/Times/Negate/operator+=,-=,*=   lattice/hal/default/dcrtpoly.h:130-190,
//                                                            dcrtpoly-impl.h:347-621
//   NativeVectorT::ModAddNoCheckEq / ModMulNoCheckEq / ModSubEq / ModMulEq(scalar)
//                                                            math/hal/intnat/mubintvecnat.h:417-423,492-503,
//                                                            mubintvecnat.cpp:273,295-304
//   NativeVectorT::SwitchModulus                             mubintvecnat.cpp:109-122
//   PolyImpl::AutomorphismTransform(k, precomp)              lattice/hal/default/poly-impl.h:366-376
//   PrecomputeAutoMap                                        core/lib/math/nbtheory2.cpp:264-275
//   KeySwitchHYBRID::EvalFastKeySwitchCoreExt                pke/lib/keyswitch/keyswitch-hybrid.cpp:402-435
//   LeveledSHEBase::EvalMultCore (2x2 tensor)                pke/lib/schemebase/base-leveledshe.cpp:608-646
#include "internal.h"

namespace ofhe {
namespace {

using v2 = ulonglong2;
__device__ __forceinline__ v2 ldv(const uint64_t* p, size_t i) {
    return reinterpret_cast<const v2*>(p)[i];
}
__device__ __forceinline__ void stv(uint64_t* p, size_t i, v2 v) {
    reinterpret_cast<v2*>(p)[i] = v;
}

// Each thread handles pairs of words; N is a power of two >= 2 so N/2 pairs per limb.
// dst may alias a or b (the in-place operators), so these two kernels carry no __restrict__.
template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(RingDev r, uint64_t* dst, const uint64_t* a,
                                                  const uint64_t* b, LimbMap lm) {
    const uint32_t limb = blockIdx.y;
    const ModConst mc   = r.mc[lm.idx[limb]];
    const uint64_t q    = mc.q;
    const size_t half   = r.N >> 1;
    const size_t off    = (size_t)lm.pos[limb] * half;  // row of the slab (LimbMap: idx = modulus, pos = row)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        v2 x = ldv(a, off + i), o;
        if (OP == EW_NEG) {
            o.x = x.x ? q - x.x : 0;
            o.y = x.y ? q - x.y : 0;
        }
        else {
            v2 y = ldv(b, off + i);
            if (OP == EW_ADD) {
                o.x = addmod(x.x, y.x, q);
                o.y = addmod(x.y, y.y, q);
            }
            else if (OP == EW_SUB) {
                o.x = submod(x.x, y.x, q);
                o.y = submod(x.y, y.y, q);
            }
            else if (OP == EW_MUL) {
                o.x = mulmod(x.x, y.x, mc);
                o.y = mulmod(x.y, y.y, mc);
            }
            else {  // EW_MAC: dst += a*b
                v2 d = ldv(dst, off + i);
                o.x  = addmod(d.x, mulmod(x.x, y.x, mc), q);
                o.y  = addmod(d.y, mulmod(x.y, y.y, mc), q);
            }
        }
        stv(dst, off + i, o);
    }
}

// MODE 0: dst = a*s   1: dst = a*s + b   2: dst = (a-b)*s   3: dst = (a-b)*s + c
template <int MODE>
__global__ void __launch_bounds__(256) scal_kernel(RingDev r, uint64_t* dst, const uint64_t* a, const uint64_t* b,
                                                    const uint64_t* c, ScalarVec sv, LimbMap lm) {
    const uint32_t limb = blockIdx.y;
    const uint64_t q    = r.mc[lm.idx[limb]].q;
    const uint64_t s = sv.s[limb], ss = sv.ss[limb];
    const size_t half = r.N >> 1;
    const size_t off  = (size_t)lm.pos[limb] * half;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        v2 x = ldv(a, off + i), o;
        if (MODE == 0) {
            o.x = shoup(x.x, s, ss, q);
            o.y = shoup(x.y, s, ss, q);
        }
        else if (MODE == 1) {
            v2 y = ldv(b, off + i);
            o.x  = addmod(shoup(x.x, s, ss, q), y.x, q);
            o.y  = addmod(shoup(x.y, s, ss, q), y.y, q);
        }
        else {
            v2 y = ldv(b, off + i);
            o.x  = shoup(submod(x.x, y.x, q), s, ss, q);
            o.y  = shoup(submod(x.y, y.y, q), s, ss, q);
            if (MODE == 3) {
                v2 z = ldv(c, off + i);
                o.x  = addmod(o.x, z.x, q);
                o.y  = addmod(o.y, z.y, q);
            }
        }
        stv(dst, off + i, o);
    }
}

// centred lift of one limb (modulus from_idx) into every modulus of lm, then * s[limb]
__global__ void __launch_bounds__(256) switch_mod_kernel(RingDev r, uint64_t* __restrict__ dst,
                                                          const uint64_t* __restrict__ src, uint32_t from_idx,
                                                          ScalarVec sv, LimbMap lm) {
    const uint32_t limb = blockIdx.y;
    const uint64_t qt   = r.mc[lm.idx[limb]].q;
    const uint64_t qf   = r.mc[from_idx].q;
    const uint64_t halfq = qf >> 1;
    const uint64_t diff  = qf > qt ? qf - qt : qt - qf;
    const uint64_t s = sv.s[limb], ss = sv.ss[limb];
    const uint64_t diffr = diff % qt;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < r.N; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = src[i];
        if (qt > qf) {
            if (v > halfq)
                v += diff;  // plain add (mubintvecnat.cpp:114-116); v stays < qt
        }
        else {
            uint64_t av = v % qt;
            uint64_t bv = v > halfq ? diffr : 0;
            v           = av >= bv ? av - bv : av + qt - bv;
        }
        dst[(size_t)lm.pos[limb] * r.N + i] = shoup(v, s, ss, qt);
    }
}

__global__ void automap_kernel(uint32_t N, uint32_t logN, uint32_t k, uint32_t* __restrict__ map) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N)
        return;
    uint32_t idx = ((((2 * j + 1) * (uint64_t)k) & (2 * (uint64_t)N - 1)) >> 1);
    map[__brev(j) >> (32 - logN)] = __brev(idx) >> (32 - logN);
}
__global__ void __launch_bounds__(256) automorph_kernel(uint32_t N, uint64_t* __restrict__ dst,
                                                         const uint64_t* __restrict__ src,
                                                         const uint32_t* __restrict__ map) {
    const size_t off = (size_t)blockIdx.y * N;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x)
        dst[off + j] = src[off + map[j]];
}

// COEFFICIENT-format automorphism X -> X^k (poly-impl.h:359-362): out[jk mod N] = (jk div N odd) ? q - in[j] : in[j].
// The reference subtracts in plain integers, so a zero coefficient with an odd wrap count comes out as q itself, not 0;
// reproduced here (bit-exactness is the contract).  j -> jk mod N is a bijection for odd k: every word written once.
__global__ void __launch_bounds__(256) automorph_coeff_kernel(RingDev r, uint64_t* __restrict__ dst,
                                                               const uint64_t* __restrict__ src, uint32_t k, LimbMap lm) {
    const uint32_t limb = blockIdx.y;
    const uint64_t q    = r.mc[lm.idx[limb]].q;
    const size_t off    = (size_t)lm.pos[limb] * r.N;
    const uint32_t mask = r.N - 1;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < r.N; j += gridDim.x * blockDim.x) {
        const uint32_t jk = j * k;
        const uint64_t v  = src[off + j];
        dst[off + (jk & mask)] = ((jk >> r.logN) & 1u) ? q - v : v;
    }
}

// e0[i] = sum_j D_j[i] * b_j[idx(i)],  e1[i] = sum_j D_j[i] * a_j[idx(i)]: products accumulated in 128 bits
// (dnum * q^2 < 2^128), one reduction per output word.
__global__ void __launch_bounds__(256) ks_inner_kernel(RingDev r, uint32_t nd, const uint64_t* __restrict__ digits,
                                                        size_t dstride, const uint64_t* __restrict__ ea,
                                                        const uint64_t* __restrict__ eb, size_t estride,
                                                        uint32_t sizeQl, uint32_t delta, uint64_t* __restrict__ e0,
                                                        uint64_t* __restrict__ e1, LimbMap lm,
                                                        const uint64_t* __restrict__ c_own, uint32_t alpha) {
    const uint32_t limb = blockIdx.y;
    const ModConst mc   = r.mc[lm.idx[limb]];
    const uint32_t kidx = limb >= sizeQl ? limb + delta : limb;
    const size_t half   = r.N >> 1;
    const size_t doff = (size_t)lm.pos[limb] * half, eoff = (size_t)kidx * half;
    // c_own: digit j's OWN rows (limbs [j alpha, (j+1) alpha) of Ql) are the original limbs of the switched polynomial
    // (keyswitch-hybrid.cpp:372-373); read them from there instead of from copies inside the digit slabs
    const uint32_t own = (c_own && limb < sizeQl) ? limb / alpha : 0xffffffffu;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        unsigned __int128 s0x = 0, s0y = 0, s1x = 0, s1y = 0;
        for (uint32_t j = 0; j < nd; ++j) {
            v2 d = ldv(j == own ? c_own : digits + j * dstride, doff + i);
            v2 b = ldv(eb + j * estride, eoff + i);
            v2 a = ldv(ea + j * estride, eoff + i);
            s0x += (unsigned __int128)d.x * b.x;
            s0y += (unsigned __int128)d.y * b.y;
            s1x += (unsigned __int128)d.x * a.x;
            s1y += (unsigned __int128)d.y * a.y;
        }
        v2 o0, o1;
        o0.x = barrett128((uint64_t)(s0x >> 64), (uint64_t)s0x, mc.q, mc.mu128_hi, mc.mu128_lo);
        o0.y = barrett128((uint64_t)(s0y >> 64), (uint64_t)s0y, mc.q, mc.mu128_hi, mc.mu128_lo);
        o1.x = barrett128((uint64_t)(s1x >> 64), (uint64_t)s1x, mc.q, mc.mu128_hi, mc.mu128_lo);
        o1.y = barrett128((uint64_t)(s1y >> 64), (uint64_t)s1y, mc.q, mc.mu128_hi, mc.mu128_lo);
        stv(e0, doff + i, o0);
        stv(e1, doff + i, o1);
    }
}

__global__ void __launch_bounds__(256) tensor_kernel(RingDev r, const uint64_t* __restrict__ a0,
                                                      const uint64_t* __restrict__ a1, const uint64_t* __restrict__ b0,
                                                      const uint64_t* __restrict__ b1, uint64_t* __restrict__ c0,
                                                      uint64_t* __restrict__ c1, uint64_t* __restrict__ c2,
                                                      LimbMap lm) {
    const uint32_t limb = blockIdx.y;
    const ModConst mc   = r.mc[lm.idx[limb]];
    const size_t half   = r.N >> 1;
    const size_t off    = (size_t)lm.pos[limb] * half;  // row of the slab (LimbMap: idx = modulus, pos = row)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        v2 x0 = ldv(a0, off + i), x1 = ldv(a1, off + i), y0 = ldv(b0, off + i), y1 = ldv(b1, off + i);
        v2 o0, o1, o2;
        o0.x = mulmod(x0.x, y0.x, mc);
        o0.y = mulmod(x0.y, y0.y, mc);
        o2.x = mulmod(x1.x, y1.x, mc);
        o2.y = mulmod(x1.y, y1.y, mc);
        // c1 = a0*b1 + a1*b0: sum the two 128-bit products, reduce once (same residue as two mulmods + add)
        unsigned __int128 sx = (unsigned __int128)x0.x * y1.x + (unsigned __int128)x1.x * y0.x;
        unsigned __int128 sy = (unsigned __int128)x0.y * y1.y + (unsigned __int128)x1.y * y0.y;
        o1.x = barrett128((uint64_t)(sx >> 64), (uint64_t)sx, mc.q, mc.mu128_hi, mc.mu128_lo);
        o1.y = barrett128((uint64_t)(sy >> 64), (uint64_t)sy, mc.q, mc.mu128_hi, mc.mu128_lo);
        stv(c0, off + i, o0);
        stv(c1, off + i, o1);
        stv(c2, off + i, o2);
    }
}

inline dim3 ew_grid(const Ring& r, uint32_t limbs, size_t items_per_limb) {
    uint32_t bx = (uint32_t)((items_per_limb + 255) / 256);
    // enough CTAs to cover the chip several times, grid-stride beyond that
    uint32_t cap = (uint32_t)(r.ctx->sm_count * 8 + limbs - 1) / limbs;
    if (cap < 1)
        cap = 1;
    if (bx > cap)
        bx = cap;
    return dim3(bx, limbs);
}

}  // namespace

void launch_ew(const Ring& r, EwOp op, uint64_t* dst, const uint64_t* a, const uint64_t* b, const LimbMap& lm,
               cudaStream_t st) {
    if (lm.n == 0)
        return;
    OFHE_REQUIRE(r.N >= 2, "ring dimension must be >= 2");
    dim3 g = ew_grid(r, lm.n, r.N / 2);
    switch (op) {
        case EW_ADD: ew_kernel<EW_ADD><<<g, 256, 0, st>>>(r.dev(), dst, a, b, lm); break;
        case EW_SUB: ew_kernel<EW_SUB><<<g, 256, 0, st>>>(r.dev(), dst, a, b, lm); break;
        case EW_MUL: ew_kernel<EW_MUL><<<g, 256, 0, st>>>(r.dev(), dst, a, b, lm); break;
        case EW_NEG: ew_kernel<EW_NEG><<<g, 256, 0, st>>>(r.dev(), dst, a, b, lm); break;
        case EW_MAC: ew_kernel<EW_MAC><<<g, 256, 0, st>>>(r.dev(), dst, a, b, lm); break;
    }
    OFHE_CUDA(cudaGetLastError());
}
void launch_mul_scalars(const Ring& r, uint64_t* dst, const uint64_t* a, const ScalarVec& sv, const LimbMap& lm,
                        cudaStream_t st) {
    if (lm.n == 0)
        return;
    scal_kernel<0><<<ew_grid(r, lm.n, r.N / 2), 256, 0, st>>>(r.dev(), dst, a, nullptr, nullptr, sv, lm);
    OFHE_CUDA(cudaGetLastError());
}
void launch_mul_scalars_add(const Ring& r, uint64_t* dst, const uint64_t* a, const ScalarVec& sv, const uint64_t* b,
                            const LimbMap& lm, cudaStream_t st) {
    if (lm.n == 0)
        return;
    scal_kernel<1><<<ew_grid(r, lm.n, r.N / 2), 256, 0, st>>>(r.dev(), dst, a, b, nullptr, sv, lm);
    OFHE_CUDA(cudaGetLastError());
}
void launch_sub_mul_scalars(const Ring& r, uint64_t* dst, const uint64_t* a, const uint64_t* b, const ScalarVec& sv,
                            const uint64_t* c, const LimbMap& lm, cudaStream_t st) {
    if (lm.n == 0)
        return;
    if (c)
        scal_kernel<3><<<ew_grid(r, lm.n, r.N / 2), 256, 0, st>>>(r.dev(), dst, a, b, c, sv, lm);
    else
        scal_kernel<2><<<ew_grid(r, lm.n, r.N / 2), 256, 0, st>>>(r.dev(), dst, a, b, nullptr, sv, lm);
    OFHE_CUDA(cudaGetLastError());
}
void launch_switch_modulus_scale(const Ring& r, uint64_t* dst, const uint64_t* src, uint32_t from_idx,
                                 const ScalarVec& sv, const LimbMap& lm, cudaStream_t st) {
    if (lm.n == 0)
        return;
    switch_mod_kernel<<<ew_grid(r, lm.n, r.N), 256, 0, st>>>(r.dev(), dst, src, from_idx, sv, lm);
    OFHE_CUDA(cudaGetLastError());
}
void launch_build_automap(uint32_t N, uint32_t logN, uint32_t k, uint32_t* map_d, cudaStream_t st) {
    automap_kernel<<<(N + 255) / 256, 256, 0, st>>>(N, logN, k, map_d);
    OFHE_CUDA(cudaGetLastError());
}
void launch_automorphism(const Ring& r, uint64_t* dst, const uint64_t* src, const uint32_t* map_d, uint32_t n_limbs,
                         cudaStream_t st) {
    if (n_limbs == 0)
        return;
    automorph_kernel<<<ew_grid(r, n_limbs, r.N), 256, 0, st>>>(r.N, dst, src, map_d);
    OFHE_CUDA(cudaGetLastError());
}
void launch_automorphism_coeff(const Ring& r, uint64_t* dst, const uint64_t* src, uint32_t k, const LimbMap& lm,
                               cudaStream_t st) {
    if (lm.n == 0)
        return;
    automorph_coeff_kernel<<<ew_grid(r, lm.n, r.N), 256, 0, st>>>(r.dev(), dst, src, k, lm);
    OFHE_CUDA(cudaGetLastError());
}
void launch_ks_inner_product(const Ring& r, uint32_t nd, const uint64_t* digits, size_t dstride,
                             const uint64_t* evk_a, const uint64_t* evk_b, size_t estride, uint32_t sizeQl,
                             uint32_t delta, uint64_t* e0, uint64_t* e1, const LimbMap& lm, cudaStream_t st,
                             const uint64_t* c_own, uint32_t alpha) {
    if (lm.n == 0)
        return;
    ks_inner_kernel<<<ew_grid(r, lm.n, r.N / 2), 256, 0, st>>>(r.dev(), nd, digits, dstride, evk_a, evk_b, estride,
                                                               sizeQl, delta, e0, e1, lm, c_own, alpha ? alpha : 1);
    OFHE_CUDA(cudaGetLastError());
}
void launch_tensor(const Ring& r, const uint64_t* a0, const uint64_t* a1, const uint64_t* b0, const uint64_t* b1,
                   uint64_t* c0, uint64_t* c1, uint64_t* c2, const LimbMap& lm, cudaStream_t st) {
    if (lm.n == 0)
        return;
    tensor_kernel<<<ew_grid(r, lm.n, r.N / 2), 256, 0, st>>>(r.dev(), a0, a1, b0, b1, c0, c1, c2, lm);
    OFHE_CUDA(cudaGetLastError());
}

}  // namespace ofhe
