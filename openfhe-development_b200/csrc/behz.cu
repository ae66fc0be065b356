// behz.cu -- the per-coefficient correction steps of the BEHZ variant of BFV multiplication.
//
// DCRTPolyImpl::FastBaseConvqToBskMontgomery (lattice/hal/default/dcrtpoly-impl.h:1694-1777), ::FastRNSFloorq
// (:1780-1830) and ::FastBaseConvSK (:1833-1921) are fast base conversions -- the multiply-accumulate parts run through
// bconv_kernel (baseconv.cu) with the reference's tables -- followed by small corrections that are NOT plain limb-wise
// modular operations: the Montgomery digit modulo m~ = 2^16, and the centred lift of the m_sk residue.  Those are the
// three kernels here; same canonical residues as the reference, including its unsigned wrap-around in ModSubFast.
#include "internal.h"

namespace ofhe {
namespace {

// r[k] = (( sum_i y_i[k] * QHatModmtilde[i] ) mod 2^16) * negQInvModmtilde  mod 2^16,
// y_i = x_i * mtildeQHatInvModq_i mod q_i                                                   (:1724-1738)
__global__ void __launch_bounds__(256) behz_mtilde_kernel(RingDev r, const uint64_t* __restrict__ in, LimbMap qmap,
                                                           ScalarVec hat, ScalarVec qhm, uint64_t neg_qinv,
                                                           uint64_t* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= r.N)
        return;
    uint64_t acc = 0;  // only the low 16 bits matter: 64-bit wrap-around is harmless, as in the reference
    for (uint32_t i = 0; i < qmap.n; ++i) {
        const uint64_t y = shoup(in[(size_t)qmap.pos[i] * r.N + k], hat.s[i], hat.ss[i], r.mc[qmap.idx[i]].q);
        acc += y * qhm.s[i];
    }
    acc &= 0xffffu;
    acc *= neg_qinv;
    out[k] = acc & 0xffffu;
}

// s_j[k] <- ( (centre(r[k]) * QModbsk_j + s_j[k]) * mtildeInvModbsk_j ) mod bsk_j                (:1760-1767)
// centre(r) = r >= 2^15 ? r + bsk_j - 2^16 : r   (m~ = 2^16 is below every modulus of Bsk)
__global__ void __launch_bounds__(256) behz_mtilde_fix_kernel(RingDev r, uint64_t* __restrict__ s, LimbMap bmap,
                                                               const uint64_t* __restrict__ rm, ScalarVec qmod,
                                                               ScalarVec minv) {
    const uint32_t j = blockIdx.y;
    const uint64_t p = r.mc[bmap.idx[j]].q;
    uint64_t* row    = s + (size_t)bmap.pos[j] * r.N;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < r.N; k += gridDim.x * blockDim.x) {
        uint64_t x = rm[k];
        if (x >= 0x8000u)
            x += p - 0x10000u;
        x      = shoup(x, qmod.s[j], qmod.ss[j], p);
        x      = addmod(x, row[k], p);
        row[k] = shoup(x, minv.s[j], minv.ss[j], p);
    }
}

// q_j[k] <- q_j[k] - (lift(alpha[k]) * BModq_j mod q_j)   mod q_j                                  (:1903-1907)
// lift(a) = a > msk/2 ? a.ModSubFast(msk, q_j) : a, with ModSubFast(b, m) = a < b ? a + m - b : a - b computed in
// 64-bit unsigned arithmetic exactly as ubintnat.h:911-915 (q_j - msk may wrap; the product takes any 64-bit operand)
__global__ void __launch_bounds__(256) behz_sk_fix_kernel(RingDev r, uint64_t* __restrict__ qc, LimbMap qmap,
                                                           const uint64_t* __restrict__ alpha, uint64_t msk,
                                                           ScalarVec bmod) {
    const uint32_t j  = blockIdx.y;
    const uint64_t q  = r.mc[qmap.idx[j]].q;
    const uint64_t hm = msk >> 1;
    uint64_t* row     = qc + (size_t)qmap.pos[j] * r.N;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < r.N; k += gridDim.x * blockDim.x) {
        uint64_t a = alpha[k];
        if (a > hm)
            a = a < msk ? a + q - msk : a - msk;
        a            = shoup(a, bmod.s[j], bmod.ss[j], q);
        const uint64_t x = row[k];
        row[k]       = x < a ? x + q - a : x - a;
    }
}

}  // namespace

void launch_behz_mtilde(const Ring& r, const uint64_t* in, const LimbMap& qmap, const ScalarVec& hat, const ScalarVec& qhm,
                        uint64_t neg_qinv, uint64_t* out, cudaStream_t st) {
    behz_mtilde_kernel<<<(r.N + 255) / 256, 256, 0, st>>>(r.dev(), in, qmap, hat, qhm, neg_qinv, out);
    OFHE_CUDA(cudaGetLastError());
}
void launch_behz_mtilde_fix(const Ring& r, uint64_t* s, const LimbMap& bmap, const uint64_t* rm, const ScalarVec& qmod,
                            const ScalarVec& minv, cudaStream_t st) {
    if (bmap.n == 0)
        return;
    behz_mtilde_fix_kernel<<<dim3(std::max(1u, std::min(64u, r.N / 256)), bmap.n), 256, 0, st>>>(r.dev(), s, bmap, rm, qmod,
                                                                                                 minv);
    OFHE_CUDA(cudaGetLastError());
}
void launch_behz_sk_fix(const Ring& r, uint64_t* qc, const LimbMap& qmap, const uint64_t* alpha, uint64_t msk,
                        const ScalarVec& bmod, cudaStream_t st) {
    if (qmap.n == 0)
        return;
    behz_sk_fix_kernel<<<dim3(std::max(1u, std::min(64u, r.N / 256)), qmap.n), 256, 0, st>>>(r.dev(), qc, qmap, alpha, msk,
                                                                                             bmod);
    OFHE_CUDA(cudaGetLastError());
}

}  // namespace ofhe
