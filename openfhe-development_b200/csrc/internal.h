// internal.h -- host-side objects behind the opaque C handles of include/ofhe_b200.h, and the kernel launchers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "modarith.cuh"

namespace ofhe {

constexpr int kMaxLimbs = 256;  // limbs per launch (e.g. all dnum x (sizeQl+sizeP) converted limbs of a ModUp)

// limb -> (registered-modulus index, row of the slab), passed by value as a kernel parameter: digits, dropped
// limbs and the P-extension are metadata-only re-indexings (SURVEY.md 7 hard part 4), never copies.
struct LimbMap {
    uint32_t n;
    uint16_t idx[kMaxLimbs];  // modulus index in the ring
    uint16_t pos[kMaxLimbs];  // row index: limb data lives at base + pos * N
};

struct PosMap {  // output rows when a transform writes to a differently laid-out slab
    uint16_t pos[kMaxLimbs];
};

// Destinations of a transform's FINAL pass when the result is pushed to several buffers at once (this rank's gather
// buffer and, over NVLink peer mappings, the same rows of every other rank's): shard.cu / ofhe_shard_*_push.
constexpr uint32_t kMaxPeers = 8;
struct PeerOut {
    uint64_t* base[kMaxPeers];
    uint32_t n = 0;
};


// device view of a ring: tables of every registered modulus
struct RingDev {
    const ModConst* mc;     // [n_moduli]
    const Twiddle* tw_fwd;  // [n_moduli][N]  T[brev(i)] = psi^i      (transformnat-impl.h:714-756)
    const Twiddle* tw_inv;  // [n_moduli][N]  Tinv[brev(i)] = psi^-i
    uint32_t N;
    uint32_t logN;
};

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define OFHE_CUDA(expr)                                                                                     \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            throw ::ofhe::Error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " @" + __FILE__ + ":" + \
                                std::to_string(__LINE__));                                                  \
    } while (0)
#define OFHE_REQUIRE(cond, msg)            \
    do {                                   \
        if (!(cond))                       \
            throw ::ofhe::Error(msg);      \
    } while (0)

// Blocking host-to-device copy for one-time table / key uploads.  cudaMemcpy from PAGEABLE memory returns once the source
// is staged, possibly before the DMA into device memory has finished, and only the legacy default stream is ordered
// behind it; callers of this library launch on their own NON-BLOCKING streams (one per host thread in the lattice
// backend), which are not.  The explicit wait on the legacy stream closes that window (round 2: intermittently wrong
// first use of a freshly uploaded evaluation key under the relinked reference's test suites).
inline void h2d_blocking(void* dst, const void* src, size_t bytes) {
    OFHE_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    OFHE_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
}

struct Context {
    int device = 0;
    int sm_count = 0;
    // Scratch arenas for the fused pipelines, ONE PER STREAM: calls issued on different streams of the same
    // context (the reference's callers sit inside their own `omp parallel for`, keyswitch-hybrid.cpp:423,
    // ckksrns-fhe.cpp:1943) never share workspace, calls on one stream are ordered by the stream.  An arena only
    // grows; an outgrown block is retired (kept until the context dies) so that kernels already queued on it and
    // CUDA graphs that captured its address stay valid.  Growing is a cudaMalloc: it is refused while the stream
    // is being captured -- pre-size with ofhe_hks_reserve() (or one eager call) before capturing.
    struct Arena {
        void* p = nullptr;
        size_t bytes = 0;
    };
    std::map<cudaStream_t, Arena> arenas;
    std::vector<void*> retired;
    std::mutex mu;
    void* get_scratch(size_t bytes, cudaStream_t st);
    ~Context();
};

struct Ring;
Context* context_create(int device);
Ring* ring_create(Context* ctx, uint32_t N, uint32_t n_moduli, const uint64_t* moduli, const uint64_t* roots);
const uint32_t* ring_automap(Ring* r, uint32_t k, cudaStream_t st);

struct Ring {
    Context* ctx = nullptr;
    uint32_t N = 0, logN = 0, n_moduli = 0;
    std::vector<uint64_t> moduli, roots;
    std::vector<ModConst> mc_h;
    ModConst* mc_d = nullptr;
    Twiddle* tw_fwd_d = nullptr;
    Twiddle* tw_inv_d = nullptr;
    std::map<uint32_t, uint32_t*> automaps;  // k -> device map (nbtheory2.cpp:264-275)
    RingDev dev() const { return RingDev{mc_d, tw_fwd_d, tw_inv_d, N, logN}; }
    ~Ring();
};

inline LimbMap make_map(const uint32_t* idx, uint32_t n) {
    OFHE_REQUIRE(n <= (uint32_t)kMaxLimbs, "too many limbs in one call");
    LimbMap m;
    m.n = n;
    for (uint32_t i = 0; i < n; ++i) {
        m.idx[i] = (uint16_t)idx[i];
        m.pos[i] = (uint16_t)i;
    }
    return m;
}

// ---- kernel launchers (all asynchronous on `st`) ----------------------------------------------------
// NTT: in -> out (in == out allowed).  Limb l of the launch lives at base + lm.pos[l] * N in both slabs.
// If out_pos is given, limb l is read from in + lm.pos[l] * N and written to out + out_pos->pos[l] * N.
// Optional epilogue of a forward transform's LAST pass: instead of storing NTT(x) for launch limb l the kernel stores
//     dst[drow[l]] = (e[erow[l]] - NTT(x)) * scal[l]  (+ addend[drow[l]] if addend != nullptr; may alias dst)
// -- the tail of ApproxModDown (dcrtpoly-impl.h:1002), so that the converted polynomial never returns to HBM.
struct NttTail {
    const uint64_t* e;
    const uint64_t* addend;
    uint64_t* dst;
    const uint64_t* scal;  // device [lm.n][2]: the scalar and its Shoup constant
    PosMap erow, drow;
};
// whether launch_ntt_fwd can fuse a tail for this ring (else the caller runs launch_sub_mul_scalars itself)
bool ntt_fwd_tail_fused(const Ring& r);
void launch_ntt_fwd(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, cudaStream_t st,
                    const PosMap* out_pos = nullptr, const NttTail* tail = nullptr);
// `fold` (device, [lm.n][4]): per-limb replacement for (nInv, Shoup, Tinv[1]*nInv, Shoup) with an extra scalar
// folded in, see make_inv_fold(); nullptr = plain inverse transform.
void launch_ntt_inv(const Ring& r, const uint64_t* in, uint64_t* out, const LimbMap& lm, cudaStream_t st,
                    const PosMap* out_pos = nullptr, const uint64_t* fold = nullptr, const PeerOut* peers = nullptr);
void ntt_cluster_occupancy(uint32_t logN, int out[2]);  // diagnostic, see ntt.cu
// host helper: the four folded constants of limb `mod_idx` for extra scalar s
void make_inv_fold(const Ring& r, uint32_t mod_idx, uint64_t s, uint64_t out4[4]);

enum EwOp { EW_ADD = 0, EW_SUB = 1, EW_MUL = 2, EW_NEG = 3, EW_MAC = 4 /* dst += a*b */ };
void launch_ew(const Ring& r, EwOp op, uint64_t* dst, const uint64_t* a, const uint64_t* b, const LimbMap& lm,
               cudaStream_t st);
// dst = a * scalar[limb]  (+ addend if not null):  DCRTPolyImpl::Times(vector<NativeInteger>)
struct ScalarVec {
    uint64_t s[kMaxLimbs];
    uint64_t ss[kMaxLimbs];  // Shoup constants
};
void launch_mul_scalars(const Ring& r, uint64_t* dst, const uint64_t* a, const ScalarVec& sv, const LimbMap& lm,
                        cudaStream_t st);
void launch_automorphism(const Ring& r, uint64_t* dst, const uint64_t* src, const uint32_t* map_d, uint32_t n_limbs,
                         cudaStream_t st);
// COEFFICIENT-format automorphism (poly-impl.h:359-362), dst != src
void launch_automorphism_coeff(const Ring& r, uint64_t* dst, const uint64_t* src, uint32_t k, const LimbMap& lm,
                               cudaStream_t st);
void launch_build_automap(uint32_t N, uint32_t logN, uint32_t k, uint32_t* map_d, cudaStream_t st);
// dst[l] = centred switch of the single limb `src` (modulus index from_idx) into every modulus of lm, times
// scalar sv[l] (Shoup): the SwitchModulus + scalar multiply of DropLastElementAndScale (dcrtpoly-impl.h:703-705)
void launch_switch_modulus_scale(const Ring& r, uint64_t* dst, const uint64_t* src, uint32_t from_idx,
                                 const ScalarVec& sv, const LimbMap& lm, cudaStream_t st);
// dst = a * s[l] + b   (the tail of DropLastElementAndScale, dcrtpoly-impl.h:708-709)
void launch_mul_scalars_add(const Ring& r, uint64_t* dst, const uint64_t* a, const ScalarVec& sv, const uint64_t* b,
                            const LimbMap& lm, cudaStream_t st);
// dst = (a - b) * s[l] (+ c if c != nullptr)   (the tail of ApproxModDown, dcrtpoly-impl.h:1002, optionally fused
// with the `cv[k] += ab[k]` of relinearisation, base-leveledshe.cpp:209-210)
void launch_sub_mul_scalars(const Ring& r, uint64_t* dst, const uint64_t* a, const uint64_t* b, const ScalarVec& sv,
                            const uint64_t* c, const LimbMap& lm, cudaStream_t st);

// BEHZ correction steps (behz.cu): dcrtpoly-impl.h:1724-1738, 1760-1767, 1903-1907
void launch_behz_mtilde(const Ring& r, const uint64_t* in, const LimbMap& qmap, const ScalarVec& hat, const ScalarVec& qhm,
                        uint64_t neg_qinv, uint64_t* out, cudaStream_t st);
void launch_behz_mtilde_fix(const Ring& r, uint64_t* s, const LimbMap& bmap, const uint64_t* rm, const ScalarVec& qmod,
                            const ScalarVec& minv, cudaStream_t st);
void launch_behz_sk_fix(const Ring& r, uint64_t* qc, const LimbMap& qmap, const uint64_t* alpha, uint64_t msk,
                        const ScalarVec& bmod, cudaStream_t st);

// fast base conversion tables (a10), device resident
struct BConv {
    uint32_t sizeQ = 0, sizeP = 0;
    LimbMap qmap, pmap;  // source limbs (rows of `in`) and target limbs (rows of `out`: pos may leave a gap)
    // device: per source limb QHatInvModq + Shoup constant; matrix QHatModp split in 30-bit halves [sizeQ][sizeP]
    uint64_t* qhatinv_d = nullptr;   // [sizeQ][2] (w, ws)
    uint32_t* mat_d = nullptr;       // [sizeQ][sizeP][2] (lo30, hi30)
    // mode 0: ApproxSwitchCRTBasis (a10).  mode 1: exact SwitchCRTBasis (a12): nu = 0.5 + sum double(y_i)*dbl[i],
    // subtract alpha_d[(size_t)nu][j].  mode 2: ScaleAndRound (a13): sources are used as they are, nu = 0.5 +
    // sum dbl[i]*double(x_i) is ADDED (mod o_j) together with x_{o_j} * extra_d[j]; xomap.pos[j] = row of x_{o_j} in `in`.
    int mode = 0;
    double* dbl_d = nullptr;         // [sizeQ]
    uint64_t* alpha_d = nullptr;     // [sizeQ+1][sizeP]
    uint64_t* extra_d = nullptr;     // [sizeP]
    PosMap xomap;
    // tensor-core path (baseconv_tc.cu): byte-digit image of the matrix in the MMA's shared-memory layout and the
    // device-resident description of the conversion; built by bconv_finalize() for <= 16 source limbs
    uint8_t* tc_bmat_d = nullptr;
    uint64_t* tc_dst_tab_d = nullptr;
    struct BConvTcDesc* tc_desc_d = nullptr;
    ~BConv();
};
struct BConvTcDesc {
    // per source limb (padded to 16): q, QHatInvModq, its Shoup constant, word offset of the input row (~0: padding)
    uint64_t src_tab[16][4];
    uint32_t sizeQ, sizeP;
    const uint8_t* bmat;      // [ceil(sizeP / 16)][256 rows][128 bytes], 128-byte swizzle
    // per target limb: p, -p^-1 mod 2^64, word offset of the output row, (pad)
    const uint64_t* dst_tab;  // [sizeP][4]
};
// called once by whoever fills a BConv (after mat_d / qhatinv_d / the maps are in place)
void bconv_finalize(const Ring& r, BConv& t);
bool bconv_tc_eligible(const Ring& r, const BConv& t);
// several approximate conversions (mode 0, <= 16 source limbs each) as ONE persistent tensor-core launch
struct BConvJob {
    const BConv* t;
    const uint64_t* in;
    uint64_t* out;
    uint32_t batch;
    size_t in_stride, out_stride;
};
void launch_bconv_tc(const Ring& r, const BConvJob* jobs, uint32_t njobs, bool premultiplied, cudaStream_t st);
// `batch` independent polynomials, `in_stride` / `out_stride` words apart
// premultiplied: the source residues already carry the factor QHatInvModq (folded into the preceding inverse NTT)
void launch_bconv(const Ring& r, const BConv& t, const uint64_t* in, uint64_t* out, uint32_t batch, size_t in_stride,
                  size_t out_stride, bool premultiplied, cudaStream_t st);

// hybrid key-switch inner product (a17)
void launch_ks_inner_product(const Ring& r, uint32_t n_digits, const uint64_t* digits, size_t digit_stride,
                             const uint64_t* evk_a, const uint64_t* evk_b, size_t evk_stride, uint32_t sizeQl,
                             uint32_t delta, uint64_t* e0, uint64_t* e1, const LimbMap& lm, cudaStream_t st,
                             const uint64_t* c_own = nullptr, uint32_t alpha = 0);
// c_own != nullptr: digit j's own rows [j alpha, (j+1) alpha) are read from the switched polynomial c_own [sizeQl][N]
// instead of from the digit slabs (which then need not hold them)
// EvalMultCore tensor (a19): c0=a0*b0, c1=a0*b1+a1*b0, c2=a1*b1
void launch_tensor(const Ring& r, const uint64_t* a0, const uint64_t* a1, const uint64_t* b0, const uint64_t* b1,
                   uint64_t* c0, uint64_t* c1, uint64_t* c2, const LimbMap& lm, cudaStream_t st);

}  // namespace ofhe
