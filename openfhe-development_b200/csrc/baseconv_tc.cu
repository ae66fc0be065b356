// baseconv_tc.cu -- RNS fast base conversion (ApproxSwitchCRTBasis, dcrtpoly-impl.h:888-932) on the 5th-generation
// tensor cores: the dense INTEGER contraction of the hybrid key switch.
//
//     out_j[r] = ( sum_i y_i[r] * M[i][j] )  mod p_j,      y_i = x_i * QHatInvModq_i mod q_i,  M[i][j] = QHatModp[i][j]
//
// is, for every target limb j, an exact 124-bit integer dot product over <= 16 source limbs followed by one
// 128 -> 64 Barrett step.  The CUDA-core kernel (baseconv.cu) spends 3 IMAD.WIDE + a 64-bit add per 60x60-bit MAC
// (215 SASS instructions per output residue, measured).  Here the dot products run on tcgen05.mma kind::i8:
//
//   * residues and matrix entries are split into their 8 bytes -- no work, the bytes of the 64-bit words ARE the
//     digits.  A tile of 128 coefficients becomes the K-major u8 operand A[128][K = 16 limbs x 8 bytes];
//   * the matrix becomes B[(i,a)][(j,c)] = byte (c-a) of M[i][j]  (0 <= c-a < 8, else 0), c = 0..15: sixteen u8
//     columns per target limb, one per digit weight 2^(8c) -- built once per conversion table, device resident in the
//     exact shared-memory image (128-byte swizzle) the MMA reads;
//   * D[r][(j,c)] = sum_k A[r][k] B[k][(j,c)] <= 120 * 255^2 < 2^23 in s32, exact, accumulated in TMEM;
//   * the epilogue reads the 15 column sums of an output with tcgen05.ld (one TMEM lane = one coefficient = one
//     thread), rebuilds S = sum_c D_c 2^(8c) < 2^124 with 7 shift-adds, 4 funnel shifts and one 128-bit add, and
//     reduces it with the same Barrett step as the CUDA-core kernel: the canonical residue, bit for bit.
//
// One persistent CTA per SM: 512 threads, all 512 TMEM columns as two 256-column accumulator buffers (16 target limbs
// each), the whole B image of the job (<= 5 chunks of 32 KiB) and two 16 KiB A tiles in shared memory.  Thread 0
// issues the MMAs (4 x M128 N256 K32 per chunk) and commits them to an mbarrier per buffer; all 16 warps run the
// epilogue of one buffer while the tensor core fills the other, and transpose the next tile ([limb][coefficient] in
// HBM -> [coefficient][limb bytes] in shared memory) through registers.  Several conversions (the digits of a ModUp,
// the two polynomials of a ModDown) are one launch: a job list, work = (job, polynomial, tile).
//
// Eligibility: mode 0 (approximate conversion), <= 16 source limbs, N >= 128.  Everything else stays on baseconv.cu.
#include <cstdlib>

#include "internal.h"

namespace ofhe {
namespace {

constexpr int kTcThreads  = 512;
constexpr int kTileM      = 128;                  // coefficients per tile = TMEM lanes
constexpr int kChunkOut   = 16;                   // target limbs per accumulator buffer
constexpr int kChunkN     = 256;                  // columns per buffer: 16 limbs x 16 digit weights
constexpr int kK          = 128;                  // 16 source limbs x 8 bytes
constexpr int kChunkBytes = kChunkN * kK;         // 32 KiB of B per chunk
constexpr int kTileBytes  = kTileM * kK;          // 16 KiB of A per tile
constexpr int kMaxChunks  = 5;                    // 80 target limbs per job (larger conversions are split into jobs)
constexpr int kMaxJobs    = 4;
constexpr int kSub        = kTcThreads / 128;     // warps per TMEM lane quarter
constexpr int kOutPerThr  = kChunkOut / kSub;     // outputs of a chunk per thread

struct TcJob {
    const BConvTcDesc* d;
    const uint64_t* in;
    uint64_t* out;
    size_t in_stride, out_stride;
    uint32_t batch;
    uint32_t j0, j1;    // target range [j0, j1) of this job, j0 a multiple of 16
    uint32_t work_end;  // exclusive prefix sum of batch * tiles over the jobs
};
struct TcLaunch {
    TcJob job[kMaxJobs];
    uint32_t njobs;
    uint32_t tiles;  // N / 128
    uint32_t premultiplied;
    uint32_t debug_col;  // diagnostics only (OFHE_BCONV_TC_DEBUG): 1 + c stores the raw column sum D_c instead of the residue
};

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 24))
            __trap();  // a lost MMA completion must fail loudly, never hang the device
    }
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows of 128 bytes, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor of tcgen05.mma kind::i8: D = s32, A = B = u8, both K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kChunkN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    // the wait is tied to the registers so that no use of them can be scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// S = sum_{c < 15} v[c] 2^(8c) (every v[c] < 2^23) reduced mod p: the value of BarrettUint128ModUint64 on the
// exact dot product (utilities-int.h:60-100), i.e. what the reference's Mul128 / accumulate loop yields.
__device__ __forceinline__ uint64_t fold_columns(const uint32_t* v, uint64_t p, uint64_t mu_hi, uint64_t mu_lo) {
    uint32_t e[8];
#pragma unroll
    for (int k = 0; k < 7; ++k)
        e[k] = v[2 * k] + (v[2 * k + 1] << 8);  // < 2^32: digit weights 2^(16k)
    e[7] = v[14];
    const uint64_t a_lo = (uint64_t)e[0] | ((uint64_t)e[2] << 32), a_hi = (uint64_t)e[4] | ((uint64_t)e[6] << 32);
    const uint64_t b_lo = (uint64_t)e[1] | ((uint64_t)e[3] << 32), b_hi = (uint64_t)e[5] | ((uint64_t)e[7] << 32);
    const unsigned __int128 s = (((unsigned __int128)a_hi << 64) | a_lo) + ((((unsigned __int128)b_hi << 64) | b_lo) << 16);
    return barrett128((uint64_t)(s >> 64), (uint64_t)s, p, mu_hi, mu_lo);
}

// ---- the kernel --------------------------------------------------------------------------------------------------
// dynamic shared memory (1024-byte aligned by hand): B image [kMaxChunks][32 KiB] | A tiles [2][16 KiB] | tables
struct TcTables {
    uint64_t dst[kMaxChunks * kChunkOut][4];  // p, mu128_hi, mu128_lo, output row
    uint64_t src[16][4];                      // q, QHatInv, Shoup(QHatInv), input row (or ~0 for padding)
    uint64_t bar[2];
    uint32_t tmem_base;
};
constexpr size_t kTcSmem = 1024 + (size_t)kMaxChunks * kChunkBytes + 2 * kTileBytes + sizeof(TcTables);

__global__ void __launch_bounds__(kTcThreads, 1) bconv_tc_kernel(RingDev r, TcLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem  = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_b   = smem;
    uint8_t* s_a   = smem + (size_t)kMaxChunks * kChunkBytes;
    TcTables* tb   = (TcTables*)(s_a + 2 * kTileBytes);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t row = tid & 127;       // coefficient of the tile: the same for the transpose and the epilogue
    const uint32_t sub = tid >> 7;        // which quarter of the source limbs / of a chunk's outputs
    const uint32_t N   = r.N;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(smem_u32(&tb->bar[0]), 1);
        mbar_init(smem_u32(&tb->bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tb->tmem_base;
    const uint32_t bar0 = smem_u32(&tb->bar[0]), bar1 = smem_u32(&tb->bar[1]);
    const uint32_t a_addr = smem_u32(s_a), b_addr = smem_u32(s_b);

    // contiguous share of the work list: at most a few job switches (= B reloads) per CTA
    const uint32_t W = L.job[L.njobs - 1].work_end;
    uint32_t w             = (uint32_t)(((uint64_t)W * blockIdx.x) / gridDim.x);
    const uint32_t w_end   = (uint32_t)(((uint64_t)W * (blockIdx.x + 1)) / gridDim.x);
    uint32_t ph0 = 0, ph1 = 0, cur = 0;
    int jb_loaded = -1;
    uint64_t nxt[4];

    // loads the four source residues of `row` for work item ww of job jb (Shoup-multiplied unless premultiplied)
    auto load_a = [&](const TcJob& J, uint32_t local, uint64_t (&y)[4]) {
        const uint32_t poly = local / L.tiles, tile = local - poly * L.tiles;
        const uint64_t* src = J.in + (size_t)poly * J.in_stride + (size_t)tile * kTileM + row;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t* s4 = tb->src[4 * sub + u];
            uint64_t v         = 0;
            if (s4[3] != ~0ull) {
                v = __ldg(src + (size_t)s4[3] * N);
                if (!L.premultiplied)
                    v = shoup(v, s4[1], s4[2], s4[0]);
            }
            y[u] = v;
        }
    };
    // [coefficient][limb bytes] image of the tile: row `row`, bytes [32 sub, 32 sub + 32), 128-byte swizzle
    auto store_a = [&](uint32_t buf, const uint64_t (&y)[4]) {
        uint8_t* base = s_a + buf * kTileBytes + row * 128;
        *reinterpret_cast<uint4*>(base + (((2 * sub) ^ (row & 7)) << 4)) =
            make_uint4((uint32_t)y[0], (uint32_t)(y[0] >> 32), (uint32_t)y[1], (uint32_t)(y[1] >> 32));
        *reinterpret_cast<uint4*>(base + (((2 * sub + 1) ^ (row & 7)) << 4)) =
            make_uint4((uint32_t)y[2], (uint32_t)(y[2] >> 32), (uint32_t)y[3], (uint32_t)(y[3] >> 32));
    };
    auto issue = [&](uint32_t buf, uint32_t chunk, uint32_t bar) {  // thread 0 only
        tc_fence_after();
#pragma unroll
        for (uint32_t k = 0; k < kK / 32; ++k)
            mma_i8(tmem + buf * kChunkN, smem_desc(a_addr + cur * kTileBytes + k * 32),
                   smem_desc(b_addr + chunk * kChunkBytes + k * 32), k);
        mma_commit(bar);
    };

    while (w < w_end) {
        int jb = 0;
        while (w >= L.job[jb].work_end)
            ++jb;
        const TcJob& J       = L.job[jb];
        const uint32_t wbase = jb ? L.job[jb - 1].work_end : 0;
        const uint32_t nout  = J.j1 - J.j0, nch = (nout + kChunkOut - 1) / kChunkOut;
        if (jb != jb_loaded) {
            // (re)load the job: tables, B image, first A tile -- everything of the previous job has been consumed
            const BConvTcDesc* d = J.d;
            if (tid < nch * kChunkOut) {
                const uint32_t j   = min(J.j0 + tid, J.j1 - 1);
                const ModConst& mc = r.mc[d->pidx[j]];
                tb->dst[tid][0] = mc.q, tb->dst[tid][1] = mc.mu128_hi, tb->dst[tid][2] = mc.mu128_lo, tb->dst[tid][3] = d->ppos[j];
            }
            if (tid >= 256 && tid < 256 + 16) {
                const uint32_t i = tid - 256;
                const bool ok    = i < d->sizeQ;
                tb->src[i][0]    = ok ? r.mc[d->qidx[i]].q : 1;
                tb->src[i][1]    = ok ? d->qhatinv[2 * i] : 0;
                tb->src[i][2]    = ok ? d->qhatinv[2 * i + 1] : 0;
                tb->src[i][3]    = ok ? (uint64_t)d->qpos[i] : ~0ull;
            }
            const uint4* bsrc = reinterpret_cast<const uint4*>(d->bmat + (size_t)(J.j0 / kChunkOut) * kChunkBytes);
            uint4* bdst       = reinterpret_cast<uint4*>(s_b);
            for (uint32_t e = tid; e < nch * (kChunkBytes / 16); e += kTcThreads)
                bdst[e] = __ldg(bsrc + e);
            __syncthreads();  // tables visible to load_a
            load_a(J, w - wbase, nxt);
            store_a(cur, nxt);
            fence_proxy_async();
            __syncthreads();
            jb_loaded = jb;
        }
        const uint32_t local = w - wbase;
        const uint32_t poly = local / L.tiles, tile = local - poly * L.tiles;
        uint64_t* outp = J.out + (size_t)poly * J.out_stride + (size_t)tile * kTileM + row;

        if (tid == 0) {
            issue(0, 0, bar0);
            if (nch > 1)
                issue(1, 1, bar1);
        }
        const bool have_next = (w + 1 < w_end) && (w + 1 < J.work_end);
        if (have_next)
            load_a(J, local + 1, nxt);  // in flight during the first epilogue

        for (uint32_t ch = 0; ch < nch; ++ch) {
            const uint32_t buf = ch & 1;
            if (buf == 0) {
                mbar_wait(bar0, ph0);
                ph0 ^= 1;
            }
            else {
                mbar_wait(bar1, ph1);
                ph1 ^= 1;
            }
            tc_fence_after();
            // this thread: outputs [kOutPerThr sub, kOutPerThr (sub + 1)) of the chunk, two at a time
#pragma unroll
            for (int pr = 0; pr < kOutPerThr / 2; ++pr) {
                const uint32_t jl = ch * kChunkOut + kOutPerThr * sub + 2 * pr;  // output index within the job
                if (jl < nout) {                                                 // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(tmem + (((warp & 3) * 32) << 16) + buf * kChunkN + (kOutPerThr * sub + 2 * pr) * 16, v);
#pragma unroll
                    for (int o = 0; o < 2; ++o)
                        if (jl + o < nout) {
                            const uint64_t* dd = tb->dst[jl + o];
                            uint64_t res       = fold_columns(v + 16 * o, dd[0], dd[1], dd[2]);
                            if (L.debug_col) {
                                res = 0;
#pragma unroll
                                for (int c = 0; c < 16; ++c)
                                    res = (L.debug_col == (uint32_t)c + 1) ? v[16 * o + c] : res;
                            }
                            outp[(size_t)dd[3] * N] = res;
                        }
                }
            }
            if (ch == 0 && have_next) {
                store_a(cur ^ 1, nxt);
                fence_proxy_async();
            }
            if (ch + 2 < nch) {
                tc_fence_before();
                __syncthreads();
                if (tid == 0)
                    issue(buf, ch + 2, buf ? bar1 : bar0);
            }
        }
        tc_fence_before();
        __syncthreads();  // both accumulator buffers drained, the next A tile visible to the tensor core
        if (have_next)
            cur ^= 1;
        else
            jb_loaded = -1;  // next item (if any) belongs to another job: full reload
        ++w;
    }

    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// B image of a conversion table, built on the device from the 30-bit split matrix: chunk ch = targets [16 ch, 16 ch+16),
// row n = 16 (j mod 16) + c, byte k = 8 i + a holds byte (c - a) of M[i][j]; rows of 128 bytes, 8-row atoms, 16-byte
// units XOR-swizzled with the row (the SWIZZLE_128B layout the MMA descriptor names).
__global__ void bconv_tc_build_kernel(const uint2* __restrict__ mat, uint32_t sizeQ, uint32_t sizeP, uint32_t nch,
                                      uint8_t* __restrict__ bmat) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nch * kChunkBytes)
        return;
    const uint32_t ch = e / kChunkBytes, rem = e % kChunkBytes, n = rem / kK, k = rem % kK;
    const uint32_t j = ch * kChunkOut + (n >> 4), c = n & 15, i = k >> 3, a = k & 7;
    uint8_t val = 0;
    if (j < sizeP && i < sizeQ && c >= a && c - a < 8) {
        const uint2 m    = mat[(size_t)i * sizeP + j];
        const uint64_t x = (uint64_t)m.x | ((uint64_t)m.y << 30);
        val              = (uint8_t)(x >> (8 * (c - a)));
    }
    bmat[(size_t)ch * kChunkBytes + (n >> 3) * 1024 + (n & 7) * 128 + ((((k >> 4) ^ (n & 7)) << 4) | (k & 15))] = val;
}

bool tc_enabled() {
    static const bool on = [] {
        const char* e = getenv("OFHE_BCONV_TC");  // 0: keep every conversion on the CUDA-core kernel
        return !(e && atoi(e) == 0);
    }();
    return on;
}

}  // namespace

void bconv_finalize(const Ring& r, BConv& t) {
    if (t.sizeQ > 16 || t.sizeQ == 0 || t.sizeP == 0 || r.N < (uint32_t)kTileM)
        return;
    const uint32_t nch = (t.sizeP + kChunkOut - 1) / kChunkOut;
    OFHE_CUDA(cudaMalloc(&t.tc_bmat_d, (size_t)nch * kChunkBytes));
    const uint32_t total = nch * kChunkBytes;
    bconv_tc_build_kernel<<<(total + 255) / 256, 256, 0, cudaStreamLegacy>>>(reinterpret_cast<const uint2*>(t.mat_d), t.sizeQ,
                                                                             t.sizeP, nch, t.tc_bmat_d);
    OFHE_CUDA(cudaGetLastError());
    BConvTcDesc h{};
    h.sizeQ   = t.sizeQ;
    h.sizeP   = t.sizeP;
    h.bmat    = t.tc_bmat_d;
    h.qhatinv = t.qhatinv_d;
    for (uint32_t i = 0; i < t.sizeQ; ++i)
        h.qidx[i] = t.qmap.idx[i], h.qpos[i] = t.qmap.pos[i];
    for (uint32_t j = 0; j < t.sizeP; ++j)
        h.pidx[j] = t.pmap.idx[j], h.ppos[j] = t.pmap.pos[j];
    OFHE_CUDA(cudaMalloc(&t.tc_desc_d, sizeof(BConvTcDesc)));
    h2d_blocking(t.tc_desc_d, &h, sizeof(h));  // also orders the build kernel (legacy stream) before any caller stream
}

bool bconv_tc_eligible(const Ring& r, const BConv& t) {
    return tc_enabled() && t.mode == 0 && t.tc_desc_d != nullptr && r.N >= (uint32_t)kTileM;
}

void launch_bconv_tc(const Ring& r, const BConvJob* jobs, uint32_t njobs, bool premultiplied, cudaStream_t st) {
    static const bool attr_set = [] {
        OFHE_CUDA(cudaFuncSetAttribute(bconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        return true;
    }();
    (void)attr_set;
    TcLaunch L{};
    L.tiles         = r.N / kTileM;
    L.premultiplied = premultiplied ? 1u : 0u;
    static const bool dbg = getenv("OFHE_BCONV_TC_DEBUG") != nullptr;
    if (dbg) {
        const char* c = getenv("OFHE_BCONV_TC_COL");
        L.debug_col   = c ? (uint32_t)atoi(c) : 0u;
    }
    auto flush      = [&] {
        if (!L.njobs)
            return;
        const uint32_t W    = L.job[L.njobs - 1].work_end;
        const uint32_t grid = W < (uint32_t)r.ctx->sm_count ? W : (uint32_t)r.ctx->sm_count;
        bconv_tc_kernel<<<grid, kTcThreads, kTcSmem, st>>>(r.dev(), L);
        OFHE_CUDA(cudaGetLastError());
        L.njobs = 0;
    };
    for (uint32_t k = 0; k < njobs; ++k) {
        const BConvJob& b = jobs[k];
        OFHE_REQUIRE(b.t && b.t->tc_desc_d && b.t->mode == 0, "tensor-core base conversion: table not eligible");
        if (b.batch == 0)
            continue;
        for (uint32_t j0 = 0; j0 < b.t->sizeP; j0 += kMaxChunks * kChunkOut) {  // <= 80 target limbs per job
            if (L.njobs == kMaxJobs)
                flush();
            TcJob& J     = L.job[L.njobs];
            J.d          = b.t->tc_desc_d;
            J.in         = b.in;
            J.out        = b.out;
            J.in_stride  = b.in_stride;
            J.out_stride = b.out_stride;
            J.batch      = b.batch;
            J.j0         = j0;
            J.j1         = j0 + kMaxChunks * kChunkOut < b.t->sizeP ? j0 + kMaxChunks * kChunkOut : b.t->sizeP;
            J.work_end   = (L.njobs ? L.job[L.njobs - 1].work_end : 0) + b.batch * L.tiles;
            ++L.njobs;
        }
    }
    flush();
}

}  // namespace ofhe
