// baseconv_tc.cu -- RNS fast base conversion (ApproxSwitchCRTBasis, dcrtpoly-impl.h:888-932) on the 5th-generation
// tensor cores: the dense INTEGER contraction of the hybrid key switch.
//
//     out_j[r] = ( sum_i y_i[r] * M[i][j] )  mod p_j,      y_i = x_i * QHatInvModq_i mod q_i,  M[i][j] = QHatModp[i][j]
//
// is, for every target limb j, an exact 124-bit integer dot product over <= 16 source limbs followed by one
// reduction mod p_j.  The CUDA-core kernel (baseconv.cu) spends 3 IMAD.WIDE + a 64-bit add per 60x60-bit MAC
// (215 SASS instructions per output residue, measured).  Here the dot products run on tcgen05.mma kind::i8:
//
//   * residues and matrix entries are split into their 8 bytes -- no work, the bytes of the 64-bit words ARE the
//     digits.  A tile of 128 coefficients becomes the K-major u8 operand A[128][K = 16 limbs x 8 bytes];
//   * the matrix becomes B[(i,a)][(j,c)] = byte (c-a) of M'[i][j]  (0 <= c-a < 8, else 0), c = 0..15: sixteen u8
//     columns per target limb, one per digit weight 2^(8c) -- built once per conversion table, device resident in the
//     exact shared-memory image (128-byte swizzle) the MMA reads.  M'[i][j] = M[i][j] * 2^64 mod p_j: the matrix is
//     kept in Montgomery form so that the reduction below needs no quotient estimate;
//   * D[r][(j,c)] = sum_k A[r][k] B[k][(j,c)] <= 120 * 255^2 < 2^23 in s32, exact, accumulated in TMEM;
//   * the epilogue reads the 15 column sums of an output with tcgen05.ld (one TMEM lane = one coefficient = one
//     thread), rebuilds S' = sum_c D_c 2^(8c) = sum_i y_i M'[i][j] < 2^124 with 7 shift-adds, 4 funnel shifts and one
//     128-bit add, and takes ONE 64-bit Montgomery step: t = (S' + (lo(S') * (-p^-1) mod 2^64) * p) / 2^64 = S (mod p)
//     with t < 2p because S' < 16 * 2^60 * p, then one conditional subtraction: the canonical residue, bit for bit
//     the value of the reference's 128-bit accumulate + BarrettUint128ModUint64.
//
// One persistent CTA per SM, 17 warps: all 512 TMEM columns as two 256-column accumulator buffers (16 target limbs
// each), the whole B image of the job (<= 5 chunks of 32 KiB) and two 16 KiB A tiles in shared memory.  Warp 16 issues
// the MMAs (4 x M128 N256 K32 per chunk, tcgen05.commit onto the buffer's `full` mbarrier); warps 0..15 run the
// epilogue of one buffer while the tensor core fills the other, and transpose the next tile ([limb][coefficient] in
// HBM -> [coefficient][limb bytes] in shared memory) through registers.  The warps meet only through mbarriers; a
// job's tables and B image arrive by TMA bulk copies.  Several conversions (the digits of a ModUp, the two
// polynomials of a ModDown) are one launch: a job list, work = (job, polynomial, tile), a contiguous share per CTA.
// Measured on B200 (profiles/r02s_*): 15 -> 45 limbs at N = 2^16 in 20.9 us (CUDA cores: 50.2 us); the MMAs occupy the
// tensor pipe 12 % of the time (kind::i8 issues at 8174 MAC/clk/SM, profiles/r02r_mma_rate.txt), the kernel is bound
// by integer issue in the epilogue at ~52 SASS instructions per output residue.
//
// Eligibility: mode 0 (approximate conversion), <= 16 source limbs, N >= 128.  Everything else stays on baseconv.cu.
#include <atomic>
#include <cstdlib>
#include <vector>

#include "internal.h"

namespace ofhe {
namespace {

constexpr int kEpiThreads = 512;                  // 16 transpose / epilogue warps
constexpr int kTcThreads  = kEpiThreads + 32;     // + the MMA warp
constexpr int kTileM      = 128;                  // coefficients per tile = TMEM lanes
#ifndef OFHE_TC_NBUF
#define OFHE_TC_NBUF 2  // accumulator buffers in TMEM: 2 x 256 columns (16 target limbs each) or 4 x 128 (8 each).
                        // Measured (B200, 15 -> 45 limbs, N = 2^16): 20.9 us with 2, 22.9 us with 4 -- the extra slack for the
                        // MMA warp costs twice the barrier traffic and halves the outputs per TMEM load; both bit-exact.
#endif
constexpr int kNBuf       = OFHE_TC_NBUF;
static_assert(kNBuf == 2 || kNBuf == 4, "two or four accumulator buffers");
constexpr int kChunkN     = 512 / kNBuf;          // columns per buffer: target limbs x 16 digit weights
constexpr int kChunkOut   = kChunkN / 16;         // target limbs per accumulator buffer
constexpr int kK          = 128;                  // 16 source limbs x 8 bytes
constexpr int kChunkBytes = kChunkN * kK;         // 32 KiB of B per chunk
constexpr int kTileBytes  = kTileM * kK;          // 16 KiB of A per tile
constexpr int kMaxChunks  = 80 / kChunkOut;       // 80 target limbs per job (larger conversions are split into jobs)
constexpr int kMaxJobs    = 4;
constexpr int kSub        = kEpiThreads / 128;    // warps per TMEM lane quarter
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
// hardware-suspended try_wait (10 ms hint); a completion that never comes (a bug) traps after a few seconds instead
// of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && spin > 400)
            __trap();
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared 1-D bulk copy (TMA), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
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
// 32 consecutive columns (two outputs' digit-weight sums) of this thread's TMEM lane; completion via tmem_wait32
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
}
// tcgen05.wait::ld covers every load of the thread issued so far; it is tied to the registers it releases so that
// no use of them can be scheduled above it
__device__ __forceinline__ void tmem_wait32(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// S' = sum_{c < 15} v[c] 2^(8c) (every v[c] < 2^23, S' < 2^64 p) taken out of Montgomery form: S' 2^-64 mod p, canonical.
// (S' + u p) is a multiple of 2^64 for u = lo(S') * ninv, ninv = -p^-1 mod 2^64; its quotient is hi(S') + hi(u p) +
// [lo(S') != 0] < 2p.
__device__ __forceinline__ uint64_t fold_columns(const uint32_t* v, uint64_t p, uint64_t ninv) {
    uint32_t e[8];
#pragma unroll
    for (int k = 0; k < 7; ++k)
        e[k] = v[2 * k] + (v[2 * k + 1] << 8);  // < 2^32: digit weights 2^(16k)
    e[7] = v[14];
    const uint64_t a_lo = (uint64_t)e[0] | ((uint64_t)e[2] << 32), a_hi = (uint64_t)e[4] | ((uint64_t)e[6] << 32);
    const uint64_t b_lo = (uint64_t)e[1] | ((uint64_t)e[3] << 32), b_hi = (uint64_t)e[5] | ((uint64_t)e[7] << 32);
    const unsigned __int128 s = (((unsigned __int128)a_hi << 64) | a_lo) + ((((unsigned __int128)b_hi << 64) | b_lo) << 16);
    const uint64_t lo = (uint64_t)s, hi = (uint64_t)(s >> 64);
    const uint64_t t  = hi + __umul64hi(lo * ninv, p) + (lo != 0);
    return t >= p ? t - p : t;
}

#ifndef OFHE_TC_DB
#define OFHE_TC_DB 0  // 1: the second output pair's TMEM load in flight during the first pair's reduction (64 live registers)
#endif
// The same residue with about half the multiplier work, for moduli of at least 32 bits: the two upper 32-bit words of
// S are folded down with c2 = 2^64 mod p and c3 = 2^96 mod p (two 32 x 60-bit products each),
//     T = (s1:s0) + s2 c2 + s3 c3  <  2^94,   T = S (mod p),
// and T takes the single-word Barrett step of modarith.cuh (quotient estimate within 2 for T < 2^(63+k), k >= 31).
__device__ __forceinline__ uint64_t fold_columns_fast(const uint32_t* v, uint64_t p, uint64_t mu, uint32_t k, uint64_t c2,
                                                      uint64_t c3) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 7; ++i)
        e[i] = v[2 * i] + (v[2 * i + 1] << 8);
    e[7] = v[14];
    const uint64_t a_lo = (uint64_t)e[0] | ((uint64_t)e[2] << 32), a_hi = (uint64_t)e[4] | ((uint64_t)e[6] << 32);
    const uint64_t b_lo = (uint64_t)e[1] | ((uint64_t)e[3] << 32), b_hi = (uint64_t)e[5] | ((uint64_t)e[7] << 32);
    const unsigned __int128 s = (((unsigned __int128)a_hi << 64) | a_lo) + ((((unsigned __int128)b_hi << 64) | b_lo) << 16);
    const uint64_t s_hi = (uint64_t)(s >> 64);
    const uint32_t s2 = (uint32_t)s_hi, s3 = (uint32_t)(s_hi >> 32);
    const uint64_t lo_part = (uint64_t)s2 * (uint32_t)c2;                                            // < 2^64
    const uint64_t lo_par2 = (uint64_t)s3 * (uint32_t)c3;
    const uint64_t hi_part = (uint64_t)s2 * (uint32_t)(c2 >> 32) + (uint64_t)s3 * (uint32_t)(c3 >> 32);  // < 2^61
    const unsigned __int128 t = (unsigned __int128)(uint64_t)s + lo_part + lo_par2 + ((unsigned __int128)hi_part << 32);
    return barrett_reduce_prod((uint64_t)(t >> 64), (uint64_t)t, p, mu, k);
}

// ---- the kernel --------------------------------------------------------------------------------------------------
// dynamic shared memory (1024-byte aligned by hand): B image [kMaxChunks][32 KiB] | A tiles [2][16 KiB] | tables
struct TcTables {
    uint64_t dst[kMaxChunks * kChunkOut][4];  // p, -p^-1 mod 2^64, word offset of the output row, (pad)
    uint64_t src[16][4];                      // q, QHatInv, Shoup(QHatInv), word offset of the input row (~0: padding)
    uint64_t full[kNBuf];                     // accumulator buffer written   (1 arrival: tcgen05.commit)
    uint64_t empty[kNBuf];                    // accumulator buffer drained   (16 arrivals: one per epilogue warp)
    uint64_t afull[2];                        // A tile staged                (16 arrivals)
    uint64_t job;                             // tables + B image of a job landed (bulk copies, byte count)
    uint32_t tmem_base;
};
constexpr size_t kTcSmem = 1024 + (size_t)kMaxChunks * kChunkBytes + 2 * kTileBytes + sizeof(TcTables);

// Warps 0..15: transpose + epilogue.  Warp 16: MMA issue.  In steady state the warps meet only through mbarriers
// (full / empty per accumulator buffer, afull per A tile), so a warp that finishes a chunk early runs ahead into the
// next one and the tensor core works across tile boundaries; block-wide barriers happen only when the job changes.
__global__ void __launch_bounds__(kTcThreads, 1) bconv_tc_kernel(RingDev r, TcLaunch L) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem  = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* s_b   = smem;
    uint8_t* s_a   = smem + (size_t)kMaxChunks * kChunkBytes;
    TcTables* tb   = reinterpret_cast<TcTables*>(s_a + 2 * kTileBytes);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma  = warp == kEpiThreads / 32;
    const uint32_t row = tid & 127;       // coefficient of the tile: the same for the transpose and the epilogue
    const uint32_t sub = (tid >> 7) & 3;  // which quarter of the source limbs / of a chunk's outputs

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int b = 0; b < kNBuf; ++b) {
            mbar_init(smem_u32(&tb->full[b]), 1);
            mbar_init(smem_u32(&tb->empty[b]), kEpiThreads / 32);
        }
        for (int b = 0; b < 2; ++b)
            mbar_init(smem_u32(&tb->afull[b]), kEpiThreads / 32);
        mbar_init(smem_u32(&tb->job), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem   = tb->tmem_base;
    const uint32_t full0  = smem_u32(&tb->full[0]), empty0 = smem_u32(&tb->empty[0]), afull0 = smem_u32(&tb->afull[0]);
    const uint32_t a_addr = smem_u32(s_a), b_addr = smem_u32(s_b);

    // contiguous share of the work list: at most a few job switches (= B reloads) per CTA
    const uint32_t W     = L.job[L.njobs - 1].work_end;
    uint32_t w           = (uint32_t)(((uint64_t)W * blockIdx.x) / gridDim.x);
    const uint32_t w_end = (uint32_t)(((uint64_t)W * (blockIdx.x + 1)) / gridDim.x);
    const uint32_t jobbar = smem_u32(&tb->job);
    uint32_t g = 0, tcnt = 0, jcnt = 0;  // chunks / tiles / jobs this CTA has been through: accumulator buffer g mod kNBuf, A buffer tcnt & 1
    int jb_loaded = -1;
    uint64_t nxt[4];

    // the four source residues of `row` for item `local` of job J: loads only, so that a prefetch never waits
    auto load_a = [&](const TcJob& J, uint32_t local, uint64_t (&y)[4]) {
        const uint32_t poly = local / L.tiles, tile = local - poly * L.tiles;
        const uint64_t* src = J.in + (size_t)poly * J.in_stride + (size_t)tile * kTileM + row;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t off = tb->src[4 * sub + u][3];
            y[u]               = off != ~0ull ? __ldg(src + off) : 0;
        }
    };
    // [coefficient][limb bytes] image of the tile: row `row`, bytes [32 sub, 32 sub + 32), 128-byte swizzle (Shoup-
    // multiplied by QHatInvModq first unless the caller folded that into the inverse NTT); then this warp's share
    // of the tile is announced to the MMA warp
    auto store_a = [&](uint32_t buf, uint64_t (&y)[4]) {
        if (!L.premultiplied) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t* s4 = tb->src[4 * sub + u];
                y[u]               = shoup(y[u], s4[1], s4[2], s4[0]);  // padding rows: 0 * 0
            }
        }
        uint8_t* base = s_a + buf * kTileBytes + row * 128;
        *reinterpret_cast<uint4*>(base + (((2 * sub) ^ (row & 7)) << 4)) =
            make_uint4((uint32_t)y[0], (uint32_t)(y[0] >> 32), (uint32_t)y[1], (uint32_t)(y[1] >> 32));
        *reinterpret_cast<uint4*>(base + (((2 * sub + 1) ^ (row & 7)) << 4)) =
            make_uint4((uint32_t)y[2], (uint32_t)(y[2] >> 32), (uint32_t)y[3], (uint32_t)(y[3] >> 32));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0)
            mbar_arrive(afull0 + 8 * buf);
    };

    while (w < w_end) {
        int jb = 0;
        while (w >= L.job[jb].work_end)
            ++jb;
        const TcJob& J       = L.job[jb];
        const uint32_t wbase = jb ? L.job[jb - 1].work_end : 0;
        const uint32_t nout  = J.j1 - J.j0, nch = (nout + kChunkOut - 1) / kChunkOut;
        const uint32_t local = w - wbase;
        if (jb != jb_loaded) {
            // (re)load the job.  Every warp is past its last wait on the previous job's accumulators, so all of that
            // job's MMAs have completed and nobody reads its tables any more once the barrier is passed.  Tables and
            // B image arrive by bulk copies on one mbarrier; nobody stages them through registers.
            __syncthreads();
            const BConvTcDesc* d = J.d;
            if (tid == kEpiThreads) {
                fence_proxy_async();  // earlier generic-proxy reads of the region vs the copy engine's writes
                mbar_expect_tx(jobbar, nch * kChunkBytes + nout * 32 + 512);
                for (uint32_t c = 0; c < nch; ++c)
                    bulk_load(b_addr + c * kChunkBytes, d->bmat + (size_t)(J.j0 / kChunkOut + c) * kChunkBytes, kChunkBytes, jobbar);
                bulk_load(smem_u32(&tb->dst[0][0]), d->dst_tab + (size_t)J.j0 * 4, nout * 32, jobbar);
                bulk_load(smem_u32(&tb->src[0][0]), &d->src_tab[0][0], 512, jobbar);
            }
            mbar_wait(jobbar, jcnt & 1);
            ++jcnt;
            if (!is_mma) {
                load_a(J, local, nxt);
                store_a(tcnt & 1, nxt);
            }
            jb_loaded = jb;
        }
        const bool have_next = (w + 1 < w_end) && (w + 1 < J.work_end);

        if (is_mma) {
            if (lane == 0) {
                mbar_wait(afull0 + 8 * (tcnt & 1), (tcnt >> 1) & 1);
                for (uint32_t ch = 0; ch < nch; ++ch, ++g) {
                    const uint32_t b = g % kNBuf;
                    mbar_wait(empty0 + 8 * b, ((g / kNBuf) & 1) ^ 1);  // the first use of each buffer passes at once
                    tc_fence_after();
#pragma unroll
                    for (uint32_t k = 0; k < kK / 32; ++k)
                        mma_i8(tmem + b * kChunkN, smem_desc(a_addr + (tcnt & 1) * kTileBytes + k * 32),
                               smem_desc(b_addr + ch * kChunkBytes + k * 32), k);
                    mma_commit(full0 + 8 * b);
                }
            }
            __syncwarp();
        }
        else {
            const uint32_t poly = local / L.tiles, tile = local - poly * L.tiles;
            uint64_t* outp      = J.out + (size_t)poly * J.out_stride + (size_t)tile * kTileM + row;
            if (have_next)
                load_a(J, local + 1, nxt);  // in flight during the first epilogue
            for (uint32_t ch = 0; ch < nch; ++ch, ++g) {
                const uint32_t b = g % kNBuf;
                mbar_wait(full0 + 8 * b, (g / kNBuf) & 1);
                tc_fence_after();
                // this thread: outputs [4 sub, 4 sub + 4) of the chunk, as two pairs; the second pair's load is in
                // flight while the first pair is folded and reduced, and the two reductions of a pair are independent
                // instruction streams for the scheduler
                const uint32_t jl0 = ch * kChunkOut + kOutPerThr * sub;  // index within the job; warp-uniform
                const int nv       = jl0 >= nout ? 0 : (int)min((uint32_t)kOutPerThr, nout - jl0);
                const uint32_t t0  = tmem + (((warp & 3) * 32) << 16) + b * kChunkN + kOutPerThr * sub * 16;
                auto release = [&] {  // every load of this buffer has landed in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(empty0 + 8 * b);
                };
                // both reductions of a pair sit in ONE basic block (the path is chosen once per pair), so that the
                // scheduler interleaves the two dependency chains
                auto emit2 = [&](const uint32_t (&v)[32], uint32_t jl, bool both) {
                    const uint64_t* d0 = tb->dst[jl];
                    const uint64_t* d1 = tb->dst[both ? jl + 1 : jl];
                    uint64_t r0 = fold_columns(v, d0[0], d0[1]);
                    uint64_t r1 = fold_columns(v + 16, d1[0], d1[1]);
                    if (L.debug_col) {
                        r0 = r1 = 0;
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (L.debug_col == (uint32_t)c + 1)
                                r0 = v[c], r1 = v[16 + c];
                    }
                    outp[d0[2]] = r0;
                    if (both)
                        outp[d1[2]] = r1;
                };
#if OFHE_TC_DB
                uint32_t va[32], vb[32];
                if (nv > 0) {
                    tmem_ld32(t0, va);
                    tmem_wait32(va);
                    if (nv > 2)
                        tmem_ld32(t0 + 32, vb);
                    else
                        release();
                    emit2(va, jl0, nv > 1);
                    if (nv > 2) {
                        tmem_wait32(vb);
                        release();
                        emit2(vb, jl0 + 2, nv > 3);
                    }
                }
                else
                    release();
#else
                uint32_t va[32];
                if (nv > 0) {
                    tmem_ld32(t0, va);
                    tmem_wait32(va);
                    if (nv <= 2)
                        release();
                    emit2(va, jl0, nv > 1);
                    if (nv > 2) {
                        tmem_ld32(t0 + 32, va);
                        tmem_wait32(va);
                        release();
                        emit2(va, jl0 + 2, nv > 3);
                    }
                }
                else
                    release();
#endif
                if (ch == 0 && have_next)
                    store_a((tcnt + 1) & 1, nxt);
            }
        }
        if (!have_next)
            jb_loaded = -1;  // the next item (if any) belongs to another job: full reload
        ++tcnt;
        ++w;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// B image of a conversion table, built on the device from the 30-bit split matrix: chunk ch = targets [16 ch, 16 ch+16),
// row n = 16 (j mod 16) + c, byte k = 8 i + a holds byte (c - a) of M'[i][j] = M[i][j] * 2^64 mod p_j; rows of 128
// bytes, 8-row atoms, 16-byte units XOR-swizzled with the row (the SWIZZLE_128B layout the MMA descriptor names).
__global__ void bconv_tc_build_kernel(RingDev r, const uint2* __restrict__ mat, const uint16_t* __restrict__ pidx,
                                      uint32_t sizeQ, uint32_t sizeP, uint32_t nch, uint8_t* __restrict__ bmat) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nch * kChunkBytes)
        return;
    const uint32_t ch = e / kChunkBytes, rem = e % kChunkBytes, n = rem / kK, k = rem % kK;
    const uint32_t j = ch * kChunkOut + (n >> 4), c = n & 15, i = k >> 3, a = k & 7;
    uint8_t val = 0;
    if (j < sizeP && i < sizeQ && c >= a && c - a < 8) {
        const ModConst mc  = r.mc[pidx[j]];
        const uint2 m      = mat[(size_t)i * sizeP + j];
        const uint64_t x   = (uint64_t)m.x | ((uint64_t)m.y << 30);
        const uint64_t r64 = barrett128(1, 0, mc.q, mc.mu128_hi, mc.mu128_lo);  // 2^64 mod p
        val                = (uint8_t)(mulmod(x, r64, mc) >> (8 * (c - a)));
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
    uint16_t* pidx_d     = nullptr;
    OFHE_CUDA(cudaMalloc(&pidx_d, sizeof(uint16_t) * t.sizeP));
    h2d_blocking(pidx_d, t.pmap.idx, sizeof(uint16_t) * t.sizeP);
    bconv_tc_build_kernel<<<(total + 255) / 256, 256, 0, cudaStreamLegacy>>>(r.dev(), reinterpret_cast<const uint2*>(t.mat_d), pidx_d,
                                                                             t.sizeQ, t.sizeP, nch, t.tc_bmat_d);
    OFHE_CUDA(cudaGetLastError());
    OFHE_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    OFHE_CUDA(cudaFree(pidx_d));
    // per-limb tables, host computed: everything a job needs arrives in shared memory by three bulk copies
    std::vector<uint64_t> hinv(2 * (size_t)t.sizeQ);
    OFHE_CUDA(cudaMemcpy(hinv.data(), t.qhatinv_d, sizeof(uint64_t) * hinv.size(), cudaMemcpyDeviceToHost));
    BConvTcDesc h{};
    h.sizeQ = t.sizeQ;
    h.sizeP = t.sizeP;
    h.bmat  = t.tc_bmat_d;
    for (uint32_t i = 0; i < 16; ++i) {
        const bool ok   = i < t.sizeQ;
        h.src_tab[i][0] = ok ? r.moduli[t.qmap.idx[i]] : 1;
        h.src_tab[i][1] = ok ? hinv[2 * i] : 0;
        h.src_tab[i][2] = ok ? hinv[2 * i + 1] : 0;
        h.src_tab[i][3] = ok ? (uint64_t)t.qmap.pos[i] * r.N : ~0ull;
    }
    std::vector<uint64_t> dst(4 * (size_t)t.sizeP);
    for (uint32_t j = 0; j < t.sizeP; ++j) {
        const uint64_t p = r.moduli[t.pmap.idx[j]];
        OFHE_REQUIRE(p & 1, "tensor-core base conversion: moduli must be odd");
        uint64_t inv = p;  // Newton: p * inv = 1 mod 2^64 (p odd: 3 correct bits to start with, doubled 5 times)
        for (int it = 0; it < 5; ++it)
            inv *= 2 - p * inv;
        uint64_t* dd = &dst[4 * (size_t)j];
        dd[0] = p, dd[1] = 0 - inv, dd[2] = (uint64_t)t.pmap.pos[j] * r.N, dd[3] = 0;
    }
    OFHE_CUDA(cudaMalloc(&t.tc_dst_tab_d, sizeof(uint64_t) * dst.size()));
    h2d_blocking(t.tc_dst_tab_d, dst.data(), sizeof(uint64_t) * dst.size());
    h.dst_tab = t.tc_dst_tab_d;
    OFHE_CUDA(cudaMalloc(&t.tc_desc_d, sizeof(BConvTcDesc)));
    h2d_blocking(t.tc_desc_d, &h, sizeof(h));  // also orders the build kernel (legacy stream) before any caller stream
}

bool bconv_tc_eligible(const Ring& r, const BConv& t) {
    return tc_enabled() && t.mode == 0 && t.tc_desc_d != nullptr && r.N >= (uint32_t)kTileM;
}

void launch_bconv_tc(const Ring& r, const BConvJob* jobs, uint32_t njobs, bool premultiplied, cudaStream_t st) {
    // the opt-in for > 48 KiB of dynamic shared memory is per device: once for every device this process launches on
    static std::atomic<bool> attr_set[64];
    const int dev = r.ctx->device;
    if (dev < 0 || dev >= 64 || !attr_set[dev].load(std::memory_order_acquire)) {
        OFHE_CUDA(cudaFuncSetAttribute(bconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
        if (dev >= 0 && dev < 64)
            attr_set[dev].store(true, std::memory_order_release);
    }
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
