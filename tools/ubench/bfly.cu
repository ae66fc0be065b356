// Microbenchmark: issue cost of the 60-bit NTT butterfly variants on sm_100a, registers only (no memory in the loop).
// Reports cycles per warp-butterfly per SM sub-partition at 3 CTAs x 256 threads per SM (the NTT kernels' occupancy),
// plus raw IMAD.WIDE / IMAD.HI / IMAD throughput with DISTINCT operands (the r01 pipes.cu reused one operand pair).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../openfhe-development_b200/csrc -o bfly bfly.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "ntt_core.cuh"

using namespace ofhe;
using namespace ofhe::nttc;

constexpr int kIters = 512;

// ---- variants of the twiddle product -------------------------------------------------------------------
__device__ __forceinline__ uint64_t shoup_ptx(uint64_t y, uint64_t w, uint64_t ws, uint64_t nq) {
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
}
// quotient from mul.hi (IMAD.HI) of the two cross products
__device__ __forceinline__ uint64_t shoup_hi(uint64_t y, uint64_t w, uint64_t ws, uint64_t nq) {
    const uint32_t y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32), s0 = (uint32_t)ws, s1 = (uint32_t)(ws >> 32);
    const uint64_t h = (uint64_t)y1 * s1 + __umulhi(y0, s1) + __umulhi(y1, s0);
    return y * w + h * nq;
}

template <int V>
__device__ __forceinline__ void bfly(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m, bool) {}

// V0: round-1 butterfly (exact quotient, [0,4q), compiler-generated conditional subtraction every stage)
template <>
__device__ __forceinline__ void bfly<0>(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m, bool) {
    const uint64_t x = X >= m.c2 ? X - m.c2 : X;
    const uint64_t v = Y * t.w + __umul64hi(Y, t.ws) * m.nq;
    X                = x + v;
    Y                = x - v + m.c2;
}
// V1: round-2 butterfly as shipped (truncated quotient in C, 4-instruction csub on alternate stages)
template <>
__device__ __forceinline__ void bfly<1>(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m, bool cs) {
    const uint64_t x = cs ? csub4i(X, m.c8) : X;
    const uint64_t v = shoup_t(Y, t.w, t.ws, m.nq);
    X                = x + v;
    Y                = x - v + m.c4;
}
// V2: same with the product in hand-written PTX
template <>
__device__ __forceinline__ void bfly<2>(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m, bool cs) {
    const uint64_t x = cs ? csub4i(X, m.c8) : X;
    const uint64_t v = shoup_ptx(Y, t.w, t.ws, m.nq);
    X                = x + v;
    Y                = x - v + m.c4;
}
// V3: cross products through mul.hi
template <>
__device__ __forceinline__ void bfly<3>(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m, bool cs) {
    const uint64_t x = cs ? csub4i(X, m.c8) : X;
    const uint64_t v = shoup_hi(Y, t.w, t.ws, m.nq);
    X                = x + v;
    Y                = x - v + m.c4;
}
// V4: V1 with a conditional subtraction on every stage
template <>
__device__ __forceinline__ void bfly<4>(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m, bool) {
    const uint64_t x = csub4i(X, m.c4);
    const uint64_t v = shoup_t(Y, t.w, t.ws, m.nq);
    X                = x + v;
    Y                = x - v + m.c4;
}
// V5: inverse (Gentleman-Sande) butterfly as shipped; V6: round-1 inverse butterfly
template <>
__device__ __forceinline__ void bfly<5>(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m, bool) {
    gs_bfly(X, Y, t, m);
}
template <>
__device__ __forceinline__ void bfly<6>(uint64_t& X, uint64_t& Y, const Twiddle& t, const Mod& m, bool) {
    const uint64_t s = X + Y;
    const uint64_t d = X - Y + m.c2;
    X                = s >= m.c2 ? s - m.c2 : s;
    Y                = d * t.w + __umul64hi(d, t.ws) * m.nq;
}

template <int V, int MINB>
__global__ void __launch_bounds__(256, MINB) kb(uint64_t* data, const Twiddle* tw, ModConst mc, long long* cyc) {
    const Mod m = make_mod(mc);
    uint64_t x[16];
    uint64_t* p = data + (size_t)blockIdx.x * 4096 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        x[i] = p[i * 256];
    Twiddle t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        t[i] = ld_tw(tw + i + (threadIdx.x & 3));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int D = 0; D < 4; ++D) {
            const int half = 8 >> D;
#pragma unroll
            for (int g = 0; g < (1 << D); ++g)
#pragma unroll
                for (int l = 0; l < half; ++l)
                    bfly<V>(x[(g << (4 - D)) + l], x[(g << (4 - D)) + l + half], t[D], m, (D & 1) != 0);
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i)
        p[i * 256] = x[i];
    if (threadIdx.x == 0)
        cyc[blockIdx.x] = t1 - t0;
}

// ---- raw instruction throughput with distinct operands ---------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 3) kraw(uint64_t* out, long long* cyc) {
    uint32_t a[8], b[8];
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i]   = threadIdx.x * 2654435761u + i * 40503u + 1;
        b[i]   = threadIdx.x * 2246822519u + i * 9973u + 7;
        acc[i] = i;
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)  // IMAD.WIDE.U32 with 64-bit accumulate, distinct multiplicands
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[(i + 1) & 7]));
            if (MODE == 1) {  // IMAD.WIDE.U32 without accumulate feeding the next one's multiplicand
                uint64_t r;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a[i]), "r"(b[(i + 1) & 7]));
                a[i] = (uint32_t)(r >> 32) | 1u;
                acc[i] = r;
            }
            if (MODE == 2) {  // IMAD.HI.U32
                uint32_t r;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b[(i + 1) & 7]));
                a[i] = r | 1u;
            }
            if (MODE == 3) {  // IMAD (32-bit) with accumulate
                uint32_t lo = (uint32_t)acc[i];
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a[i]), "r"(b[(i + 1) & 7]));
                acc[i] = lo;
            }
        }
    }
    const long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s += acc[i] + a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0)
        cyc[blockIdx.x] = t1 - t0;
}

static double avg_cycles(long long* d, int n) {
    static long long h[4096];
    cudaMemcpy(h, d, sizeof(long long) * n, cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n; ++i)
        s += (double)h[i];
    return s / n;
}

template <int V, int MINB = 3>
void run_b(const char* name, uint64_t* data, Twiddle* tw, ModConst mc, long long* cyc, int grid, int per_sm = 3) {
    grid = 148 * per_sm;
    kb<V, MINB><<<grid, 256>>>(data, tw, mc, cyc);
    kb<V, MINB><<<grid, 256>>>(data, tw, mc, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s: %s\n", name, cudaGetErrorString(e));
        return;
    }
    // per SMSP: 3 CTAs x 8 warps / 4 = 6 warps, each kIters * 32 butterflies
    const double c = avg_cycles(cyc, grid) / ((double)kIters * 32 * 2 * per_sm);
    printf("[regcap %3d, %d CTAs/SM] %-52s cycles/warp-butterfly/SMSP = %6.2f   => floor %.3f us per 2^16 limb-NTT at 1.965 GHz\n", name, 65536 / (256 * MINB) > 255 ? 255 : 65536 / (256 * MINB), per_sm, name, c,
           16384.0 * c / (148 * 4) / 1965.0);
}
template <int M>
void run_r(const char* name, uint64_t* data, long long* cyc, int grid) {
    kraw<M><<<grid, 256>>>(data, cyc);
    kraw<M><<<grid, 256>>>(data, cyc);
    cudaDeviceSynchronize();
    const double c = avg_cycles(cyc, grid) / (4096.0 * 8 * 6);
    printf("%-52s clk/warp-instr/SMSP = %5.2f\n", name, c);
}

int main() {
    const int grid = 148 * 8;
    uint64_t* data;
    Twiddle* tw;
    long long* cyc;
    cudaMalloc(&data, sizeof(uint64_t) * 4096 * grid);
    cudaMemset(data, 1, sizeof(uint64_t) * 4096 * grid);
    cudaMalloc(&tw, sizeof(Twiddle) * 64);
    cudaMemset(tw, 3, sizeof(Twiddle) * 64);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    ModConst mc = {};
    mc.q        = 576460752303415297ull;  // a 59-bit NTT prime
    run_r<0>("IMAD.WIDE.U32 acc64, distinct operands", data, cyc, grid);
    run_r<1>("IMAD.WIDE.U32 RZ addend, dependent chain x8", data, cyc, grid);
    run_r<2>("IMAD.HI.U32, dependent chain x8", data, cyc, grid);
    run_r<3>("IMAD 32-bit acc, distinct operands", data, cyc, grid);
    run_b<0>("V0 fwd round-1 (exact quotient, csub every stage)", data, tw, mc, cyc, grid);
    run_b<1>("V1 fwd round-2 shipped (truncated, csub alt. stages)", data, tw, mc, cyc, grid);
    run_b<2>("V2 fwd round-2, product in PTX", data, tw, mc, cyc, grid);
    run_b<3>("V3 fwd round-2, cross products via mul.hi", data, tw, mc, cyc, grid);
    run_b<4>("V4 fwd truncated, csub every stage", data, tw, mc, cyc, grid);
    run_b<5>("V5 inv round-2 shipped", data, tw, mc, cyc, grid);
    // occupancy / register-cap sweep of the shipped forward and inverse butterflies
    run_b<2, 1>("V2 fwd", data, tw, mc, cyc, grid, 1);
    run_b<2, 2>("V2 fwd", data, tw, mc, cyc, grid, 1);
    run_b<2, 2>("V2 fwd", data, tw, mc, cyc, grid, 2);
    run_b<2, 3>("V2 fwd", data, tw, mc, cyc, grid, 1);
    run_b<2, 3>("V2 fwd", data, tw, mc, cyc, grid, 2);
    run_b<2, 3>("V2 fwd", data, tw, mc, cyc, grid, 3);
    run_b<2, 4>("V2 fwd", data, tw, mc, cyc, grid, 3);
    run_b<2, 4>("V2 fwd", data, tw, mc, cyc, grid, 4);
    run_b<2, 5>("V2 fwd", data, tw, mc, cyc, grid, 5);
    run_b<2, 6>("V2 fwd", data, tw, mc, cyc, grid, 6);
    run_b<5, 2>("V5 inv", data, tw, mc, cyc, grid, 2);
    run_b<5, 4>("V5 inv", data, tw, mc, cyc, grid, 4);
    run_b<5, 6>("V5 inv", data, tw, mc, cyc, grid, 6);
    run_b<6>("V6 inv round-1", data, tw, mc, cyc, grid);
    return 0;
}
