// Microbenchmark: issue cost of IMAD.WIDE.U32 (64-bit accumulate), 32-bit IMAD, DFMA, IADD3 and mixes on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <int MODE>
__global__ void k(uint64_t* out, uint32_t a0, uint32_t b0, double d0) {
    uint64_t acc[8];
    double facc[8];
    uint32_t a = a0 + threadIdx.x, b = b0 + threadIdx.x;
    double fa = d0 + threadIdx.x, fb = d0 * 3.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] = i + threadIdx.x;
        facc[i] = i;
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 3 || MODE == 4)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a), "r"(b));
            if (MODE == 1) {
                uint32_t lo = (uint32_t)acc[i];
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a), "r"(b));
                acc[i] = lo;
            }
            if (MODE == 2 || MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(facc[i]) : "d"(fa), "d"(fb));
            if (MODE == 4) {
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(facc[i]) : "d"(fa), "d"(fb));
                asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(facc[i]) : "d"(fb), "d"(fa));
            }
            if (MODE == 5) {
                uint32_t lo = (uint32_t)acc[i], hi = (uint32_t)(acc[i] >> 32);
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
                acc[i] = ((uint64_t)hi << 32) | lo;
            }
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
    double fs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s += acc[i];
        fs += facc[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint64_t)fs;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (uint64_t)(t1 - t0);
}

template <int MODE>
void run(const char* name, int warps_per_smsp, int instr_per_iter) {
    uint64_t* d;
    cudaMalloc(&d, 148 * 1024 * 8);
    int threads = warps_per_smsp * 4 * 32;
    k<MODE><<<148, threads>>>(d, 3, 5, 1.5);
    k<MODE><<<148, threads>>>(d, 3, 5, 1.5);
    cudaDeviceSynchronize();
    uint64_t clk;
    cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
    double per = (double)clk / ((double)kIters * instr_per_iter * warps_per_smsp);
    printf("%-28s warps/SMSP=%d  clk/warp-instr=%.2f\n", name, warps_per_smsp, per);
    cudaFree(d);
}

int main() {
    for (int w : {1, 2, 4, 8}) {
        run<0>("IMAD.WIDE.U32 (acc64)", w, 8);
        run<1>("IMAD (32-bit)", w, 8);
        run<2>("DFMA", w, 8);
        run<3>("WIDE + DFMA (per pair)", w, 8);
        run<4>("WIDE + 2 DFMA (per triple)", w, 8);
        run<5>("IADD3 + IADD3.X (per pair)", w, 8);
    }
    return 0;
}
