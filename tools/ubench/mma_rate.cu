// mma_rate.cu -- issue rate of tcgen05.mma (cta_group::1, M = 128, N = 256, operands in shared memory, K-major,
// 128-byte swizzle) by kind on this GPU: cycles per instruction and MACs per clock per SM.  One CTA per SM, thread 0
// issues `reps` x 4 K-steps alternating between two accumulator buffers, one commit at the end.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else if constexpr (KIND == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else if constexpr (KIND == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) rate_kernel(uint32_t idesc, int reps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += 128)
        reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base, a = smem_u32(smem), b = a + 16 * 1024;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k)
                mma<KIND>(tmem + (r & 1) * 256, smem_desc(a + 32 * k), smem_desc(b + 32 * k), idesc, k);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        for (int spin = 0; !done && spin < 4000; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0)
            *cycles = done ? t1 - t0 : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int KIND>
void run(const char* name, uint32_t idesc, int k_elems) {
    long long* d;
    cudaMalloc(&d, 8);
    const int reps = 256, smem = 49 * 1024 + 1024;
    cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 2; ++it)
        rate_kernel<KIND><<<148, 128, smem>>>(idesc, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c   = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (reps * 4);
    printf("%-28s M128 N256 K%-3d  %8.1f clk/instr  %8.0f MAC/clk/SM  (%s)\n", name, k_elems, per, 128.0 * 256 * k_elems / per,
           e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    const uint32_t mn = ((256u >> 3) << 17) | ((128u >> 4) << 24);
    run<0>("kind::i8 u8 x u8 -> s32", (2u << 4) | mn, 32);
    run<0>("kind::i8 s8 x s8 -> s32", (2u << 4) | (1u << 7) | (1u << 10) | mn, 32);
    run<1>("kind::f8f6f4 e4m3 -> f32", (1u << 4) | mn, 32);
    run<2>("kind::f16 bf16 -> f32", (1u << 4) | (1u << 7) | (1u << 10) | mn, 16);
    run<2>("kind::f16 f16 -> f32", (1u << 4) | mn, 16);
    run<3>("kind::tf32 -> f32", (1u << 4) | (2u << 7) | (2u << 10) | mn, 8);
    return 0;
}
