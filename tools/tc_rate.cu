// How fast does one thread issue small tcgen05.mma (kind::tf32, M=128, N in {16,32,64}, K=8, operands in shared
// memory, no swizzle)?  Prints cycles per MMA for back-to-back issue with one commit + wait per group of 20,
// (a) all into distinct accumulators, (b) in accumulate chains of 2 as the sweep does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tc_rate tools/tc_rate.cu && tools/tc_rate
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(uint32_t d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nmov.b64 da, {%1, %3};\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n}\n" ::"r"(d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k_rate(int N, int iters, int chain, int adist, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 160 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 160 * 1024 + 16);
    for (int i = threadIdx.x; i < 40 * 1024; i += 128) reinterpret_cast<float*>(sm)[i] = 1.0f;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t hi = (256u >> 4) | (1u << 14);
        const uint32_t alo = ((smem_u32(sm) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
        const uint32_t blo = ((smem_u32(sm + 96 * 1024) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int g = 0; g < 20; ++g) {
                uint32_t col = chain ? (uint32_t)((g / 2) * N) % 512 : (uint32_t)(g * N) % 512;
                if (col + N > 512) col = 0;
                mma(tmem + col, alo + (uint32_t)((g % adist) * 4096 >> 4), blo + (uint32_t)(g * 512 >> 4), hi, idesc, chain ? (g & 1) : 0);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
            asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(it & 1) : "memory");
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
    long long* d; CK(cudaMalloc(&d, 8));
    CK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024));
    for (int N : {16, 32, 64, 128})
        for (int chain = 0; chain < 2; ++chain)
            for (int adist : {1, 20}) {
                k_rate<<<148, 128, 161 * 1024>>>(N, 2000, chain, adist, d);
                CK(cudaDeviceSynchronize());
                long long c; CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
                printf("N=%3d chain=%d distinctA=%2d: %.1f cycles per MMA (group of 20 + commit + wait: %.0f)\n", N, chain, adist, (double)c / (2000.0 * 20), (double)c / 2000.0);
            }
    return 0;
}
