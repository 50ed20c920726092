// Microbenchmark: FFMA vs FFMA2 (fma.rn.f32x2) vs MUFU issue rates on this device.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/peak_ffma2.cu -o tools/peak_ffma2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    if (MODE == 0) {
        float a[8], x = 1.0f + 1e-7f * threadIdx.x, y = 1e-9f;
        for (int i = 0; i < 8; ++i) a[i] = 0.1f * i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 64; ++u) a[u & 7] = __fmaf_rn(a[u & 7], x, y);
        float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 123.f) out[0] = s;
    } else if (MODE == 1) {
        u64 a[8], x = 0x3f8000013f800001ull, y = 0x3089705f3089705full;
        for (int i = 0; i < 8; ++i) a[i] = 0x3dcccccd3dcccccdull + i + threadIdx.x;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 64; ++u) a[u & 7] = fma2(a[u & 7], x, y);
        u64 s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
        if (s == 123) out[0] = 1.f;
    } else if (MODE == 3 || MODE == 4) {
        u64 a[8], x = 0x3f8000013f800001ull;
        for (int i = 0; i < 8; ++i) a[i] = 0x3dcccccd3dcccccdull + i + threadIdx.x;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 64; ++u) {
                if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[u & 7]) : "l"(x));
                else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[u & 7]) : "l"(x));
            }
        u64 s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
        if (s == 123) out[0] = 1.f;
    } else if (MODE == 5 || MODE == 6) {
        // kernel-like mix: 6 FFMA2 : 1 MUFU (: 2 FSEL)
        u64 a[6], x = 0x3f8000013f800001ull, y = 0x3089705f3089705full;
        float m[4], sel[4];
        for (int i = 0; i < 6; ++i) a[i] = 0x3dcccccd3dcccccdull + i + threadIdx.x;
        for (int i = 0; i < 4; ++i) { m[i] = -0.001f * (i + 1); sel[i] = 0.5f * i; }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 16; ++u) {
#pragma unroll
                for (int q = 0; q < 6; ++q) a[q] = fma2(a[q], x, y);
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[u & 3]));
                if (MODE == 6) {
                    asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0f00000000; selp.f32 %0, %0, %2, p;}" : "+f"(sel[u & 3]) : "f"(m[(u + 1) & 3]), "f"(m[(u + 2) & 3]));
                    asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0f3f000000; selp.f32 %0, %0, %2, p;}" : "+f"(sel[(u + 1) & 3]) : "f"(m[(u + 3) & 3]), "f"(m[(u + 2) & 3]));
                }
            }
        u64 s = 0; for (int i = 0; i < 6; ++i) s ^= a[i];
        for (int i = 0; i < 4; ++i) s += (u64)(m[i] + sel[i]);
        if (s == 123) out[0] = 1.f;
    } else {
        // mixed: 1 FFMA2 + 1 MUFU + 1 FSEL-ish (ALU) per slot group
        u64 a[4], x = 0x3f8000013f800001ull, y = 0x3089705f3089705full;
        float m[4];
        for (int i = 0; i < 4; ++i) { a[i] = 0x3dcccccd3dcccccdull + i + threadIdx.x; m[i] = -0.001f * (i + 1); }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                a[u & 3] = fma2(a[u & 3], x, y);
                a[(u + 1) & 3] = fma2(a[(u + 1) & 3], x, y);
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[u & 3]));
            }
        u64 s = 0; for (int i = 0; i < 4; ++i) s ^= a[i] + (u64)m[i];
        if (s == 123) out[0] = 1.f;
    }
}
int main() {
    float* out; cudaMalloc(&out, 256);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 8, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 7; ++mode) {
        float best = 1e9;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, 256>>>(out, iters);
            else if (mode == 1) k<1><<<blocks, 256>>>(out, iters);
            else if (mode == 2) k<2><<<blocks, 256>>>(out, iters);
            else if (mode == 3) k<3><<<blocks, 256>>>(out, iters);
            else if (mode == 4) k<4><<<blocks, 256>>>(out, iters);
            else if (mode == 5) k<5><<<blocks, 256>>>(out, iters);
            else k<6><<<blocks, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        double inst = (double)blocks * 256 * iters * (mode == 2 ? 96 : mode == 5 ? 112 : mode == 6 ? 176 : 64);
        double lanes_fma = (double)blocks * 256 * iters * (mode == 0 ? 64 : mode >= 5 ? 192 : 128);
        printf("mode %d: %.3f ms  thread-instr/s %.3e  warp-instr/clk/SM (at 1.965GHz) %.3f  FMA TFLOP/s %.2f\n", mode, best,
               inst / (best * 1e-3), inst / 32 / (best * 1e-3) / p.multiProcessorCount / 1.965e9, 2 * lanes_fma / (best * 1e-3) / 1e12);
    }
    return 0;
}
