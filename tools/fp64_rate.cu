// FP64 issue rates on this GPU: DFMA on the vector pipe against DMMA (mma.sync.m8n8k4.f64) on the tensor cores.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    double* out;
    cudaMalloc(&out, (size_t)sms * 4 * 512 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int threads : {128, 256, 512}) {
        for (int which = 0; which < 2; ++which) {
            const int iters = 20000;
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_dfma<<<sms * 2, threads>>>(out, iters, 1.0000001, 1e-9);
                else k_dmma<<<sms * 2, threads>>>(out, iters, 1.0000001, 1e-9);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double fl = which == 0 ? (double)sms * 2 * threads * iters * 16 * 2
                                         : (double)sms * 2 * (threads / 32) * iters * 8 * (8 * 8 * 4 * 2);
            printf("%s threads/CTA %d (2 CTAs/SM): %.3f ms, %.2f TFLOP/s\n", which == 0 ? "DFMA" : "DMMA m8n8k4", threads, best,
                   fl / best * 1e-9);
        }
    }
    printf("%s, %d SMs\n", pr.name, sms);
    return 0;
}
