// Probe of the tcgen05 pieces the tensor-core sweep relies on (run on a B200):
//   one CTA, D[128 x 16] = A[128 x K] . B[16 x K]^T with kind::tf32, both operands K-major in shared memory
//   without swizzle, accumulator in TMEM, read back with tcgen05.ld.32x32b.
// It answers, on the hardware, which of the two (LBO, SBO) readings of the no-swizzle K-major descriptor is the
// right one, that the instruction descriptor encodes M=128 / N=16 / tf32 / fp32 accumulate as intended, that a
// second MMA with enable_input_d accumulates, and how the tensor core treats the low 13 mantissa bits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tc_probe tools/tc_probe.cu && tools/tc_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (sm_100)
    return d;                    // layout type 0 (no swizzle), base offset 0
}

struct ProbeParams {
    const float* A;   // [128][16]  row-major (row, k)
    const float* B;   // [16][16]   row-major (n, k)
    float* D;         // [128][16]
    uint32_t lbo, sbo;
    int ksteps;       // 1 or 2 (K = 8 or 16)
};

// shared layout of one k-step of an operand with R rows: [row group][k chunk (2)][row in group (8)][4 floats]
__device__ __forceinline__ uint32_t op_off(int row, int k) {   // k in 0..7
    return (uint32_t)((row >> 3) * 256 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

__global__ void __launch_bounds__(128, 1) k_probe(ProbeParams P) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* sA = sm;                 // 2 k-steps x 4096
    unsigned char* sB = sm + 8192;          // 2 k-steps x 512
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192 + 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + 8192 + 1024 + 16);
    const int tid = threadIdx.x;
    for (int s = 0; s < 2; ++s)
        for (int k = 0; k < 8; ++k) {
            *reinterpret_cast<float*>(sA + s * 4096 + op_off(tid, k)) = P.A[tid * 16 + s * 8 + k];
            if (tid < 16) *reinterpret_cast<float*>(sB + s * 512 + op_off(tid, k)) = P.B[tid * 16 + s * 8 + k];
        }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        // instruction descriptor: fp32 accumulate, tf32 x tf32, K-major both, N = 16, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        for (int s = 0; s < P.ksteps; ++s) {
            uint64_t da = make_desc(smem_u32(sA + s * 4096), P.lbo, P.sbo);
            uint64_t db = make_desc(smem_u32(sB + s * 512), P.lbo, P.sbo);
            uint32_t acc = s > 0 ? 1u : 0u;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // wait for the MMAs
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(0)
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) P.D[tid * 16 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    std::vector<float> A(128 * 16), B(16 * 16), D(128 * 16);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX + 0.5f;
    for (auto& v : B) v = (float)rand() / RAND_MAX - 0.3f;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    const uint32_t combos[2][2] = {{128, 256}, {256, 128}};   // (LBO, SBO)
    for (int c = 0; c < 2; ++c)
        for (int ks = 1; ks <= 2; ++ks) {
            ProbeParams P = {dA, dB, dD, combos[c][0], combos[c][1], ks};
            CK(cudaMemset(dD, 0xFF, D.size() * 4));
            k_probe<<<1, 128, 16384>>>(P);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double e_exact = 0, e_trunc = 0, e_rna = 0;
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 16; ++j) {
                    double r0 = 0, r1 = 0, r2 = 0;
                    for (int k = 0; k < 8 * ks; ++k) {
                        r0 += (double)A[i * 16 + k] * B[j * 16 + k];
                        r1 += (double)tf32_trunc(A[i * 16 + k]) * tf32_trunc(B[j * 16 + k]);
                        r2 += (double)tf32_rna(A[i * 16 + k]) * tf32_rna(B[j * 16 + k]);
                    }
                    double d = D[i * 16 + j];
                    e_exact = fmax(e_exact, fabs(d - r0));
                    e_trunc = fmax(e_trunc, fabs(d - r1));
                    e_rna = fmax(e_rna, fabs(d - r2));
                }
            printf("LBO=%u SBO=%u ksteps=%d: max|D - exact|=%.3e  |D - trunc-tf32|=%.3e  |D - rna-tf32|=%.3e  D[0][0]=%g D[5][3]=%g\n",
                   combos[c][0], combos[c][1], ks, e_exact, e_trunc, e_rna, D[0], D[5 * 16 + 3]);
        }
    // accumulation precision: operands already tf32-exact, compare with float64
    for (auto& v : A) v = tf32_rna(v);
    for (auto& v : B) v = tf32_rna(v);
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int c = 0; c < 2; ++c) {
        ProbeParams P = {dA, dB, dD, combos[c][0], combos[c][1], 2};
        k_probe<<<1, 128, 16384>>>(P);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double emax = 0, smax = 0;
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 16; ++j) {
                double r = 0, sa = 0;
                for (int k = 0; k < 16; ++k) { r += (double)A[i * 16 + k] * B[j * 16 + k]; sa += fabs((double)A[i * 16 + k] * B[j * 16 + k]); }
                emax = fmax(emax, fabs(D[i * 16 + j] - r) / sa);
                smax = fmax(smax, sa);
            }
        printf("tf32-exact operands, LBO=%u SBO=%u: max |D - float64| / sum|terms| = %.3e\n", combos[c][0], combos[c][1], emax);
    }
    printf("probe done\n");
    return 0;
}
