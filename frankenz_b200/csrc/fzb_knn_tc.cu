// Tensor-core candidate scan for the brute-force kNN search (frankenz/knn.py:362-365, the K cKDTree queries).
//
// For p = 2 the squared distance is |f - q|^2 = |q'|^2 + v,  v = |f'|^2 - 2 f'.q',  with f' = f - c, q' = q - c and c
// the mean training feature (distances do not depend on c; centring keeps the numbers small).  v is a GEMM with
// K = Nf (north_star: "a |x|^2 - 2x.y + |y|^2 contraction with padded K = filters"), so the 5th-generation tensor cores
// evaluate it:  D[128 queries][256 rows] = A[queries][K] . B[rows][K]^T, kind::tf32, operands K-major in shared memory,
// accumulators in TMEM (2 x 256 columns, double buffered), one elected thread issuing.  tf32 keeps 11 significant bits,
// so both sides are split x = hi + lo and a band takes three K-slots (hi*hi + hi*lo + lo*hi); |f'|^2 is formed in
// float64 and enters through three more slots (three tf32 pieces against 1.0).  K = 3 Nf + 3 -> 16 / 24 / 24 for 4 / 5 / 6
// bands, i.e. 2 or 3 K = 8 instructions per (128 x 256) block of distance evaluations.
//
// The scan only GENERATES CANDIDATES: every consumer thread owns one query (= one TMEM lane), reads its 256 values of a
// block with tcgen05.ld, reduces 32 of them at a time with 3-input minima and only looks at individual values when the
// minimum beats the threshold of its candidate list (the KC = k + 8 smallest v seen so far, kept in shared memory).
// The float64 re-rank of fzb_knn.cu then orders the candidates by their exact (numpy-bit) distances and accepts the
// result only if the k-th exact distance lies below the smallest distance an EXCLUDED row can have,
//     d_k^2 < tau_v + |q'|^2 - delta,   delta = c0 (|q'|^2 + max_j |f'_j|^2),
// where tau_v is the list threshold and delta bounds the error of v (operand splitting 3 x 2^-21 |f'||q'|, fp32
// accumulation of the tensor core); c0 = 4e-6, four times the bound of the splitting alone, and the re-rank records
// the largest error it actually sees on the candidates (FzbStats.knn_tc_err) so that tests and bench.py can assert the
// margin.  (query, tree) pairs that fail the test are searched again by the all-float64 kernel (k_knn_exact), so the
// neighbour lists stay exact whatever the tensor cores do.
//
// CTA = 10 warps: warps 0-7 consumers (128 queries; warps w and w + 4 read the same 32 TMEM lanes and take half of the
// 256 columns each, with a list of their own: they report like two row splits), warp 8 TMA producer (row tiles of 256
// rows, built once per fzb_knn_build), warp 9 MMA issuer.
#include <type_traits>

#include <math_constants.h>

#include "fzb_sweep_common.cuh"

namespace {
using namespace fzbsweep;
#include "fzb_sweep_tc.cuh"      // tcgen05 / mbarrier helpers (the sweep kernel template itself is not instantiated here)

constexpr int KT_ROWS = 256;                 // training rows per tile = N of one MMA
constexpr int KT_Q = 128;                    // queries per CTA = M
constexpr int KT_CW = 8;                     // consumer warps: warps w and w + 4 share 32 queries, half of the columns each
constexpr int KT_THREADS = (KT_CW + 2) * 32;
constexpr int KT_NSTAGE = 3;
constexpr int KT_KCMAX = 48;                 // candidate list capacity (k + 8): shared-memory budget
__host__ __device__ constexpr int kt_ksteps(int nf) { return (3 * nf + 3 + 7) / 8; }
__host__ __device__ constexpr int kt_tile_bytes(int nf) { return kt_ksteps(nf) * KT_ROWS * 32; }
__host__ __device__ constexpr int kt_a_bytes(int nf) { return kt_ksteps(nf) * KT_Q * 32; }
__host__ __device__ constexpr size_t kt_smem(int nf, int kc) {
    return (size_t)kt_a_bytes(nf) + (size_t)KT_NSTAGE * kt_tile_bytes(nf) + (size_t)kc * 2 * KT_Q * 8 + 32 * 2 * KT_Q * 4 + 256;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- tile builder: one thread per (tree, row) ----------------------------------------------------------------------------
struct KtBuildParams {
    const float* feats;      // [K][stride] float32 features
    int64_t stride, Nm;
    int K, Nf;
    const double* centre;    // [Nf]
    unsigned char* tiles;    // [K][ntile][tile_bytes]
    int64_t ntile;
    double* fmax2;           // [K]: max |f'|^2 per tree (atomic max on the bit pattern: values are >= 0)
};

__global__ void k_knn_build_tiles(KtBuildParams P) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t per_tree = P.ntile * KT_ROWS;
    if (g >= (int64_t)P.K * per_tree) return;
    const int t = (int)(g / per_tree);
    const int64_t row = g - (int64_t)t * per_tree;
    const int nf = P.Nf, ks = kt_ksteps(nf);
    float v[32];
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    if (row < P.Nm) {
        const float* f = P.feats + (size_t)t * P.stride + (size_t)row * nf;
        double n2 = 0.0;
        for (int b = 0; b < nf; ++b) {
            const double fp = (double)f[b] - P.centre[b];
            n2 += fp * fp;
            const float x = (float)fp;
            const float hi = tf32_rn(x), lo = tf32_rn((float)(fp - (double)hi));
            v[3 * b] = hi; v[3 * b + 1] = lo; v[3 * b + 2] = hi;          // against (q_hi, q_hi, q_lo)
        }
        const float n0 = tf32_rn((float)n2);
        const float n1 = tf32_rn((float)(n2 - (double)n0));
        const float n2c = tf32_rn((float)(n2 - (double)n0 - (double)n1));
        v[3 * nf] = n0; v[3 * nf + 1] = n1; v[3 * nf + 2] = n2c;
        if (isfinite(n2)) atomicMax(reinterpret_cast<unsigned long long*>(P.fmax2 + t), (unsigned long long)__double_as_longlong(n2));
        else v[3 * nf] = 1e30f;           // non-finite features never become candidates (the float64 kernel handles them)
    } else {
        v[3 * nf] = 1e30f;                // padding rows of the last tile
    }
    unsigned char* T = P.tiles + ((size_t)t * P.ntile + (size_t)(row / KT_ROWS)) * kt_tile_bytes(nf);
    const int r = (int)(row % KT_ROWS);
    unsigned char* dst = T + (r >> 3) * 256 + (r & 7) * 16;
    for (int s = 0; s < ks; ++s)
        for (int c = 0; c < 2; ++c)
            *reinterpret_cast<float4*>(dst + s * (KT_ROWS * 32) + c * 128) =
                make_float4(v[s * 8 + c * 4], v[s * 8 + c * 4 + 1], v[s * 8 + c * 4 + 2], v[s * 8 + c * 4 + 3]);
}

__global__ void k_knn_centre(const float* feats, int64_t stride, int64_t Nm, int K, int Nf, double* sums) {
    // sums[b] += sum over rows of feature b (all trees); grid-stride, one atomic per thread and band
    double acc[FZB_FAST_MAXF];
    for (int b = 0; b < Nf; ++b) acc[b] = 0.0;
    const int64_t total = (int64_t)K * Nm;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(g / Nm);
        const float* f = feats + (size_t)t * stride + (size_t)(g - (int64_t)t * Nm) * Nf;
        for (int b = 0; b < Nf; ++b) {
            const float x = f[b];
            if (isfinite(x)) acc[b] += (double)x;
        }
    }
    for (int b = 0; b < Nf; ++b) atomicAdd(sums + b, acc[b]);
}
__global__ void k_knn_centre_finish(double* sums, int Nf, double inv_n) {
    if (threadIdx.x < Nf) sums[threadIdx.x] *= inv_n;
}

// ---- the scan ----------------------------------------------------------------------------------------------------------
struct KtScanParams {
    const unsigned char* tiles;     // [K][ntile][tile_bytes]
    int64_t ntile, Nm;
    const double* q;                // [No][Nf] float64 queries
    const double* centre;           // [Nf]
    int64_t No;
    int KC;
    int tiles_per_split;
    float* cand_d;                  // [(query * K + tree) * nsp + split][KC]   approximate v (not yet + |q'|^2)
    int* cand_i;
};

template <int NF>
__global__ void __launch_bounds__(KT_THREADS, 1) k_knn_scan_tc(KtScanParams P) {
    constexpr int KS = kt_ksteps(NF), TILE_BYTES = kt_tile_bytes(NF), A_BYTES = kt_a_bytes(NF);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* objA = smem_raw;
    unsigned char* stage = smem_raw + A_BYTES;
    constexpr int NT = 2 * KT_Q;        // consumer threads: two per query (column halves)
    float* list_d = reinterpret_cast<float*>(stage + (size_t)KT_NSTAGE * TILE_BYTES);        // [KC][256]
    int* list_i = reinterpret_cast<int*>(list_d + (size_t)P.KC * NT);
    float* stash = reinterpret_cast<float*>(list_i + (size_t)P.KC * NT);                    // [32][256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stash + 32 * NT);
    uint64_t* tile_full = bars;                    // [NSTAGE] TMA -> MMA
    uint64_t* tile_empty = bars + KT_NSTAGE;       // [NSTAGE] MMA commit -> TMA
    uint64_t* acc_full = bars + 2 * KT_NSTAGE;     // [2] MMA commit -> consumers
    uint64_t* acc_empty = bars + 2 * KT_NSTAGE + 2;   // [2] consumer warps -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * KT_NSTAGE + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = blockIdx.y, sp = blockIdx.z, nsp = gridDim.z, K = gridDim.y;
    const int64_t t0 = (int64_t)sp * P.tiles_per_split;
    int64_t t1 = t0 + P.tiles_per_split;
    if (t1 > P.ntile) t1 = P.ntile;
    const int nt = (int)(t1 - t0);
    const unsigned char* tiles = P.tiles + ((size_t)t * P.ntile + (size_t)t0) * TILE_BYTES;

    if (tid == 0) {
        for (int s = 0; s < KT_NSTAGE; ++s) { mbar_init(&tile_full[s], 1); mbar_init(&tile_empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], KT_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == KT_CW + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const int qrow = tid & (KT_Q - 1);                        // consumer threads: query row (TMEM lane) ...
    const int chalf = tid >> 7;                               // ... and column half
    const int64_t oq = (int64_t)blockIdx.x * KT_Q + qrow;
    if (warp < KT_CW)
        for (int c = 0; c < P.KC; ++c) { list_d[c * NT + tid] = CUDART_INF_F; list_i[c * NT + tid] = -1; }
    if (warp < 4) {
        // A operand row of this query: (-2 q'_hi, -2 q'_hi, -2 q'_lo) per band, then (1, 1, 1) for |f'|^2
        float arow[8 * KS];
#pragma unroll
        for (int i = 0; i < 8 * KS; ++i) arow[i] = 0.f;
        const int64_t oo = oq < P.No ? oq : P.No - 1;
#pragma unroll
        for (int b = 0; b < NF; ++b) {
            const double qp = P.q[oo * NF + b] - P.centre[b];
            const float hi = tf32_rn((float)qp), lo = tf32_rn((float)(qp - (double)hi));
            arow[3 * b] = -2.f * hi; arow[3 * b + 1] = -2.f * hi; arow[3 * b + 2] = -2.f * lo;
        }
        arow[3 * NF] = 1.f; arow[3 * NF + 1] = 1.f; arow[3 * NF + 2] = 1.f;
        unsigned char* dst = objA + (tid >> 3) * 256 + (tid & 7) * 16;
#pragma unroll
        for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                *reinterpret_cast<float4*>(dst + s * (KT_Q * 32) + c * 128) =
                    make_float4(arow[s * 8 + c * 4], arow[s * 8 + c * 4 + 1], arow[s * 8 + c * 4 + 2], arow[s * 8 + c * 4 + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == KT_CW) {
        // ===== TMA producer =======================================================================================
        if (lane == 0) {
            for (int it = 0; it < nt; ++it) {
                const int st = it % KT_NSTAGE, n = it / KT_NSTAGE;
                if (n > 0) mbar_wait_hint(&tile_empty[st], (uint32_t)((n - 1) & 1));
                mbar_expect_tx(&tile_full[st], TILE_BYTES);
                bulk_g2s(stage + (size_t)st * TILE_BYTES, tiles + (size_t)it * TILE_BYTES, TILE_BYTES, &tile_full[st]);
            }
        }
    } else if (warp == KT_CW + 1) {
        // ===== MMA issuer ===========================================================================================
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KT_ROWS >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t d0 = tc_desc(smem_u32(objA), 128u, 256u);
        const uint32_t a_lo = (uint32_t)d0, desc_hi = (uint32_t)(d0 >> 32);
        for (int it = 0; it < nt; ++it) {
            const int st = it % KT_NSTAGE, n = it / KT_NSTAGE;
            const uint32_t buf = it & 1, use = it >> 1;
            mbar_wait_hint(&tile_full[st], (uint32_t)(n & 1));
            mbar_wait_hint(&acc_empty[buf], (use & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t b_lo = (uint32_t)tc_desc(smem_u32(stage + (size_t)st * TILE_BYTES), 128u, 256u);
                const uint32_t dcol = tmem + buf * KT_ROWS;
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const uint32_t ak = a_lo + (k * (KT_Q * 32) >> 4), bk = b_lo + (k * (KT_ROWS * 32) >> 4);
                    if (k == 0) tc_mma<false>(dcol, ak, bk, desc_hi, idesc);
                    else tc_mma<true>(dcol, ak, bk, desc_hi, idesc);
                }
                tc_commit(&acc_full[buf]);
                tc_commit(&tile_empty[st]);
            }
            __syncwarp();
        }
    } else {
        // ===== consumers: thread <-> query (TMEM lane) ==============================================================
        // The candidate list is a binary max-heap in shared memory (entry c of thread t at [c][t], root = entry 1 = the
        // threshold tau): a row that beats tau replaces the root and sifts down, ~2 log2(KC) shared-memory reads.
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + chalf * (KT_ROWS / 2);
        float tau = CUDART_INF_F;
        const int KC = P.KC;
        float* hd = list_d - NT;            // 1-based
        int* hi = list_i - NT;
        auto insert = [&](float d, int idx) {
            int pos = 1;
            while (true) {
                const int l = 2 * pos;
                if (l > KC) break;
                const float dl = hd[l * NT + tid];
                const float dr = (l + 1 <= KC) ? hd[(l + 1) * NT + tid] : -CUDART_INF_F;
                const int c = (dr > dl) ? l + 1 : l;
                const float dc = fmaxf(dl, dr);
                if (!(dc > d)) break;
                hd[pos * NT + tid] = dc;
                hi[pos * NT + tid] = hi[c * NT + tid];
                pos = c;
            }
            hd[pos * NT + tid] = d;
            hi[pos * NT + tid] = idx;
            tau = hd[NT + tid];
        };
        for (int it = 0; it < nt; ++it) {
            const uint32_t buf = it & 1, use = it >> 1;
            mbar_wait_hint(&acc_full[buf], use & 1);
            tc_fence_after();
            const int rowbase = (int)((t0 + it) * KT_ROWS);
#pragma unroll 1
            for (int c = 0; c < KT_ROWS / 64; ++c) {
                float v[32];
                tmem_ld32(lane_addr + buf * KT_ROWS + c * 32, v);
                if (c == KT_ROWS / 64 - 1) {     // everything this warp needs of the block is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                float m4[4];                     // four independent chains of 3-input minima
#pragma unroll
                for (int j = 0; j < 4; ++j) m4[j] = fminf(v[8 * j], v[8 * j + 1]);
#pragma unroll
                for (int i = 2; i < 8; i += 2)
#pragma unroll
                    for (int j = 0; j < 4; ++j) m4[j] = fminf(m4[j], fminf(v[8 * j + i], v[8 * j + i + 1]));
                const float m = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
                if (m < tau) {                   // rare after the first few thousand rows
                    // park the 32 values in shared memory and walk them with one compact loop (an unrolled chain of 32
                    // inlined inserts thrashes the instruction cache)
#pragma unroll
                    for (int i = 0; i < 32; ++i) stash[i * NT + tid] = v[i];
                    const int r0 = rowbase + chalf * (KT_ROWS / 2) + c * 32;
#pragma unroll 1
                    for (int i = 0; i < 32; ++i) {
                        const float x = stash[i * NT + tid];
                        if (x < tau && r0 + i < (int)P.Nm) insert(x, r0 + i);
                    }
                }
            }
        }
        if (oq < P.No) {      // the two column halves report like two row splits
            const size_t base = (((size_t)oq * K + t) * (2 * nsp) + 2 * sp + chalf) * KC;
            for (int c = 0; c < KC; ++c) { P.cand_d[base + c] = list_d[c * NT + tid]; P.cand_i[base + c] = list_i[c * NT + tid]; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == KT_CW + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int NF>
int launch_scan_tc(fzb_context* h, const KtScanParams& P, dim3 grid) {
    const size_t smem = kt_smem(NF, P.KC);
    FZB_CUDA(cudaFuncSetAttribute(k_knn_scan_tc<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_knn_scan_tc<NF><<<grid, KT_THREADS, smem, h->stream>>>(P);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

// Build the centred, split row tiles of every tree (called by fzb_knn_build; Nf in 4..6, Nm < 2^31)
int fzb_knn_tc_build(fzb_context* h) {
    const int nf = h->knn_Nf, K = h->knn_K;
    const int64_t Nm = h->knn_Nm;
    h->knn_tc_valid = false;
    if (nf < 4 || nf > 6 || Nm >= ((int64_t)1 << 31) - KT_ROWS || Nm < 4096) return 0;
    const int64_t ntile = (Nm + KT_ROWS - 1) / KT_ROWS;
    const size_t bytes = (size_t)K * ntile * kt_tile_bytes(nf);
    if (h->knn_tiles.reserve(bytes + 256) || h->knn_aux.reserve((size_t)(FZB_FAST_MAXF + K + 8) * 8)) return 1;
    double* centre = h->knn_aux.as<double>();
    double* fmax2 = centre + FZB_FAST_MAXF;
    FZB_CUDA(cudaMemsetAsync(h->knn_aux.p, 0, (size_t)(FZB_FAST_MAXF + K + 8) * 8, h->stream));
    k_knn_centre<<<h->sm_count * 4, 256, 0, h->stream>>>(h->knn_feats.as<float>(), h->knn_stride, Nm, K, nf, centre);
    k_knn_centre_finish<<<1, 32, 0, h->stream>>>(centre, nf, 1.0 / ((double)K * (double)Nm));
    KtBuildParams B = {};
    B.feats = h->knn_feats.as<float>(); B.stride = h->knn_stride; B.Nm = Nm; B.K = K; B.Nf = nf; B.centre = centre;
    B.tiles = h->knn_tiles.as<unsigned char>(); B.ntile = ntile; B.fmax2 = fmax2;
    const int64_t total = (int64_t)K * ntile * KT_ROWS;
    k_knn_build_tiles<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(B);
    fzb_count_launch(h, 3);
    FZB_CUDA(cudaGetLastError());
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->knn_ntile = ntile;
    h->knn_tc_valid = true;
    return 0;
}

int fzb_knn_tc_kcmax() { return KT_KCMAX; }
int fzb_knn_tc_lists() { return 2; }     // candidate lists per (query, tree, row split)

// candidate scan of `No` queries against every tree; cand buffers laid out like those of the fp32 scan
int fzb_knn_tc_scan(fzb_context* h, const double* d_q, int64_t No, int KC, int nsp, int tiles_per_split, float* cand_d,
                    int* cand_i) {
    KtScanParams P = {};
    P.tiles = h->knn_tiles.as<unsigned char>(); P.ntile = h->knn_ntile; P.Nm = h->knn_Nm; P.q = d_q;
    P.centre = h->knn_aux.as<double>(); P.No = No; P.KC = KC; P.tiles_per_split = tiles_per_split;
    P.cand_d = cand_d; P.cand_i = cand_i;
    dim3 grid((unsigned)((No + KT_Q - 1) / KT_Q), (unsigned)h->knn_K, (unsigned)nsp);
    switch (h->knn_Nf) {
        case 4: return launch_scan_tc<4>(h, P, grid);
        case 5: return launch_scan_tc<5>(h, P, grid);
        case 6: return launch_scan_tc<6>(h, P, grid);
        default: break;
    }
    fzb_set_error("tensor-core kNN scan: unsupported filter count %d", h->knn_Nf);
    return 2;
}
