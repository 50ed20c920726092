// Brute-force replacement of the K cKDTrees of frankenz/knn.py (:186 build, :362-365 query) and the
// order-preserving union of the K*k hits (pandas.unique, knn.py:368).
//
// Exactness contract: distances are float64 Minkowski-p between the float64 query and the
// float32-rounded training features promoted to float64 (cKDTree stores doubles), accumulated band by
// band without FMA contraction, i.e. the same bits numpy produces for sum((f - q)**2, axis=1).  Hits
// are ordered by (distance, index), so exact ties resolve to the lowest index.
#include <math_constants.h>

#include "fzb_common.cuh"

namespace {

constexpr int KT = 256;

__device__ __forceinline__ bool lex_less(double da, long long ia, double db, long long ib) {
    return da < db || (da == db && ia < ib);
}

// one CTA per (query, tree); each warp keeps a sorted top-k list in shared memory
__global__ void __launch_bounds__(KT) k_knn_exact(const float* __restrict__ feats, int64_t tstride, int K, int64_t Nm, int Nf,
                                                  const double* __restrict__ q, int64_t No, int k, double p, int pmode,
                                                  int64_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                                  const int64_t* __restrict__ items, const int* __restrict__ n_items) {
    extern __shared__ double sm[];
    const int nw = KT / 32;
    double* s_q = sm;                                  // Nf
    double* l_d = s_q + FZB_MAXF;                      // nw * k
    long long* l_i = reinterpret_cast<long long*>(l_d + (size_t)nw * k);   // nw * k
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* my_d = l_d + (size_t)warp * k;
    long long* my_i = l_i + (size_t)warp * k;

    const int64_t total = items ? (int64_t)*n_items : No * K;
    for (int64_t it_ = blockIdx.x; it_ < total; it_ += gridDim.x) {
        const int64_t item = items ? items[it_] : it_;
        const int64_t o = item / K;
        const int t = (int)(item % K);
        __syncthreads();
        if (tid < Nf) s_q[tid] = q[o * Nf + tid];
        __syncthreads();
        const float* F = feats + (size_t)t * tstride;
        int cnt = 0;                        // warp-uniform
        double worst_d = CUDART_INF;
        long long worst_i = 0x7fffffffffffffffll;
        for (int64_t r0 = (int64_t)warp * 32; r0 < Nm; r0 += KT) {
            int64_t r = r0 + lane;
            double d = CUDART_INF;
            bool valid = r < Nm;
            if (valid) {
                const float* f = F + r * Nf;
                double acc = 0.0;
                for (int b = 0; b < Nf; ++b) {
                    double df = __dsub_rn((double)f[b], s_q[b]);
                    if (pmode == 2) acc = __dadd_rn(acc, __dmul_rn(df, df));
                    else if (pmode == 1) acc = __dadd_rn(acc, fabs(df));
                    else if (pmode == 0) acc = fmax(acc, fabs(df));
                    else acc = __dadd_rn(acc, pow(fabs(df), p));
                }
                d = acc;
                if (isnan(d)) d = CUDART_INF;
            }
            bool want = valid && (cnt < k || lex_less(d, r, worst_d, worst_i));
            unsigned bal = __ballot_sync(0xffffffffu, want);
            while (bal) {
                int src = __ffs(bal) - 1;
                bal &= bal - 1;
                double cd = __shfl_sync(0xffffffffu, d, src);
                long long ci = r0 + src;
                if (cnt == k && !lex_less(cd, ci, worst_d, worst_i)) continue;   // threshold moved meanwhile
                // position = number of list entries smaller than the candidate
                int part = 0;
                for (int i = lane; i < cnt; i += 32) part += lex_less(my_d[i], my_i[i], cd, ci) ? 1 : 0;
                for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
                int pos = part;
                int newcnt = cnt < k ? cnt + 1 : k;
                // shift [pos, newcnt-1) one slot to the right (read all, then write)
                for (int base = ((newcnt - 2 - pos) / 32) * 32 + pos; base >= pos; base -= 32) {
                    int i = base + lane;
                    double td = 0.0;
                    long long ti = 0;
                    bool mv = i <= newcnt - 2;
                    if (mv) { td = my_d[i]; ti = my_i[i]; }
                    __syncwarp();
                    if (mv) { my_d[i + 1] = td; my_i[i + 1] = ti; }
                    __syncwarp();
                }
                if (lane == 0) { my_d[pos] = cd; my_i[pos] = ci; }
                __syncwarp();
                cnt = newcnt;
                if (cnt == k) { worst_d = my_d[k - 1]; worst_i = my_i[k - 1]; }
            }
        }
        // pad short lists, then rank-merge the nw lists
        for (int i = cnt + lane; i < k; i += 32) { my_d[i] = CUDART_INF; my_i[i] = 0x7fffffffffffffffll; }
        __syncthreads();
        const int tot = nw * k;
        for (int e = tid; e < tot; e += KT) {
            double de = l_d[e];
            long long ie = l_i[e];
            if (ie == 0x7fffffffffffffffll) continue;
            int rank = 0;
            for (int j = 0; j < tot; ++j) rank += lex_less(l_d[j], l_i[j], de, ie) ? 1 : 0;
            if (rank < k) {
                size_t w = ((size_t)o * K + t) * k + rank;
                out_idx[w] = ie;
                if (out_dist) {
                    double dd = de;
                    if (pmode == 2) dd = sqrt(de);
                    else if (pmode == 3) dd = pow(de, 1.0 / p);
                    out_dist[w] = dd;
                }
            }
        }
    }
}


// ---- fp32 scan + float64 re-rank ---------------------------------------------------------------------
// Scan: one thread owns R queries (features in registers) and keeps, per query, the KC = k + 8 smallest fp32
// squared distances seen so far in a small local-memory list; the training rows are staged tile by tile into
// shared memory with TMA bulk copies and broadcast to all threads, so each row costs one compare per query in the
// common case.  Re-rank: the candidates are re-evaluated in float64 (numpy's bits) and ordered by (distance,
// index).  The result is accepted only if the k-th exact distance is strictly below the smallest distance any
// EXCLUDED row can have, d >= sqrt(tau (1 - 5e-7)) - 2^-24 |q|, where tau is the fp32 list threshold (fp32
// evaluation error and float rounding of the query, triangle inequality).  Otherwise the (query, tree) pair is
// appended to a list that k_knn_exact re-does from scratch.
constexpr int KS_T = 256;       // threads per CTA
constexpr int KS_R = 4;         // queries per thread
constexpr int KS_TM = 1024;     // rows per shared-memory tile
constexpr int KS_KCMAX = 64;    // candidate list capacity

__device__ __forceinline__ uint32_t ks_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ks_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nKW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra KD_%=;\nbra KW_%=;\nKD_%=:\n}\n" ::"r"(
            ks_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// packed FP32 (two queries per 64-bit register pair): the same roundings as the scalar __fsub_rn / __fmul_rn / __fmaf_rn
typedef unsigned long long kf2;
__device__ __forceinline__ kf2 ks_pack2(float a, float b) {
    kf2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void ks_unpack2(kf2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ kf2 ks_add2(kf2 a, kf2 b) {
    kf2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ kf2 ks_mul2(kf2 a, kf2 b) {
    kf2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ kf2 ks_fma2(kf2 a, kf2 b, kf2 c) {
    kf2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int NF>
__global__ void __launch_bounds__(KS_T, 2) k_knn_scan(const float* __restrict__ feats, int64_t tstride, int64_t Nm,
                                                      const double* __restrict__ q, int64_t No, int KC,
                                                      int64_t rows_per_split, float* __restrict__ cand_d,
                                                      int* __restrict__ cand_i) {
    extern __shared__ __align__(128) unsigned char ks_raw[];
    float* stage = reinterpret_cast<float*>(ks_raw);                                   // 2 x KS_TM x NF
    uint64_t* bars = reinterpret_cast<uint64_t*>(ks_raw + (size_t)2 * KS_TM * NF * sizeof(float));
    const int tid = threadIdx.x;
    const int t = blockIdx.y, sp = blockIdx.z, nsp = gridDim.z, K = gridDim.y;
    const float* F = feats + (size_t)t * tstride;
    float qf[KS_R][NF];
    kf2 nq2[KS_R / 2][NF];        // (-q, -q') of query pairs: f - q = f + (-q) with the same rounding
    float ld[KS_R][KS_KCMAX];
    int li[KS_R][KS_KCMAX];
    float tau[KS_R];
    int pmax[KS_R];
    int64_t oq[KS_R];
#pragma unroll
    for (int r = 0; r < KS_R; ++r) {
        int64_t o = (int64_t)blockIdx.x * (KS_T * KS_R) + (int64_t)r * KS_T + tid;
        oq[r] = o;
        int64_t oo = o < No ? o : No - 1;
#pragma unroll
        for (int b = 0; b < NF; ++b) qf[r][b] = (float)q[oo * NF + b];
        for (int c = 0; c < KC; ++c) { ld[r][c] = CUDART_INF_F; li[r][c] = -1; }
        tau[r] = CUDART_INF_F;
        pmax[r] = 0;
    }
#pragma unroll
    for (int r = 0; r < KS_R / 2; ++r)
#pragma unroll
        for (int b = 0; b < NF; ++b) nq2[r][b] = ks_pack2(-qf[2 * r][b], -qf[2 * r + 1][b]);
    const int64_t row0 = (int64_t)sp * rows_per_split;
    int64_t row1 = row0 + rows_per_split;
    if (row1 > Nm) row1 = Nm;
    const int nt = (int)((row1 - row0 + KS_TM - 1) / KS_TM);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ks_smem_u32(&bars[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int it) {
        int64_t first = row0 + (int64_t)it * KS_TM;
        int cnt = (int)((row1 - first) < KS_TM ? (row1 - first) : KS_TM);
        uint32_t bytes = ((uint32_t)cnt * NF * sizeof(float) + 15u) & ~15u;   // tail reads stay inside the K x Nm x NF array + pad
        uint64_t* bar = &bars[it & 1];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ks_smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ks_smem_u32(stage + (size_t)(it & 1) * KS_TM * NF)),
                     "l"(F + first * NF), "r"(bytes), "r"(ks_smem_u32(bar))
                     : "memory");
    };
    if (tid == 0) {
        for (int it = 0; it < 2 && it < nt; ++it) issue(it);
    }
    for (int it = 0; it < nt; ++it) {
        ks_mbar_wait(&bars[it & 1], (uint32_t)((it >> 1) & 1));
        const float* tile = stage + (size_t)(it & 1) * KS_TM * NF;
        const int64_t first = row0 + (int64_t)it * KS_TM;
        const int cnt = (int)((row1 - first) < KS_TM ? (row1 - first) : KS_TM);
        auto one_row = [&](const float* f, int jj) {
            float dd[KS_R];
#pragma unroll
            for (int rp = 0; rp < KS_R / 2; ++rp) {       // two queries per packed instruction
                kf2 d = ks_add2(ks_pack2(f[0], f[0]), nq2[rp][0]);
                kf2 acc = ks_mul2(d, d);
#pragma unroll
                for (int b = 1; b < NF; ++b) {
                    d = ks_add2(ks_pack2(f[b], f[b]), nq2[rp][b]);
                    acc = ks_fma2(d, d, acc);
                }
                ks_unpack2(acc, dd[2 * rp], dd[2 * rp + 1]);
            }
#pragma unroll
            for (int r = 0; r < KS_R; ++r) {
                const float d2 = dd[r];
                if (d2 < tau[r]) {            // rare after warm-up: replace the current maximum of the list
                    ld[r][pmax[r]] = d2;
                    li[r][pmax[r]] = (int)(first + jj);
                    float mx = -1.f;
                    int pm = 0;
                    for (int c = 0; c < KC; ++c) {
                        float v = ld[r][c];
                        if (v > mx) { mx = v; pm = c; }
                    }
                    tau[r] = mx;
                    pmax[r] = pm;
                }
            }
        };
#pragma unroll 2
        for (int jj = 0; jj < cnt; ++jj) {
            float f[NF];
#pragma unroll
            for (int b = 0; b < NF; ++b) f[b] = tile[jj * NF + b];
            one_row(f, jj);
        }
        __syncthreads();
        if (tid == 0 && it + 2 < nt) issue(it + 2);
    }
#pragma unroll
    for (int r = 0; r < KS_R; ++r) {
        if (oq[r] < No) {
            size_t base = (((size_t)oq[r] * K + t) * nsp + sp) * KC;
            for (int c = 0; c < KC; ++c) { cand_d[base + c] = ld[r][c]; cand_i[base + c] = li[r][c]; }
        }
    }
}

// one warp per (query, tree): exact float64 distances of the candidates, order by (distance, index), verify
// tc_aux != null: the candidates come from the tensor-core scan (fzb_knn_tc.cu): cand_d holds v = |f'|^2 - 2 f'.q' in the
// frame centred on tc_aux[0..NF) and its error is absolute, |v - v_true| <= tc_c0 (|q'|^2 + max_j |f'_j|^2)
// (tc_aux[FZB_FAST_MAXF + tree] = that maximum); tc_stat receives the largest error seen, in the same units.
__global__ void k_knn_rerank(const float* __restrict__ feats, int64_t tstride, int K, int64_t Nm, int NF, const double* __restrict__ q,
                             int64_t No, int k, int KC, int nsp, const float* __restrict__ cand_d,
                             const int* __restrict__ cand_i, int64_t* __restrict__ out_idx,
                             double* __restrict__ out_dist, int64_t* __restrict__ redo, int* __restrict__ n_redo,
                             const double* __restrict__ tc_aux, double tc_c0, unsigned long long* __restrict__ tc_stat) {
    extern __shared__ double rs[];
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, w = threadIdx.x >> 5;
    const int C = KC * nsp;
    double* sd = rs + (size_t)w * C * 2;
    long long* si = reinterpret_cast<long long*>(sd + C);
    const int64_t item = (int64_t)blockIdx.x * wpb + w;
    if (item >= No * K) return;
    const int64_t o = item / K;
    const int t = (int)(item % K);
    const float* F = feats + (size_t)t * tstride;
    const size_t base = (size_t)item * C;
    double qn = 0.0, qc = 0.0;
    for (int b = 0; b < NF; ++b) {
        qn += q[o * NF + b] * q[o * NF + b];
        if (tc_aux) { const double d = q[o * NF + b] - tc_aux[b]; qc += d * d; }
    }
    const double eq = sqrt(qn) * 5.9604644775390625e-08;     // 2^-24 |q|
    const double tc_scale = tc_aux ? qc + tc_aux[FZB_FAST_MAXF + t] : 0.0;
    double worst = 0.0;
    float tau = CUDART_INF_F;          // smallest per-split threshold = smallest fp32 distance of any excluded row
    for (int s = 0; s < nsp; ++s) {
        float mx = -CUDART_INF_F;
        for (int c = lane; c < KC; c += 32) mx = fmaxf(mx, cand_d[base + (size_t)s * KC + c]);
        for (int sh = 16; sh > 0; sh >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
        tau = fminf(tau, mx);
    }
    int nvalid = 0;
    for (int c = lane; c < C; c += 32) {
        int r = cand_i[base + c];
        double acc = CUDART_INF;
        if (r >= 0) {
            acc = 0.0;
            for (int b = 0; b < NF; ++b) {
                double df = __dsub_rn((double)F[(size_t)r * NF + b], q[o * NF + b]);
                acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
            if (isnan(acc)) acc = CUDART_INF;
            ++nvalid;
            if (tc_aux && isfinite(acc)) worst = fmax(worst, fabs((double)cand_d[base + c] + qc - acc));
        }
        sd[c] = acc;
        si[c] = r >= 0 ? r : 0x7fffffffffffffffll;
    }
    if (tc_aux && tc_stat) {
        worst = tc_scale > 0.0 ? worst / tc_scale : 0.0;
        for (int sh = 16; sh > 0; sh >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, sh));
        if (lane == 0 && worst > 0.0) atomicMax(tc_stat, (unsigned long long)__double_as_longlong(worst));
    }
    for (int sh = 16; sh > 0; sh >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, sh);
    __syncwarp();
    double dk = CUDART_INF;
    for (int c = lane; c < C; c += 32) {
        double de = sd[c];
        long long ie = si[c];
        if (ie == 0x7fffffffffffffffll) continue;
        int rank = 0;
        for (int j = 0; j < C; ++j) rank += lex_less(sd[j], si[j], de, ie) ? 1 : 0;
        if (rank < k) {
            size_t wout = (size_t)item * k + rank;
            out_idx[wout] = ie;
            if (out_dist) out_dist[wout] = sqrt(de);
            if (rank == k - 1) dk = de;
        }
    }
    for (int sh = 16; sh > 0; sh >>= 1) dk = fmin(dk, __shfl_xor_sync(0xffffffffu, dk, sh));
    // accept only if no excluded row can belong to the top k
    bool ok = nvalid >= k && isfinite(dk);
    if (ok && !isinf(tau)) {
        if (tc_aux) {
            ok = dk * (1.0 + 1e-12) < (double)tau + qc - tc_c0 * tc_scale;
        } else {
            double dmin_excl = sqrt((double)tau * (1.0 - 5e-7)) - eq;
            ok = sqrt(dk) < dmin_excl;
        }
    }
    if (!ok && lane == 0) redo[atomicAdd(n_redo, 1)] = item;
}

// ordered union: one warp per object
__global__ void k_union(const int64_t* __restrict__ idx, int64_t No, int W, int64_t* __restrict__ nbr,
                        int64_t* __restrict__ nnbr) {
    const int lane = threadIdx.x & 31;
    const int64_t o = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (o >= No) return;
    const int64_t* v = idx + (size_t)o * W;
    int64_t* out = nbr + (size_t)o * W;
    int base = 0;
    for (int i0 = 0; i0 < W; i0 += 32) {
        int i = i0 + lane;
        bool keep = false;
        if (i < W) {
            long long a = v[i];
            keep = true;
            for (int j = 0; j < i; ++j)
                if (v[j] == a) { keep = false; break; }
        }
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) out[base + __popc(bal & ((1u << lane) - 1u))] = v[i];
        base += __popc(bal);
    }
    __syncwarp();
    for (int i = base + lane; i < W; i += 32) out[i] = -99;   // knn.py:343
    if (lane == 0) nnbr[o] = base;
}

}  // namespace

static int knn_exact_launch(fzb_context* h, const double* d_q, int64_t No, int k, double p, int pmode, int64_t* d_idx,
                            double* d_dist, const int64_t* items, const int* n_items, int64_t max_items) {
    size_t smem = sizeof(double) * FZB_MAXF + (size_t)(KT / 32) * k * 16;
    FZB_CHECK(smem <= 200 * 1024, "k=%d too large for the kNN kernel", k);
    FZB_CUDA(cudaFuncSetAttribute(k_knn_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = max_items < (int64_t)h->sm_count * 8 ? max_items : (int64_t)h->sm_count * 8;
    if (grid < 1) grid = 1;
    k_knn_exact<<<(unsigned)grid, KT, smem, h->stream>>>(h->knn_feats.as<float>(), h->knn_stride, h->knn_K, h->knn_Nm, h->knn_Nf, d_q,
                                                         No, k, p, pmode, d_idx, d_dist, items, n_items);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

template <int NF>
static int knn_scan_launch(fzb_context* h, const double* d_q, int64_t No, int KC, int nsp, int64_t rows_per_split,
                           float* cand_d, int* cand_i) {
    size_t smem = (size_t)2 * KS_TM * NF * sizeof(float) + 2 * sizeof(uint64_t);
    FZB_CUDA(cudaFuncSetAttribute(k_knn_scan<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((No + KS_T * KS_R - 1) / (KS_T * KS_R)), (unsigned)h->knn_K, (unsigned)nsp);
    k_knn_scan<NF><<<grid, KS_T, smem, h->stream>>>(h->knn_feats.as<float>(), h->knn_stride, h->knn_Nm, d_q, No, KC, rows_per_split,
                                                    cand_d, cand_i);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

// error bound of the tensor-core scan's v, in units of (|q'|^2 + max |f'|^2): see fzb_knn_tc.cu
static double env_c0() {
    const char* e = getenv("FZB_KNN_TC_C0");
    return e ? atof(e) : 4e-6;
}

int fzb_knn_query_dev(fzb_context* h, const double* d_q, int64_t No, int k, double p, int64_t* d_idx, double* d_dist) {
    int pmode = (p == 2.0) ? 2 : (p == 1.0) ? 1 : (!(p > 0) || std::isinf(p)) ? 0 : 3;
    const int nf = h->knn_Nf;
    const int K = h->knn_K;
    const int64_t Nm = h->knn_Nm;
    const int KC = k + 8;
    // tensor-core candidate scan (fzb_knn_tc.cu): measured SLOWER than the fp32 CUDA-core scan on B200 (1.24e12 against
    // 1.42e12 distance evaluations/s at 1M rows x 20 trees: every accumulator has to be read back from TMEM at 64 B/clk/SM),
    // so it is opt-in (FZB_KNN_TC=1) and the CUDA-core scan is the product path
    const bool use_tc = pmode == 2 && h->knn_tc_valid && KC <= fzb_knn_tc_kcmax() && getenv("FZB_KNN_EXACT_ONLY") == nullptr;
    const bool fast = use_tc || (pmode == 2 && nf >= 4 && nf <= 6 && KC <= KS_KCMAX && Nm >= 4096 &&
                                 Nm < ((int64_t)1 << 31) && getenv("FZB_KNN_EXACT_ONLY") == nullptr);
    if (!fast) return knn_exact_launch(h, d_q, No, k, p, pmode, d_idx, d_dist, nullptr, nullptr, No * K);

    // row splits so that small query batches still fill the GPU
    const int64_t qper = use_tc ? 128 : KS_T * KS_R, rtile = use_tc ? 256 : KS_TM;
    int64_t qtiles = (No + qper - 1) / qper;
    int64_t want = (int64_t)h->sm_count * (use_tc ? 2 : 8);
    int64_t nsp = (want + qtiles * K - 1) / (qtiles * K);
    int64_t max_sp = (Nm + 4 * rtile - 1) / (4 * rtile);
    if (nsp > max_sp) nsp = max_sp;
    if (nsp > 32) nsp = 32;
    if (nsp < 1) nsp = 1;
    int64_t rows_per_split = ((Nm + nsp - 1) / nsp + rtile - 1) / rtile * rtile;
    nsp = (Nm + rows_per_split - 1) / rows_per_split;
    // process the queries in chunks that bound the candidate buffers (~2 GB)
    const int nlist = (int)nsp * (use_tc ? fzb_knn_tc_lists() : 1);     // candidate lists per (query, tree)
    size_t per_q = (size_t)K * nlist * KC * 8;
    int64_t chunk = (int64_t)(((size_t)2 << 30) / per_q);
    if (chunk < KS_T * KS_R) chunk = KS_T * KS_R;
    if (chunk > No) chunk = No;
    if (h->knn_cand.reserve((size_t)chunk * per_q + 256) || h->knn_redo.reserve((size_t)chunk * K * 8 + 64)) return 1;
    float* cand_d = h->knn_cand.as<float>();
    int* cand_i = reinterpret_cast<int*>(cand_d + (size_t)chunk * K * nlist * KC);
    int* n_redo = h->knn_redo.as<int>();
    int64_t* redo = reinterpret_cast<int64_t*>(n_redo + 4);
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        int64_t nc = No - o0 < chunk ? No - o0 : chunk;
        const double* qq = d_q + o0 * nf;
        int rc = use_tc ? fzb_knn_tc_scan(h, qq, nc, KC, (int)nsp, (int)(rows_per_split / rtile), cand_d, cand_i)
               : nf == 4 ? knn_scan_launch<4>(h, qq, nc, KC, (int)nsp, rows_per_split, cand_d, cand_i)
               : nf == 5 ? knn_scan_launch<5>(h, qq, nc, KC, (int)nsp, rows_per_split, cand_d, cand_i)
                         : knn_scan_launch<6>(h, qq, nc, KC, (int)nsp, rows_per_split, cand_d, cand_i);
        if (rc) return rc;
        FZB_CUDA(cudaMemsetAsync(n_redo, 0, 16, h->stream));
        const int wpb = 4;
        size_t smem = (size_t)wpb * KC * nlist * 16;
        FZB_CHECK(smem <= 200 * 1024, "kNN re-rank: too many candidates");
        FZB_CUDA(cudaFuncSetAttribute(k_knn_rerank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t items = nc * K;
        k_knn_rerank<<<(unsigned)((items + wpb - 1) / wpb), wpb * 32, smem, h->stream>>>(
            h->knn_feats.as<float>(), h->knn_stride, K, Nm, nf, qq, nc, k, KC, nlist, cand_d, cand_i, d_idx + (size_t)o0 * K * k,
            d_dist ? d_dist + (size_t)o0 * K * k : nullptr, redo, n_redo, use_tc ? h->knn_aux.as<double>() : nullptr,
            env_c0(), use_tc ? reinterpret_cast<unsigned long long*>(n_redo + 2) : nullptr);
        fzb_count_launch(h);
        FZB_CUDA(cudaGetLastError());
        // the rare (query, tree) pairs that failed the exactness test are re-done by the float64 kernel
        if (knn_exact_launch(h, qq, nc, k, p, pmode, d_idx + (size_t)o0 * K * k,
                             d_dist ? d_dist + (size_t)o0 * K * k : nullptr, redo, n_redo, 4096))
            return 1;
        int nr[4] = {0, 0, 0, 0};
        FZB_CUDA(cudaMemcpyAsync(nr, n_redo, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        FZB_CUDA(cudaStreamSynchronize(h->stream));
        h->stats.knn_redo += nr[0];
        if (use_tc) {
            double w;
            memcpy(&w, nr + 2, 8);
            if (w > h->stats.knn_tc_err) h->stats.knn_tc_err = w;
        }
        h->stats.knn_tc = use_tc ? 1 : 0;
    }
    return 0;
}

int fzb_knn_union_dev(fzb_context* h, const int64_t* d_idx, int64_t No, int Kk, int64_t* d_neighbors,
                      int64_t* d_nneighbors) {
    int wpb = 8;
    int64_t grid = (No + wpb - 1) / wpb;
    k_union<<<(unsigned)grid, wpb * 32, 0, h->stream>>>(d_idx, No, Kk, d_neighbors, d_nneighbors);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}
