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
__global__ void __launch_bounds__(KT) k_knn_exact(const float* __restrict__ feats, int K, int64_t Nm, int Nf,
                                                  const double* __restrict__ q, int64_t No, int k, double p, int pmode,
                                                  int64_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    extern __shared__ double sm[];
    const int nw = KT / 32;
    double* s_q = sm;                                  // Nf
    double* l_d = s_q + FZB_MAXF;                      // nw * k
    long long* l_i = reinterpret_cast<long long*>(l_d + (size_t)nw * k);   // nw * k
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* my_d = l_d + (size_t)warp * k;
    long long* my_i = l_i + (size_t)warp * k;

    for (int64_t item = blockIdx.x; item < No * K; item += gridDim.x) {
        const int64_t o = item / K;
        const int t = (int)(item % K);
        __syncthreads();
        if (tid < Nf) s_q[tid] = q[o * Nf + tid];
        __syncthreads();
        const float* F = feats + (size_t)t * Nm * Nf;
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

int fzb_knn_query_dev(fzb_context* h, const double* d_q, int64_t No, int k, double p, int64_t* d_idx, double* d_dist) {
    int pmode = (p == 2.0) ? 2 : (p == 1.0) ? 1 : (!(p > 0) || std::isinf(p)) ? 0 : 3;
    size_t smem = sizeof(double) * FZB_MAXF + (size_t)(KT / 32) * k * 16;
    FZB_CHECK(smem <= 200 * 1024, "k=%d too large for the kNN kernel", k);
    FZB_CUDA(cudaFuncSetAttribute(k_knn_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t items = No * h->knn_K;
    int64_t grid = items < (int64_t)h->sm_count * 8 ? items : (int64_t)h->sm_count * 8;
    k_knn_exact<<<(unsigned)grid, KT, smem, h->stream>>>(h->knn_feats.as<float>(), h->knn_K, h->knn_Nm, h->knn_Nf, d_q,
                                                         No, k, p, pmode, d_idx, d_dist);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
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
