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


// ---- fp32 threshold filter + float64 re-rank ---------------------------------------------------------
// Candidate search for one (query, tree): the KC = k + 8 rows with the smallest fp32 squared distances.  A streaming
// top-KC list per thread (round 1) spent 85 % of its issued warp instructions in the divergent insert path
// (profiles/r2_knn_scan_ncu.md: 12.4 active threads per instruction, 1.1e9 local-memory loads per launch).  Here the
// scan keeps no list at all: a row is APPENDED to a per-(query, tree) buffer when its distance is <= a threshold tau
// that is known to be >= the KC-th smallest distance, and the exact KC smallest are selected from the buffer
// afterwards.  tau comes from the same machinery applied to nested prefixes of the rows, 512 -> x16 -> x16 -> all
// (stage 0 appends every row of the 512-row prefix, k_knn_select returns its KC-th smallest distance, which bounds
// the KC-th smallest of any superset): a stage appends KC x ratio <= 528 rows on average whatever the size of the
// set.  The scan copy of the features stores row r' = 64-way interleave of the original order (r' -> (r' % 64) m +
// r' / 64, m = ceil(Nm / 64)), so that a prefix samples the whole set even when the caller's rows are sorted, every
// value twice (v, v) so that one LDS.128 yields two packed f32x2 operands, rows padded to 16 bytes (NaN rows beyond
// Nm: never appended).  An overflowing buffer (more than `cap` rows under tau: adversarial row order) sends the
// (query, tree) pair to the all-float64 kernel, like a failed exactness test.
// Re-rank: the candidates are re-evaluated in float64 (numpy's bits) and ordered by (distance, index).  The result
// is accepted only if the k-th exact distance is strictly below the smallest distance any EXCLUDED row can have,
// d >= sqrt(tau (1 - 5e-7)) - 2^-24 |q|, where tau is the largest fp32 distance of the candidates (every excluded
// row has an fp32 distance >= it; fp32 evaluation error and float rounding of the query, triangle inequality).
// Otherwise the (query, tree) pair is appended to a list that k_knn_exact re-does from scratch.
constexpr int KS_T = 256;       // threads per CTA
constexpr int KS_R = 4;         // queries per thread (difference form)
constexpr int KS_RDOT = 8;      // queries per thread (dot form)
constexpr int KS_TM = 1024;     // rows per shared-memory tile
constexpr int KS_KCMAX = 256;   // largest k + 8
constexpr int KS_IL = 64;       // interleave factor of the scan copy
constexpr int KS_P0 = 512;      // rows of the first prefix (all appended)
constexpr int KS_CAP = 2048;    // largest number of buffered rows per (query, tree) (select kernel's shared memory)
// Two forms of the fp32 distance.  DIFF: sum_b (f_b - q_b)^2, 2 NF FMA-pipe operations per distance (relative error
// 5e-7).  DOT: v = |f'|^2 - 2 f'.q' in the frame centred on the mean training feature c (f' = float(f - c), q' =
// float(q - c); |f - q|^2 = v + |q'|^2), NF FMA-pipe operations per distance with |f'|^2 stored next to the row; its
// error is absolute, |v_fp32 - v| <= c0 (|f'|^2 + |q'|^2) with c0 = 2e-6 >= (2 NF + 1) 2^-24, and the re-rank's
// exactness test accounts for it without a global bound on |f'| (see k_knn_rerank).
enum { KS_DIFF = 0, KS_DOT = 2 };
__host__ __device__ constexpr int ks_rowq(int nf, int form = KS_DIFF) { return ((form == KS_DOT ? nf + 1 : nf) * 8 + 15) / 16; }   // 16-byte words per scan row
constexpr double KS_C0 = 2e-6;

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

// scan copy: [tree][row r'][ks_rowq(NF, form)] 16-byte words = (v, v) per band (DOT: of f', then (|f'|^2, |f'|^2)), zero padded
__global__ void k_knn_build_scan(const float* __restrict__ feats, int64_t tstride, int64_t Nm, int NF, int K, int64_t m,
                                 int64_t Ns, int form, const double* __restrict__ centre, float* __restrict__ scan) {
    const int rq = ks_rowq(NF, form);
    const int64_t total = (int64_t)K * Ns;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(e / Ns);
        const int64_t rp = e % Ns;
        const int64_t i = (rp % KS_IL) * m + rp / KS_IL;
        float* dst = scan + (size_t)e * rq * 4;
        double ff = 0.0;
        for (int b = 0; b < rq * 2; ++b) {
            float v = 0.f;
            if (b < NF) {
                v = i < Nm ? feats[(size_t)t * tstride + (size_t)i * NF + b] : CUDART_NAN_F;
                if (form == KS_DOT) { v = (float)((double)v - centre[b]); ff += (double)v * (double)v; }
            } else if (b == NF && form == KS_DOT) {
                v = (float)ff;
            }
            dst[2 * b] = v;
            dst[2 * b + 1] = v;
        }
    }
}

// mean finite feature over all trees (the centre of the DOT form)
__global__ void k_knn_mean(const float* feats, int64_t stride, int64_t Nm, int K, int Nf, double* sums) {
    double acc[FZB_MAXF + 1];
    for (int b = 0; b <= Nf; ++b) acc[b] = 0.0;
    const int64_t total = (int64_t)K * Nm;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(g / Nm);
        const float* f = feats + (size_t)t * stride + (size_t)(g - (int64_t)t * Nm) * Nf;
        bool fin = true;
        for (int b = 0; b < Nf; ++b) fin = fin && isfinite(f[b]);
        if (fin) {
            for (int b = 0; b < Nf; ++b) acc[b] += (double)f[b];
            acc[Nf] += 1.0;
        }
    }
    for (int b = 0; b <= Nf; ++b) atomicAdd(sums + b, acc[b]);
}
__global__ void k_knn_mean_finish(double* sums, int Nf) {
    if (threadIdx.x < Nf) sums[threadIdx.x] = sums[Nf] > 0.0 ? sums[threadIdx.x] / sums[Nf] : 0.0;
}

// One thread owns R queries (features in registers) of one tree and one row split; the rows are staged tile by
// tile into shared memory with TMA bulk copies and broadcast to all threads.  Rows with distance <= tau are appended
// (fp32 distance, scan row) to the buffer of the (query, tree, split).
// A launch covers the trees [t0, t0 + gridDim.y); buffers, counts and thresholds are indexed (query, tree - t0).
// tau_q != null: the threshold of every tree of the launch is tau_fac x the squared-distance threshold tau_q[query]
// (the k+8-th smallest distance found in another tree: the trees are noise realisations of one training set).
template <int NF, int FORM, int R>
__global__ void __launch_bounds__(KS_T, 2) k_knn_filter(const float* __restrict__ scan, int64_t Ns, int64_t nrows,
                                                        const double* __restrict__ q, const double* __restrict__ centre,
                                                        int64_t No, const float* __restrict__ tau_in, int64_t rows_per_split,
                                                        int cap, uint2* __restrict__ buf, int* __restrict__ cnt_out, int t0,
                                                        const float* __restrict__ tau_q, float tau_fac) {
    constexpr int RQ = ks_rowq(NF, FORM);
    extern __shared__ __align__(128) unsigned char ks_raw[];
    ulonglong2* stage = reinterpret_cast<ulonglong2*>(ks_raw);                        // 2 x KS_TM x RQ
    uint64_t* bars = reinterpret_cast<uint64_t*>(ks_raw + (size_t)2 * KS_TM * RQ * 16);
    const int tid = threadIdx.x;
    const int t = blockIdx.y, sp = blockIdx.z, nsp = gridDim.z, K = gridDim.y;
    const ulonglong2* F = reinterpret_cast<const ulonglong2*>(scan) + (size_t)(t0 + t) * Ns * RQ;
    // DIFF: (-q, -q') of query pairs, f - q = f + (-q) with the same rounding; DOT: (-2 q', -2 q'')
    kf2 nq2[R / 2][NF];
    float tau[R];
    int cnt[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t o = (int64_t)blockIdx.x * (KS_T * R) + (int64_t)r * KS_T + tid;
        tau[r] = (o < No) ? (tau_in ? tau_in[o * K + t] : CUDART_INF_F) : -CUDART_INF_F;     // padding queries append nothing
        if (tau_q && o < No) {
            // DOT: the thresholds are v = d^2 - |q'|^2, so d^2 is scaled: v' = fac v + (fac - 1) |q'|^2
            float qq = 0.f;
            if (FORM == KS_DOT)
                for (int b = 0; b < NF; ++b) { const float qc = (float)(q[o * NF + b] - centre[b]); qq = fmaf(qc, qc, qq); }
            tau[r] = fmaf(tau_fac, tau_q[o], (tau_fac - 1.f) * qq);
        }
        cnt[r] = 0;
    }
#pragma unroll
    for (int r = 0; r < R / 2; ++r) {
        const int64_t o0 = (int64_t)blockIdx.x * (KS_T * R) + (int64_t)(2 * r) * KS_T + tid, o1 = o0 + KS_T;
        const int64_t a0 = o0 < No ? o0 : No - 1, a1 = o1 < No ? o1 : No - 1;
#pragma unroll
        for (int b = 0; b < NF; ++b) {
            if (FORM == KS_DOT) nq2[r][b] = ks_pack2(-2.f * (float)(q[a0 * NF + b] - centre[b]), -2.f * (float)(q[a1 * NF + b] - centre[b]));
            else nq2[r][b] = ks_pack2(-(float)q[a0 * NF + b], -(float)q[a1 * NF + b]);
        }
    }
    const int64_t row0 = (int64_t)sp * rows_per_split;
    int64_t row1 = row0 + rows_per_split;
    if (row1 > nrows) row1 = nrows;
    const int nt = row1 > row0 ? (int)((row1 - row0 + KS_TM - 1) / KS_TM) : 0;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ks_smem_u32(&bars[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int it) {
        const int64_t first = row0 + (int64_t)it * KS_TM;
        const int n = (int)((row1 - first) < KS_TM ? (row1 - first) : KS_TM);
        const uint32_t bytes = (uint32_t)n * RQ * 16;
        uint64_t* bar = &bars[it & 1];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ks_smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ks_smem_u32(stage + (size_t)(it & 1) * KS_TM * RQ)),
                     "l"(F + first * RQ), "r"(bytes), "r"(ks_smem_u32(bar))
                     : "memory");
    };
    if (tid == 0) {
        for (int it = 0; it < 2 && it < nt; ++it) issue(it);
    }
    const int64_t rstride = (int64_t)KS_T * K * nsp * cap;        // buffer entries between consecutive queries of a thread
    uint2* const buf0 = buf + ((((int64_t)blockIdx.x * (KS_T * R) + tid) * K + t) * nsp + sp) * cap;
    auto append = [&](int r, float d2, int row) {
        if (cnt[r] < cap) buf0[r * rstride + cnt[r]] = make_uint2(__float_as_uint(d2), (unsigned)row);
        ++cnt[r];
    };
    for (int it = 0; it < nt; ++it) {
        ks_mbar_wait(&bars[it & 1], (uint32_t)((it >> 1) & 1));
        const ulonglong2* tile = stage + (size_t)(it & 1) * KS_TM * RQ;
        const int64_t first = row0 + (int64_t)it * KS_TM;
        const int n = (int)((row1 - first) < KS_TM ? (row1 - first) : KS_TM);
#pragma unroll 4
        for (int jj = 0; jj < n; ++jj) {
            kf2 f[2 * RQ];
#pragma unroll
            for (int i = 0; i < RQ; ++i) { const ulonglong2 v = tile[jj * RQ + i]; f[2 * i] = v.x; f[2 * i + 1] = v.y; }
            float dd[R];
#pragma unroll
            for (int rp = 0; rp < R / 2; ++rp) {       // two queries per packed instruction
                kf2 acc;
                if (FORM == KS_DOT) {
                    acc = f[NF];
#pragma unroll
                    for (int b = 0; b < NF; ++b) acc = ks_fma2(f[b], nq2[rp][b], acc);
                } else {
                    kf2 d = ks_add2(f[0], nq2[rp][0]);
                    acc = ks_mul2(d, d);
#pragma unroll
                    for (int b = 1; b < NF; ++b) {
                        d = ks_add2(f[b], nq2[rp][b]);
                        acc = ks_fma2(d, d, acc);
                    }
                }
                ks_unpack2(acc, dd[2 * rp], dd[2 * rp + 1]);
            }
            bool any = false;
#pragma unroll
            for (int r = 0; r < R; ++r) any = any || (dd[r] <= tau[r]);
            if (any) {            // rare: ~KC x ratio rows per (query, tree) and stage
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float d2 = dd[r];
                    asm volatile("" : "+f"(d2));      // keeps the per-query compares out of the common path
                    if (d2 <= tau[r]) append(r, d2, (int)(first + jj));
                }
            }
        }
        __syncthreads();
        if (tid == 0 && it + 2 < nt) issue(it + 2);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t o = (int64_t)blockIdx.x * (KS_T * R) + (int64_t)r * KS_T + tid;
        if (o < No) cnt_out[((size_t)o * K + t) * nsp + sp] = cnt[r];
    }
}

// One warp per (query, tree): the KC-th smallest fp32 distance of the buffered rows (bisection on the order-preserving
// integer image of the floats) -> tau_out; final stage: the KC smallest rows -> cand_d / cand_i (ORIGINAL row indices).
// A pair whose buffer overflowed or holds fewer than KC rows gets tau_out = +inf (intermediate stage: the next one
// overflows too) and an empty candidate list (final stage: the re-rank sends it to the all-float64 kernel).
// order-preserving map of float bit patterns to unsigned integers (negative values: the DOT form) and back
__device__ __forceinline__ unsigned int ks_key(unsigned int b) { return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ unsigned int ks_unkey(unsigned int k) { return (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k; }

__global__ void __launch_bounds__(256) k_knn_select(const uint2* __restrict__ buf, const int* __restrict__ cnt, int nsp, int cap,
                                                    int64_t nitems, int KC, float* __restrict__ tau_out,
                                                    float* __restrict__ cand_d, int* __restrict__ cand_i, int64_t m,
                                                    unsigned int* __restrict__ n_overflow, int Kw, int t0, int Ktot) {
    extern __shared__ unsigned int sel_sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    unsigned int* keys = sel_sm + (size_t)w * KS_CAP;
    const int64_t item = (int64_t)blockIdx.x * wpb + w;
    if (item >= nitems) return;
    // candidates go to the slot of (query, tree) among all Ktot trees; the launch covers the trees [t0, t0 + Kw)
    const int64_t oitem = (item / Kw) * Ktot + t0 + (item % Kw);
    int n = 0;
    bool over = false;
    for (int s = 0; s < nsp; ++s) {
        int c = cnt[(size_t)item * nsp + s];
        if (c > cap) { over = true; c = cap; }
        const uint2* src = buf + ((size_t)item * nsp + s) * cap;
        for (int i = lane; i < c; i += 32) keys[n + i] = ks_key(src[i].x);
        n += c;
    }
    __syncwarp();
    const bool bad = over || n < KC;
    if (bad) {
        if (tau_out && lane == 0) tau_out[item] = CUDART_INF_F;
        if (cand_d)
            for (int c = lane; c < KC; c += 32) { cand_d[(size_t)oitem * KC + c] = CUDART_INF_F; cand_i[(size_t)oitem * KC + c] = -1; }
        if (over && cand_d && lane == 0 && n_overflow) atomicAdd(n_overflow, 1u);
        return;
    }
    unsigned int v = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned int trial = v | (1u << bit);
        int c = 0;
        for (int i = lane; i < n; i += 32) c += keys[i] < trial ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c < KC) v = trial;
    }
    // v = the KC-th smallest key: fewer than KC keys are below it, at least KC are <= it
    if (tau_out && lane == 0) tau_out[item] = __uint_as_float(ks_unkey(v));
    if (cand_d) {
        int nless = 0;
        for (int i = lane; i < n; i += 32) nless += keys[i] < v ? 1 : 0;
        nless = __reduce_add_sync(0xffffffffu, nless);
        int at_less = 0, at_eq = nless;        // output positions: the keys < v first, then keys == v up to KC
        for (int s = 0; s < nsp; ++s) {
            int c = cnt[(size_t)item * nsp + s];
            if (c > cap) c = cap;
            const uint2* src = buf + ((size_t)item * nsp + s) * cap;
            for (int i0 = 0; i0 < c; i0 += 32) {
                const int i = i0 + lane;
                uint2 e = make_uint2(0u, 0u);
                if (i < c) e = src[i];
                const unsigned int key = i < c ? ks_key(e.x) : 0xffffffffu;
                const bool lt = key < v, eq = key == v;
                const unsigned bl = __ballot_sync(0xffffffffu, lt), be = __ballot_sync(0xffffffffu, eq);
                const unsigned below = (1u << lane) - 1u;
                int pos = -1;
                if (lt) pos = at_less + __popc(bl & below);
                else if (eq) { pos = at_eq + __popc(be & below); if (pos >= KC) pos = -1; }
                if (pos >= 0) {
                    const int64_t rp = e.y;
                    cand_d[(size_t)oitem * KC + pos] = __uint_as_float(e.x);
                    cand_i[(size_t)oitem * KC + pos] = (int)((rp % KS_IL) * m + rp / KS_IL);
                }
                at_less += __popc(bl);
                at_eq += __popc(be);
            }
        }
    }
}

// one warp per (query, tree): exact float64 distances of the candidates, order by (distance, index), verify
// form KS_DIFF: cand_d = fp32 squared distances (relative error <= 5e-7).
// form 1 (tensor-core scan, fzb_knn_tc.cu): cand_d holds v = |f'|^2 - 2 f'.q' in the frame centred on aux[0..NF) and its
//   error is absolute, |v - v_true| <= c0 (|q'|^2 + max_j |f'_j|^2) (aux[FZB_FAST_MAXF + tree] = that maximum).
// form KS_DOT (filter scan): the same v from the fp32 FMA chain with f' = float(f - c), q' = float(q - c):
//   |v_fp32 - v| <= c0 (|f'_j|^2 + |q'|^2).  With D the k-th exact distance and Q = |q'|, an excluded row with
//   |f'_j| >= A = (D + Q)(1 + 1e-6) is farther than D by the triangle inequality; one with |f'_j| < A has
//   |f'_j - q'|^2 > tau + Q^2 - c0 (A^2 + Q^2) =: L2, i.e. a true distance > sqrt(L2) - 2^-24 (A + Q) (rounding of f'
//   and q').  The result is accepted when D is strictly below that.
// stat receives the largest error seen on the candidates in units of the bound's scale (forms 1 and 2).
__global__ void k_knn_rerank(const float* __restrict__ feats, int64_t tstride, int K, int64_t Nm, int NF, const double* __restrict__ q,
                             int64_t No, int k, int KC, int nsp, const float* __restrict__ cand_d,
                             const int* __restrict__ cand_i, int64_t* __restrict__ out_idx,
                             double* __restrict__ out_dist, int64_t* __restrict__ redo, int* __restrict__ n_redo, int form,
                             const double* __restrict__ aux, double c0, unsigned long long* __restrict__ stat) {
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
        if (form == 1) { const double d = q[o * NF + b] - aux[b]; qc += d * d; }
        if (form == KS_DOT) { const double d = (double)(float)(q[o * NF + b] - aux[b]); qc += d * d; }
    }
    const double eq = sqrt(qn) * 5.9604644775390625e-08;     // 2^-24 |q|
    const double tc_scale = form == 1 ? qc + aux[FZB_FAST_MAXF + t] : 0.0;
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
            double ff = 0.0;
            for (int b = 0; b < NF; ++b) {
                double df = __dsub_rn((double)F[(size_t)r * NF + b], q[o * NF + b]);
                acc = __dadd_rn(acc, __dmul_rn(df, df));
                if (form == KS_DOT) { const double fp = (double)(float)((double)F[(size_t)r * NF + b] - aux[b]); ff += fp * fp; }
            }
            if (isnan(acc)) acc = CUDART_INF;
            ++nvalid;
            if (form == 1 && isfinite(acc)) worst = fmax(worst, fabs((double)cand_d[base + c] + qc - acc));
            if (form == KS_DOT && isfinite(acc) && ff + qc > 0.0) worst = fmax(worst, fabs((double)cand_d[base + c] + qc - acc) / (ff + qc));
        }
        sd[c] = acc;
        si[c] = r >= 0 ? r : 0x7fffffffffffffffll;
    }
    if (form != KS_DIFF && stat) {
        if (form == 1) worst = tc_scale > 0.0 ? worst / tc_scale : 0.0;
        for (int sh = 16; sh > 0; sh >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, sh));
        if (lane == 0 && worst > 0.0) atomicMax(stat, (unsigned long long)__double_as_longlong(worst));
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
        if (form == 1) {
            ok = dk * (1.0 + 1e-12) < (double)tau + qc - c0 * tc_scale;
        } else if (form == KS_DOT) {
            const double D = sqrt(dk), Q = sqrt(qc);
            const double A = (D + Q) * (1.0 + 1e-6);
            const double L2 = (double)tau + qc - c0 * (A * A + qc);
            ok = L2 > 0.0 && D < sqrt(L2) - 5.9604644775390625e-08 * (A + Q);
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

// scan copy of the features (fzb_knn_build): see the layout comment above.  FZB_KNN_FORM=diff selects the difference
// form of the fp32 distance (default: dot)
int fzb_knn_scan_build(fzb_context* h) {
    const int nf = h->knn_Nf, K = h->knn_K;
    const int64_t Nm = h->knn_Nm;
    h->knn_scan_valid = false;
    h->knn_share_off = false;
    if (nf < 4 || nf > 6 || Nm < 4096 || Nm >= ((int64_t)1 << 31) - KS_IL) return 0;
    const char* e = getenv("FZB_KNN_FORM");
    const int form = (e && strcmp(e, "diff") == 0) ? KS_DIFF : KS_DOT;
    const int64_t m = (Nm + KS_IL - 1) / KS_IL, Ns = m * KS_IL;
    if (h->knn_scan.reserve((size_t)K * Ns * ks_rowq(nf, form) * 16 + 256) || h->knn_centre.reserve((FZB_MAXF + 2) * 8)) return 1;
    double* centre = h->knn_centre.as<double>();
    FZB_CUDA(cudaMemsetAsync(centre, 0, (FZB_MAXF + 2) * 8, h->stream));
    if (form == KS_DOT) {
        k_knn_mean<<<h->sm_count * 4, 256, 0, h->stream>>>(h->knn_feats.as<float>(), h->knn_stride, Nm, K, nf, centre);
        k_knn_mean_finish<<<1, 32, 0, h->stream>>>(centre, nf);
        fzb_count_launch(h);
        fzb_count_launch(h);
    }
    k_knn_build_scan<<<h->sm_count * 8, 256, 0, h->stream>>>(h->knn_feats.as<float>(), h->knn_stride, Nm, nf, K, m, Ns, form, centre,
                                                             h->knn_scan.as<float>());
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->knn_m = m;
    h->knn_Ns = Ns;
    h->knn_form = form;
    h->knn_scan_valid = true;
    return 0;
}

struct KnnWindow {
    int t0, Kw;              // trees [t0, t0 + Kw)
    const float* tau_q;      // per-query squared-distance threshold shared by the trees of the window (null: per item)
    float tau_fac;
};

template <int NF, int FORM, int R>
static int knn_filter_launch(fzb_context* h, const double* d_q, int64_t No, int64_t nrows, const float* tau_in, int nsp,
                             int64_t rows_per_split, int cap, uint2* buf, int* cnt, const KnnWindow& W) {
    size_t smem = (size_t)2 * KS_TM * ks_rowq(NF, FORM) * 16 + 2 * sizeof(uint64_t);
    FZB_CUDA(cudaFuncSetAttribute(k_knn_filter<NF, FORM, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((No + KS_T * R - 1) / (KS_T * R)), (unsigned)W.Kw, (unsigned)nsp);
    k_knn_filter<NF, FORM, R><<<grid, KS_T, smem, h->stream>>>(h->knn_scan.as<float>(), h->knn_Ns, nrows, d_q, h->knn_centre.as<double>(),
                                                               No, tau_in, rows_per_split, cap, buf, cnt, W.t0, W.tau_q, W.tau_fac);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}
// queries per thread: 4 in the difference form, 8 in the dot form (half the FMA work per distance: the shared-memory
// loads and the branch of a row are spread over more distances)
static int knn_filter_r(int form) { return form == KS_DOT ? KS_RDOT : KS_R; }
static int knn_filter_dispatch(fzb_context* h, const double* d_q, int64_t No, int64_t nrows, const float* tau_in, int nsp,
                               int64_t rows_per_split, int cap, uint2* buf, int* cnt, const KnnWindow& W) {
    const int nf = h->knn_Nf;
#define FZB_KF(NF_) \
    return h->knn_form == KS_DOT ? knn_filter_launch<NF_, KS_DOT, KS_RDOT>(h, d_q, No, nrows, tau_in, nsp, rows_per_split, cap, buf, cnt, W) \
                                 : knn_filter_launch<NF_, KS_DIFF, KS_R>(h, d_q, No, nrows, tau_in, nsp, rows_per_split, cap, buf, cnt, W)
    if (nf == 4) { FZB_KF(4); }
    if (nf == 5) { FZB_KF(5); }
    FZB_KF(6);
#undef FZB_KF
}

// row splits of a filter launch: fill the GPU and even out the last wave
static void knn_row_splits(fzb_context* h, int64_t ctas_per_split, int64_t nrows, bool split, int64_t* nsp_out, int64_t* rps_out) {
    const int64_t tiles = (nrows + KS_TM - 1) / KS_TM, slots = (int64_t)h->sm_count * 2;
    int64_t nsp = 1;
    if (split) {
        double best = 1e30;
        for (int64_t c = 1; c <= 32 && c <= tiles; ++c) {
            if ((tiles + c - 1) / c < 4 && c > 1) break;            // at least four tiles per split
            const double waves = (double)(ctas_per_split * c) / (double)slots;
            const double cost = std::ceil(waves) / waves * (1.0 + 0.01 * c);
            if (cost < best - 1e-9) { best = cost; nsp = c; }
        }
    }
    const int64_t rows_per_split = ((nrows + nsp - 1) / nsp + KS_TM - 1) / KS_TM * KS_TM;
    *nsp_out = (nrows + rows_per_split - 1) / rows_per_split;
    *rps_out = rows_per_split;
}

static int knn_select_launch(fzb_context* h, const uint2* buf, const int* cnt, int nsp, int cap, int64_t items, int KC, float* tau_out,
                             float* cand_d, int* cand_i, unsigned int* n_overflow, const KnnWindow& W) {
    const int wpb = 8;
    const size_t smem = (size_t)wpb * KS_CAP * sizeof(unsigned int);
    FZB_CUDA(cudaFuncSetAttribute(k_knn_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_knn_select<<<(unsigned)((items + wpb - 1) / wpb), wpb * 32, smem, h->stream>>>(buf, cnt, nsp, cap, items, KC, tau_out, cand_d, cand_i,
                                                                                 h->knn_m, n_overflow, W.Kw, W.t0, h->knn_K);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

// Candidate search of the trees of a window by nested prefixes (see above): leaves the KC smallest rows per (query, tree)
// in cand_d / cand_i and, if asked, the KC-th smallest distance in tau_final[query x tree of the window].
static int knn_staged_search(fzb_context* h, const double* d_q, int64_t No, int KC, uint2* buf, float* tau2, int* cnt,
                             float* cand_d, int* cand_i, unsigned int* n_overflow, const KnnWindow& W, float* tau_final) {
    const int64_t Ns = h->knn_Ns, items = No * W.Kw;
    // prefixes: KS_P0 rows, then a constant ratio up to all rows; a stage appends KC x ratio rows on average (relative
    // scatter 1 / sqrt(KC)), kept below 3/8 of the buffer
    double ratio_max = std::min(16.0, 0.375 * KS_CAP / KC);
    if (ratio_max < 2.0) ratio_max = 2.0;
    int J = 1;
    while (std::pow(ratio_max, J) * KS_P0 < (double)Ns) ++J;
    const double ratio = std::pow((double)Ns / KS_P0, 1.0 / J);
    const int64_t qper = (int64_t)KS_T * knn_filter_r(h->knn_form);
    const int64_t qtiles = (No + qper - 1) / qper;
    for (int j = 0; j <= J; ++j) {
        int64_t nrows = j == J ? Ns : (int64_t)std::ceil(KS_P0 * std::pow(ratio, j));
        if (nrows > Ns) nrows = Ns;
        int64_t nsp, rows_per_split;
        knn_row_splits(h, qtiles * W.Kw, nrows, j > 0, &nsp, &rows_per_split);
        const int cap = KS_CAP / (int)nsp;
        const float* tin = j == 0 ? nullptr : tau2 + (size_t)((j - 1) & 1) * items;
        float* tout = tau2 + (size_t)(j & 1) * items;
        if (knn_filter_dispatch(h, d_q, No, nrows, tin, (int)nsp, rows_per_split, cap, buf, cnt, W)) return 1;
        const bool last = j == J;
        if (knn_select_launch(h, buf, cnt, (int)nsp, cap, items, KC, last ? tau_final : tout, last ? cand_d : nullptr,
                              last ? cand_i : nullptr, n_overflow, W))
            return 1;
    }
    return 0;
}

// All trees of one query chunk.  The K trees are Monte-Carlo realisations of ONE training set (knn.py:175-188), so the
// k+8-th smallest distance of a query is nearly the same in all of them (C4: within +-25 %, 99.9 % below x1.32): the
// staged search runs for tree 0 only and 1.5 x its threshold filters the other trees in a single sweep (a quarter of the
// appended rows, no prefix stages).  A pair whose sweep finds fewer than k+8 rows (or overflows) is re-done by the
// float64 kernel like any failed exactness test; if that ever exceeds 1 % of a chunk the shortcut is switched off
// for the handle.  aux: tau2 = 2 x items floats, tauq = No floats.
static int knn_filter_search(fzb_context* h, const double* d_q, int64_t No, int KC, uint2* buf, float* tau2, float* tauq, int* cnt,
                             float* cand_d, int* cand_i, unsigned int* n_overflow, bool* shared) {
    const int K = h->knn_K;
    const bool share = K > 1 && !h->knn_share_off && getenv("FZB_KNN_NO_SHARE") == nullptr;
    *shared = share;
    if (!share) return knn_staged_search(h, d_q, No, KC, buf, tau2, cnt, cand_d, cand_i, n_overflow, KnnWindow{0, K, nullptr, 1.f}, nullptr);
    if (knn_staged_search(h, d_q, No, KC, buf, tau2, cnt, cand_d, cand_i, n_overflow, KnnWindow{0, 1, nullptr, 1.f}, tauq)) return 1;
    const char* e = getenv("FZB_KNN_SHARE_FAC");
    const KnnWindow W{1, K - 1, tauq, e ? (float)atof(e) : 1.5f};
    const int64_t qper = (int64_t)KS_T * knn_filter_r(h->knn_form);
    int64_t nsp, rows_per_split;
    knn_row_splits(h, ((No + qper - 1) / qper) * W.Kw, h->knn_Ns, true, &nsp, &rows_per_split);
    const int cap = KS_CAP / (int)nsp;
    if (knn_filter_dispatch(h, d_q, No, h->knn_Ns, nullptr, (int)nsp, rows_per_split, cap, buf, cnt, W)) return 1;
    return knn_select_launch(h, buf, cnt, (int)nsp, cap, No * W.Kw, KC, nullptr, cand_d, cand_i, n_overflow, W);
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
    const bool fast = use_tc || (pmode == 2 && h->knn_scan_valid && KC <= KS_KCMAX && getenv("FZB_KNN_EXACT_ONLY") == nullptr);
    if (!fast) return knn_exact_launch(h, d_q, No, k, p, pmode, d_idx, d_dist, nullptr, nullptr, No * K);

    // tensor-core scan: row splits so that small query batches still fill the GPU
    int64_t nsp = 1, rows_per_split = Nm;
    const int64_t rtile = 256;
    if (use_tc) {
        const int64_t qtiles = (No + 127) / 128;
        const int64_t want = (int64_t)h->sm_count * 2;
        nsp = (want + qtiles * K - 1) / (qtiles * K);
        const int64_t max_sp = (Nm + 4 * rtile - 1) / (4 * rtile);
        if (nsp > max_sp) nsp = max_sp;
        if (nsp > 32) nsp = 32;
        if (nsp < 1) nsp = 1;
        rows_per_split = ((Nm + nsp - 1) / nsp + rtile - 1) / rtile * rtile;
        nsp = (Nm + rows_per_split - 1) / rows_per_split;
    }
    // process the queries in chunks that bound the candidate buffers (tensor-core scan: ~2 GB of lists; filter scan:
    // KS_CAP buffered rows per (query, tree), up to 12 GB)
    const int nlist = use_tc ? (int)nsp * fzb_knn_tc_lists() : 1;     // candidate lists per (query, tree)
    const size_t per_q = (size_t)K * nlist * KC * 8;
    const size_t per_q_buf = use_tc ? 0 : (size_t)K * (KS_CAP * 8 + 8 + 32 * 4) + 4;
    int64_t chunk = use_tc ? (int64_t)(((size_t)2 << 30) / per_q) : (int64_t)(((size_t)12 << 30) / per_q_buf);
    const int64_t qtile = use_tc ? 128 : (int64_t)KS_T * knn_filter_r(h->knn_form);
    chunk = chunk / qtile * qtile;        // whole query tiles
    if (chunk < qtile) chunk = qtile;
    if (chunk > No) chunk = No;
    if (h->knn_cand.reserve((size_t)chunk * per_q + 256) || h->knn_redo.reserve((size_t)chunk * K * 8 + 64)) return 1;
    if (!use_tc && h->knn_buf.reserve((size_t)chunk * per_q_buf + 256)) return 1;
    float* cand_d = h->knn_cand.as<float>();
    int* cand_i = reinterpret_cast<int*>(cand_d + (size_t)chunk * K * nlist * KC);
    int* n_redo = h->knn_redo.as<int>();
    int64_t* redo = reinterpret_cast<int64_t*>(n_redo + 4);
    uint2* fbuf = h->knn_buf.as<uint2>();
    float* ftau = reinterpret_cast<float*>(fbuf + (size_t)chunk * K * KS_CAP);
    float* ftauq = ftau + (size_t)2 * chunk * K;
    int* fcnt = reinterpret_cast<int*>(ftauq + chunk);
    double ms_search = 0.0;
    bool shared = false;
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        int64_t nc = No - o0 < chunk ? No - o0 : chunk;
        const double* qq = d_q + o0 * nf;
        FZB_CUDA(cudaEventRecord(h->ev[2], h->stream));
        FZB_CUDA(cudaMemsetAsync(n_redo, 0, 16, h->stream));
        int rc = use_tc ? fzb_knn_tc_scan(h, qq, nc, KC, (int)nsp, (int)(rows_per_split / rtile), cand_d, cand_i)
                        : knn_filter_search(h, qq, nc, KC, fbuf, ftau, ftauq, fcnt, cand_d, cand_i,
                                            reinterpret_cast<unsigned int*>(n_redo + 1), &shared);
        if (rc) return rc;
        const int wpb = 4;
        size_t smem = (size_t)wpb * KC * nlist * 16;
        FZB_CHECK(smem <= 200 * 1024, "kNN re-rank: too many candidates");
        FZB_CUDA(cudaFuncSetAttribute(k_knn_rerank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t items = nc * K;
        k_knn_rerank<<<(unsigned)((items + wpb - 1) / wpb), wpb * 32, smem, h->stream>>>(
            h->knn_feats.as<float>(), h->knn_stride, K, Nm, nf, qq, nc, k, KC, nlist, cand_d, cand_i, d_idx + (size_t)o0 * K * k,
            d_dist ? d_dist + (size_t)o0 * K * k : nullptr, redo, n_redo, use_tc ? 1 : h->knn_form,
            use_tc ? h->knn_aux.as<double>() : h->knn_centre.as<double>(), use_tc ? env_c0() : KS_C0,
            reinterpret_cast<unsigned long long*>(n_redo + 2));
        fzb_count_launch(h);
        FZB_CUDA(cudaGetLastError());
        // the rare (query, tree) pairs that failed the exactness test are re-done by the float64 kernel
        if (knn_exact_launch(h, qq, nc, k, p, pmode, d_idx + (size_t)o0 * K * k,
                             d_dist ? d_dist + (size_t)o0 * K * k : nullptr, redo, n_redo, 4096))
            return 1;
        FZB_CUDA(cudaEventRecord(h->ev[3], h->stream));
        int nr[4] = {0, 0, 0, 0};
        FZB_CUDA(cudaMemcpyAsync(nr, n_redo, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        FZB_CUDA(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
        ms_search += ms;
        h->stats.knn_redo += nr[0];
        if (shared && (int64_t)nr[0] * 100 > nc * K) h->knn_share_off = true;      // the trees are not alike: staged search for all
        {   // largest error of the candidates' fp32 values in units of the bound's scale (tensor-core / dot forms)
            double w;
            memcpy(&w, nr + 2, 8);
            if (w > h->stats.knn_tc_err) h->stats.knn_tc_err = w;
        }
        h->stats.knn_overflow += nr[1];
        h->stats.knn_tc = use_tc ? 1 : 0;
    }
    h->stats.ms_scan += ms_search;       // CUDA-event time of the search alone (candidate scan, select, re-rank, re-dos)
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
