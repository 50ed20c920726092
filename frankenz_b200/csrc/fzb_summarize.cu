// PDF summary statistics on the device: frankenz/pdf.py:899-1074 (`pdfs_summarize`), SURVEY.md section 8f rank 2.
//
// Per object: row sum in numpy's pairwise order and in-place style renormalisation (pdf.py:980), mean (:983), mode
// (:986), sequential CDF (:989, numpy's cumsum order, so that plateaus compare equal exactly as they do there),
// quantiles / median / Monte-Carlo draw by inverse-CDF interpolation (:994-997, numpy.interp semantics), the risk
// curve risk[g] = sum_t pdf[t] (1 - kernel[t, g]) (:1024, a (No x Ng) x (Ng x Ng) float64 GEMM: the bulk of the work)
// and its arg-min `best` (:1025), the standard deviation about each of the four estimators (:1028-1036), the risk at
// each of them (:1066-1068); stage 2 adds the probability within +-width of each estimator (:1038-1062) once the
// caller has evaluated its `wconf_func` at them.  This file is compiled without FMA contraction (numpy.interp is
// plain C); the GEMM uses explicit fma().
//
// k_summarize: one CTA = 32 objects x 384 threads (16 objects for grids above 704 points).  The PDF tile lives in shared
// memory TRANSPOSED ([grid point][object], 16-byte object pairs XOR-swizzled by the grid point so that the transposing
// store, the lane = object readers and the DMMA fragment loads are all (nearly) conflict free).  Warps 0-10 own the risk
// product on the FP64 tensor cores (mma.sync.m8n8k4.f64, measured 37 TFLOP/s on B200 against 34 for DFMA, tools/fp64_rate.cu;
// what decides is that a DMMA needs 1/8 of the issue slots and 1/6 of the shared-memory loads per FMA): warp w holds the
// NOBJ x 64 accumulator block of grid columns [64 w, 64 w + 64) in registers; four loss rows per step arrive by bulk-TMA
// into one half of an 8-row shared-memory ring (rows skewed by 64 bytes: fragment loads without bank conflicts), and the
// LAST warp to have its fragments of a half in registers requests the refill at once, two steps ahead of its use.
// 32 objects per CTA is what takes the kernel off the L2 roof: every CTA streams the whole (Ng x Ng) loss matrix, 3.9 MB
// at Ng = 701, i.e. 123 GB of L2 -> SM traffic per 1M objects instead of 490 GB at 8 objects (round 1, 88 ms per 1M).
// Warp 11 meanwhile runs the one inherently sequential piece with lane = object: the numpy-order CDF, kept as one
// checkpoint per 64 grid points (a later look-up re-adds at most 64 terms in the same order, so it is bit-identical).
// After the GEMM, with the FP64 pipe free again: mean and mode (grid slices per warp), the quantiles (numpy.interp with
// xp = CDF needs only c[j], c[j + 1] around the last c[j] <= q: found behind the last checkpoint <= q), the risk rows
// through the ring's memory 8 objects at a time (arg-min, interpolation at the four estimators), the standard deviations
// and the confidence look-ups.  numpy's pairwise row sum runs leaf-parallel (its <= 128-element leaves are enumerated on
// the host, the combination tree is the same recursion).
#include <algorithm>
#include <cfloat>

#include "fzb_common.cuh"

namespace {

constexpr int GW = 11;                 // GEMM warps: with the scan warp 12 = three register-allocation granules of 4 warps
constexpr int ST2 = (GW + 1) * 32;     // + the scan / producer warp: 384 threads, 168 registers each
constexpr int NGMAX = GW * 8 * 12;     // 1056 grid points: 12 column blocks of 8 per GEMM warp (8 blocks = 704 >= the 701-point grid)
constexpr int CKS = 64;                // CDF checkpoint spacing
constexpr int RO = 8;                  // risk rows per round through the ring's memory = one DMMA row block

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sm_bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sm_bar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sm_bar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sm_u32(bar)) : "memory");
}
__device__ __forceinline__ void sm_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm_u32(dst)),
                 "l"(src), "r"(bytes), "r"(sm_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sm_bar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nSW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra SD_%=;\nbra SW_%=;\nSD_%=:\n}\n" ::"r"(
            sm_u32(bar)),
        "r"(parity)
        : "memory");
}
// element (grid point t, object o) of the transposed, pair-swizzled PDF tile
template <int NOBJ>
__device__ __forceinline__ int pidx(int t, int o) {
    const int sw = (((t & 3) << 2) | ((t >> 2) & 3)) & (NOBJ / 2 - 1);     // 4 consecutive grid points -> 4 distinct 64-byte windows
    return t * NOBJ + ((((o >> 1) ^ sw) << 1) | (o & 1));
}

// numpy's pairwise summation of a contiguous float64 row (numpy/core/src/umath/loops_utils.h, DOUBLE_pairwise_sum); the
// row is object o of the tile, elements [base, base + n)
template <int NOBJ>
__device__ double np_pairwise_sum(const double* tile, int o, int base, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += tile[pidx<NOBJ>(base + i, o)];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = tile[pidx<NOBJ>(base + j, o)];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += tile[pidx<NOBJ>(base + i + j, o)];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += tile[pidx<NOBJ>(base + i, o)];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const double left = np_pairwise_sum<NOBJ>(tile, o, base, n2);          // depth <= 4 for the supported grids (n <= 1024)
    const double right = np_pairwise_sum<NOBJ>(tile, o, base + n2, n - n2);
    return left + right;
}

// the combination tree of the same recursion over precomputed leaf sums ls[0], ls[stride], ... (in leaf order)
__device__ double np_pairwise_combine(int n, const double* ls, int stride, int& idx) {
    if (n <= 128) return ls[(idx++) * stride];
    int n2 = n / 2;
    n2 -= n2 % 8;
    const double left = np_pairwise_combine(n2, ls, stride, idx);
    const double right = np_pairwise_combine(n - n2, ls, stride, idx);
    return left + right;
}

// numpy.interp(x, xp, fp) for one point, xp non-decreasing (numpy/core/src/multiarray/compiled_base.c, arr_interp):
// j = last index with xp[j] <= x; exact hits return fp[j]; the slope is formed per point
__device__ __forceinline__ double np_interp_finish(double x, double xj, double xn, double fj, double fn) {
    const double slope = (fn - fj) / (xn - xj);
    double r = slope * (x - xj) + fj;
    if (r != r) {
        r = slope * (x - xn) + fn;
        if (r != r && fj == fn) r = fj;
    }
    return r;
}

__device__ int np_search(double x, const double* xp, int n) {      // last index with xp[j] <= x (n - 1 >= j >= -1)
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

__device__ double np_interp(double x, const double* xp, const double* fp, int n) {
    if (x != x) return x;
    if (x > xp[n - 1]) return fp[n - 1];
    if (x < xp[0]) return fp[0];
    const int j = np_search(x, xp, n);
    if (j == n - 1) return fp[j];
    if (xp[j] == x) return fp[j];
    return np_interp_finish(x, xp[j], xp[j + 1], fp[j], fp[j + 1]);
}

struct SummParams {
    const double* pdfs;     // device (No x Ng) rows of this launch
    const double* pgrid;    // [Ng]
    const double* loss;     // (4 ceil(Ng / 4) x LP), LP = 88 NNB: 1 - kernel[truth, guess], padding zero, rows 16-byte aligned
    const double* urand;    // [No]
    int64_t No;
    int Ng, LP, renorm;
    double* rowsum;         // [No]
    double* cdf;            // (No x Ng) out, kept for stage 2 (nullable)
    double *est, *sd, *risk, *quant, *mc;   // [4][Ntot] each (mc: [Ntot]); column offset o0
    double* conf;           // nullable [4][Ntot]: probability within +-wfac (1 + estimator), the default `wconf_func`
    double wfac;            // (pdf.py:1038-1062), formed from the CDF checkpoints in shared memory
    int64_t Ntot, o0;
    int nleaf, leaf_base[16], leaf_n[16];   // leaves of numpy's pairwise summation of a row of Ng elements
    int ncopy;              // copies of the loss matrix, `copy_stride` doubles apart: CTAs marching through ONE copy in
    size_t copy_stride;     // lockstep would all hit the same few L2 slices at the same time
};

// shared-memory layout in doubles: tile (4 ceil(Ng / 4) rows), ring of 8 loss rows with a 64-byte skew per row (>= the risk
// rows and the sd partials that reuse it), checkpoints, 8 per-object rows, then ints and barriers
__host__ __device__ inline size_t summ_ring_doubles(int Ng, int LP, int nobj) {
    size_t ring = (size_t)8 * (LP + 8);
    const size_t part = (size_t)GW * 4 * nobj, rsk = (size_t)RO * Ng;
    if (ring < part) ring = part;
    if (ring < rsk) ring = rsk;
    return ring;
}
__host__ __device__ inline size_t summ_smem_bytes(int Ng, int LP, int nobj) {
    return ((size_t)((Ng + 3) / 4 * 4) * nobj + summ_ring_doubles(Ng, LP, nobj) + (size_t)((Ng + CKS - 1) / CKS) * nobj +
            (size_t)8 * nobj) * 8 + (size_t)2 * nobj * 4 + 4 * 8;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// sequential (numpy.cumsum order) CDF value c[j] of object o from the checkpoints
template <int NOBJ>
__device__ double cdf_at(const double* tile, const double* ckpt, int o, int j) {
    const int k = j / CKS;
    double run = ckpt[k * NOBJ + o];
    for (int t = k * CKS + 1; t <= j; ++t) run = run + tile[pidx<NOBJ>(t, o)];
    return run;
}

template <int NNB, int NOBJ>
__global__ void __launch_bounds__(ST2, 1) k_summarize(SummParams P) {
    extern __shared__ __align__(128) double sm_s[];
    constexpr int NMB = NOBJ / 8;                               // DMMA row blocks (8 objects each)
    const int Ng = P.Ng, LP = P.LP, RP = LP + 8, NK = (Ng + 3) / 4;
    double* spdf = sm_s;                                        // [4 NK][NOBJ], swizzled; rows >= Ng zero
    double* ring = spdf + (size_t)NK * 4 * NOBJ;                // [2 halves][4 rows][RP]; later risk rows [8][Ng], then sd partials
    double* ckpt = ring + summ_ring_doubles(Ng, LP, NOBJ);      // [ceil(Ng / CKS)][NOBJ]
    double* s_mean = ckpt + (size_t)((Ng + CKS - 1) / CKS) * NOBJ;   // [NOBJ]
    double* s_med = s_mean + NOBJ;
    double* s_sum = s_med + NOBJ;
    double* s_pts = s_sum + NOBJ;                               // [4][NOBJ]
    double* s_clast = s_pts + 4 * NOBJ;                         // [NOBJ]: c[Ng - 1]
    int* s_imax = reinterpret_cast<int*>(s_pts + 5 * NOBJ);     // [NOBJ] (+ NOBJ spare)
    int* s_cnt = s_imax + NOBJ;                                 // [2]: GEMM warps that hold their fragments of the half
    uint64_t* full = reinterpret_cast<uint64_t*>(s_imax + 2 * NOBJ);   // [2]: the four rows of a half have landed
    uint64_t* empty = full + 2;                                        // [2]: every GEMM warp holds its fragments of the half
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ob = (int64_t)blockIdx.x * NOBJ;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            sm_bar_init(&full[s], 1);
            sm_bar_init(&empty[s], GW * 32);
            s_cnt[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t row_bytes = (uint32_t)LP * 8;
    const double* lossc = P.loss + (size_t)(blockIdx.x % (unsigned)P.ncopy) * P.copy_stride;
    auto issue = [&](int k) {                                   // loss rows 4k .. 4k + 3 into half k & 1
        const int h = k & 1;
        sm_bar_expect(&full[h], 4 * row_bytes);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            sm_bulk_g2s(ring + (size_t)(h * 4 + i) * RP, lossc + (size_t)(4 * k + i) * LP, row_bytes, &full[h]);
    };
    if (tid == GW * 32)
        for (int k = 0; k < 2 && k < NK; ++k) issue(k);
    for (int i = tid; i < NOBJ * (NK * 4 - Ng); i += ST2) spdf[(size_t)Ng * NOBJ + i] = 0.0;
    {   // up to three object rows per warp, lanes along the grid (coalesced), the rows interleaved for loads in flight
        constexpr int NR = (NOBJ + GW) / (GW + 1);
        const double* src[NR];
        bool live[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int o = warp + r * (GW + 1);
            live[r] = o < NOBJ && ob + o < P.No;
            src[r] = P.pdfs + (size_t)(ob + (live[r] ? o : 0)) * Ng;
        }
#pragma unroll 4
        for (int t = lane; t < Ng; t += 32) {
            double v[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) v[r] = live[r] ? src[r][t] : 0.0;
#pragma unroll
            for (int r = 0; r < NR; ++r)
                if (warp + r * (GW + 1) < NOBJ) spdf[pidx<NOBJ>(t, warp + r * (GW + 1))] = v[r];
        }
    }
    __syncthreads();
    if (P.renorm) {
        // numpy's pairwise sum, its leaves (<= 128 elements each, host-enumerated) in parallel, lane = object
        for (int i = tid; i < P.nleaf * NOBJ; i += ST2) {
            const int L = i / NOBJ, o = i - L * NOBJ;
            ckpt[i] = np_pairwise_sum<NOBJ>(spdf, o, P.leaf_base[L], P.leaf_n[L]);
        }
        __syncthreads();
        if (tid < NOBJ) {
            int idx = 0;
            const double s = np_pairwise_combine(Ng, ckpt + tid, NOBJ, idx);
            s_sum[tid] = s;
            if (ob + tid < P.No && P.rowsum) P.rowsum[P.o0 + ob + tid] = s;
        }
        __syncthreads();
        {   // lane = object within a row of the tile: the divisor stays in a register, four divisions in flight
            const int o = tid % NOBJ;
            const double sdiv = s_sum[o];
#pragma unroll 4
            for (int t = tid / NOBJ; t < Ng; t += ST2 / NOBJ) {
                const int a = pidx<NOBJ>(t, o);
                spdf[a] = spdf[a] / sdiv;
            }
        }
        __syncthreads();
    }

    double acc[NMB][NNB][2];
    if (warp < GW) {
        // ---- risk curves on the FP64 tensor cores: C[o][g] += pdf[o][t] loss[t][g], 8 x 8 x 4 per DMMA; this warp owns
        // columns [warp 8 NNB, (warp + 1) 8 NNB) of all NOBJ objects.  Fragments (PTX ISA, m8n8k4 .f64): A[lane / 4][lane % 4],
        // B[lane % 4][lane / 4], C[lane / 4][2 (lane % 4) + {0, 1}].  The half is refilled as soon as the last warp has its
        // fragments in registers, i.e. before the DMMAs that use them: the rows have two whole steps to arrive.
#pragma unroll
        for (int mb = 0; mb < NMB; ++mb)
#pragma unroll
            for (int nb = 0; nb < NNB; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
        const int lr = lane & 3, lq = lane >> 2;
        const int colbase = warp * 8 * NNB + lq;
        double af[NMB];
#pragma unroll
        for (int mb = 0; mb < NMB; ++mb) af[mb] = spdf[pidx<NOBJ>(lr, mb * 8 + lq)];
        for (int j = 0; j < NK; ++j) {
            const int h = j & 1;
            sm_bar_wait(&full[h], (uint32_t)(j >> 1) & 1u);
            const double* rb = ring + (size_t)(h * 4 + lr) * RP + colbase;
            double bf[NNB];
#pragma unroll
            for (int nb = 0; nb < NNB; ++nb) bf[nb] = rb[nb * 8];
            sm_bar_arrive(&empty[h]);                           // every lane: its own reads of the half are done
            __syncwarp();
            if (lane == 0) {
                if (atomicAdd(&s_cnt[h], 1) == GW - 1) {        // the last warp to hold its fragments refills the half:
                    s_cnt[h] = 0;                               // every other warp arrived before it counted, so the
                    if (j + 2 < NK) {                           // wait below returns at once (it is the acquire)
                        sm_bar_wait(&empty[h], (uint32_t)(j >> 1) & 1u);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue(j + 2);
                    }
                }
            }
            // the PDF fragments of the next step do not wait for anything: they are loaded behind this step's DMMAs
            const int tn = min(4 * (j + 1), 4 * (NK - 1)) + lr;
#pragma unroll
            for (int mb = 0; mb < NMB; ++mb) {
#pragma unroll
                for (int nb = 0; nb < NNB; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], af[mb], bf[nb]);
                af[mb] = spdf[pidx<NOBJ>(tn, mb * 8 + lq)];
            }
        }
    } else if (lane < NOBJ) {
        // ---- the one sequential piece, lane = object: the CDF in numpy.cumsum order, one checkpoint per CKS grid points
        // (and the CDF rows themselves when stage 2 wants them).  The FP64 pipe is saturated by the DMMAs meanwhile, so
        // nothing else runs here.
        const int o = lane;
        const int64_t og = ob + o;
        double run = 0.0;
        double* crow = (P.cdf && og < P.No) ? P.cdf + (size_t)(P.o0 + og) * Ng : nullptr;
        for (int t = 0; t < Ng; ++t) {
            const double p = spdf[pidx<NOBJ>(t, o)];
            run = (t == 0) ? p : run + p;
            if ((t & (CKS - 1)) == 0) ckpt[(t / CKS) * NOBJ + o] = run;
            if (crow) crow[t] = run;
        }
        s_clast[o] = run;
    }
    __syncthreads();

    // ---- mean and mode: grid slices per warp, lane = object; partials through the ring's memory --------------------
    {
        double* pm = ring;                                      // [GW][NOBJ] partial means
        double* px = pm + GW * NOBJ;                            // [GW][NOBJ] slice maxima
        int* pi = reinterpret_cast<int*>(px + GW * NOBJ);       // [GW][NOBJ] their first positions
        if (warp < GW && lane < NOBJ) {
            const int o = lane;
            const int ts = (Ng + GW - 1) / GW;
            const int t1 = min(Ng, (warp + 1) * ts);
            double mean = 0.0, pmax = -DBL_MAX;
            int imax = 0x7fffffff;
            for (int t = warp * ts; t < t1; ++t) {
                const double p = spdf[pidx<NOBJ>(t, o)];
                mean = fma(p, P.pgrid[t], mean);
                if (p > pmax) { pmax = p; imax = t; }
            }
            pm[warp * NOBJ + o] = mean;
            px[warp * NOBJ + o] = pmax;
            pi[warp * NOBJ + o] = imax;
        }
        __syncthreads();
        if (tid < NOBJ) {
            double mean = 0.0, pmax = -DBL_MAX;
            int imax = 0x7fffffff;
            for (int w = 0; w < GW; ++w) {
                mean += pm[w * NOBJ + tid];
                if (px[w * NOBJ + tid] > pmax) { pmax = px[w * NOBJ + tid]; imax = pi[w * NOBJ + tid]; }   // first maximum
            }
            s_mean[tid] = mean;
            s_imax[tid] = (imax == 0x7fffffff) ? 0 : imax;     // all-NaN row: numpy returns the first NaN; not reproduced
        }
    }
    // ---- quantiles, median, Monte-Carlo draw: numpy.interp with xp = CDF needs j = the last index with c[j] <= q and
    // c[j], c[j + 1]; the checkpoints are monotone, so j lies behind the last checkpoint <= q and is found by re-adding at
    // most CKS terms in the order of the cumulative sum
    if (tid < 6 * NOBJ) {
        const int o = tid % NOBJ, e = tid / NOBJ;
        const int64_t og = ob + o;
        const double q = (e == 5) ? ((og < P.No) ? P.urand[P.o0 + og] : 0.5)
                                  : (e == 0 ? 0.025 : e == 1 ? 0.16 : e == 2 ? 0.5 : e == 3 ? 0.84 : 0.975);
        const int nck = (Ng + CKS - 1) / CKS;
        double r;
        if (q != q) r = q;
        else if (q > s_clast[o]) r = P.pgrid[Ng - 1];
        else if (q < ckpt[o]) r = P.pgrid[0];
        else {
            int k = 0;
            while (k + 1 < nck && ckpt[(k + 1) * NOBJ + o] <= q) ++k;
            int j = k * CKS;
            double cj = ckpt[k * NOBJ + o], cn = cj, c = cj;
            for (int t = j + 1; t < Ng; ++t) {
                c = c + spdf[pidx<NOBJ>(t, o)];
                if (c <= q) { j = t; cj = c; }
                else { cn = c; break; }
            }
            if (j == Ng - 1 || cj == q) r = P.pgrid[j];
            else r = np_interp_finish(q, cj, cn, P.pgrid[j], P.pgrid[j + 1]);
        }
        if (e == 2) s_med[o] = r;
        if (og < P.No) {
            const int64_t qo = P.o0 + og;
            if (e == 5) P.mc[qo] = r;
            else if (e != 2) P.quant[(size_t)(e < 2 ? e : e - 1) * P.Ntot + qo] = r;
        }
    }
    __syncthreads();

    // ---- risk rows through the ring's memory, 8 objects at a time: arg-min, interpolation at the estimators ------
    double* srsk = ring;
#pragma unroll
    for (int r = 0; r < NMB; ++r) {
        if (warp < GW) {
            const int lr = lane & 3, lq = lane >> 2;
#pragma unroll
            for (int nb = 0; nb < NNB; ++nb) {
                const int col = warp * 8 * NNB + nb * 8 + 2 * lr;
                if (col < Ng) srsk[(size_t)lq * Ng + col] = acc[r][nb][0];
                if (col + 1 < Ng) srsk[(size_t)lq * Ng + col + 1] = acc[r][nb][1];
            }
        }
        __syncthreads();
        if (warp < RO) {
            const int o = RO * r + warp;
            const int64_t og = ob + o;
            const double* rr = srsk + (size_t)warp * Ng;
            double rmin = DBL_MAX;
            int imin = 0x7fffffff;
            for (int t = lane; t < Ng; t += 32)
                if (rr[t] < rmin) { rmin = rr[t]; imin = t; }
            for (int s = 16; s > 0; s >>= 1) {
                const double orr = __shfl_xor_sync(0xffffffffu, rmin, s);
                const int oj = __shfl_xor_sync(0xffffffffu, imin, s);
                if (orr < rmin || (orr == rmin && oj < imin)) { rmin = orr; imin = oj; }   // first minimum (numpy.argmin)
            }
            if (imin == 0x7fffffff) imin = 0;
            const double pts[4] = {s_mean[o], s_med[o], P.pgrid[s_imax[o]], P.pgrid[imin]};
            if (lane < 4) {
                s_pts[lane * NOBJ + o] = pts[lane];
                if (og < P.No) {
                    const int64_t q = P.o0 + og;
                    P.est[(size_t)lane * P.Ntot + q] = pts[lane];
                    P.risk[(size_t)lane * P.Ntot + q] = np_interp(pts[lane], P.pgrid, rr, Ng);
                }
            }
        }
        __syncthreads();
    }

    // ---- standard deviation about each estimator: grid slices per warp, lane = object ------------------------------
    double* part = ring;                       // [GW][4][NOBJ]
    if (warp < GW && lane < NOBJ) {
        const int o = lane;
        const int ts = (Ng + GW - 1) / GW;
        const int t1 = min(Ng, (warp + 1) * ts);
        const double pts[4] = {s_pts[o], s_pts[NOBJ + o], s_pts[2 * NOBJ + o], s_pts[3 * NOBJ + o]};
        double sd[4] = {0.0, 0.0, 0.0, 0.0};
        for (int t = warp * ts; t < t1; ++t) {
            const double g = P.pgrid[t];
            const double p = spdf[pidx<NOBJ>(t, o)];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double d = g - pts[e];
                sd[e] = fma(d * d, p, sd[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) part[(size_t)(warp * 4 + e) * NOBJ + o] = sd[e];
    }
    __syncthreads();
    if (tid < 4 * NOBJ) {
        const int e = tid / NOBJ, o = tid - e * NOBJ;
        double s = 0.0;
        for (int w = 0; w < GW; ++w) s += part[(size_t)(w * 4 + e) * NOBJ + o];
        if (ob + o < P.No) P.sd[(size_t)e * P.Ntot + P.o0 + ob + o] = sqrt(s);
    }
    // ---- probability within +-width of each estimator (numpy.interp with xp = grid, fp = CDF) ----------------------
    if (P.conf && tid < 8 * NOBJ) {
        const int o = tid >> 3, e = (tid >> 1) & 3, side = tid & 1;
        const double pt = s_pts[e * NOBJ + o];
        const double w = (1. + pt) * P.wfac;
        const double x = side ? pt + w : pt - w;
        double v;
        if (x != x) v = x;
        else if (x > P.pgrid[Ng - 1]) v = cdf_at<NOBJ>(spdf, ckpt, o, Ng - 1);
        else if (x < P.pgrid[0]) v = cdf_at<NOBJ>(spdf, ckpt, o, 0);
        else {
            const int j = np_search(x, P.pgrid, Ng);
            const double cj = cdf_at<NOBJ>(spdf, ckpt, o, j);
            if (j == Ng - 1 || P.pgrid[j] == x) v = cj;
            else {
                const double cn = cj + spdf[pidx<NOBJ>(j + 1, o)];
                v = np_interp_finish(x, P.pgrid[j], P.pgrid[j + 1], cj, cn);
            }
        }
        const double other = __shfl_xor_sync(0xffffffffu, v, 1);
        if (side == 0 && ob + o < P.No) P.conf[(size_t)e * P.Ntot + P.o0 + ob + o] = other - v;
    }
}

struct ConfParams {
    const double* pgrid;
    const double* cdf;      // (No x Ng)
    const double *points, *widths;   // [4][No]
    double* conf;           // [4][No]
    int64_t No;
    int Ng;
};

__global__ void k_conf(ConfParams P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * P.No) return;
    const int64_t o = i % P.No;
    const double pt = P.points[i], w = P.widths[i];
    const double* c = P.cdf + (size_t)o * P.Ng;
    const double lo = np_interp(pt - w, P.pgrid, c, P.Ng);
    const double hi = np_interp(pt + w, P.pgrid, c, P.Ng);
    P.conf[i] = hi - lo;
}

template <int NNB, int NOBJ>
int launch_summ_t(fzb_context* h, const SummParams& P, int64_t nobj) {
    const size_t smem = summ_smem_bytes(P.Ng, P.LP, NOBJ);
    FZB_CUDA(cudaFuncSetAttribute(k_summarize<NNB, NOBJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_summarize<NNB, NOBJ><<<(unsigned)((nobj + NOBJ - 1) / NOBJ), ST2, smem, h->stream>>>(P);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

// column blocks per GEMM warp for a grid of Ng points (instantiated: 2, 4, 8 with 32 objects per CTA; 12 with 16)
int summ_nnb(int Ng) { return Ng <= 176 ? 2 : Ng <= 352 ? 4 : Ng <= 704 ? 8 : 12; }
int summ_lp(int Ng) { return GW * 8 * summ_nnb(Ng); }
int summ_copies() {
    const char* e = getenv("FZB_SUMM_COPIES");
    const int c = e ? atoi(e) : 1;      // measured: 1, 4, 8, 16 copies run alike on B200 (no L2 slice hot-spotting)
    return c < 1 ? 1 : c > 32 ? 32 : c;
}
void summ_leaves(int base, int n, SummParams& P) {
    if (n <= 128) {
        P.leaf_base[P.nleaf] = base;
        P.leaf_n[P.nleaf++] = n;
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    summ_leaves(base, n2, P);
    summ_leaves(base + n2, n - n2, P);
}
size_t summ_copy_stride(int Ng) { return (((size_t)((Ng + 3) / 4 * 4) * summ_lp(Ng) * 8 + 255) / 256 * 256) / 8; }

int launch_summ(fzb_context* h, const SummParams& P, int64_t nobj) {
    switch (summ_nnb(P.Ng)) {
        case 2: return launch_summ_t<2, 32>(h, P, nobj);
        case 4: return launch_summ_t<4, 32>(h, P, nobj);
        case 8: return launch_summ_t<8, 32>(h, P, nobj);
        default: return launch_summ_t<12, 16>(h, P, nobj);
    }
}

// the loss matrix padded to the GEMM's shape (bulk-copy sources: rows 16-byte aligned), padding zeroed
int upload_loss(fzb_context* h, DevBuf& d_loss, const double* loss, int Ng) {
    const int LP = summ_lp(Ng), rows = (Ng + 3) / 4 * 4;
    const int ncopy = summ_copies();
    const size_t one = ((size_t)rows * LP * 8 + 255) / 256 * 256;
    if (d_loss.reserve(one * ncopy)) return 1;
    FZB_CUDA(cudaMemsetAsync(d_loss.p, 0, one, h->stream));
    FZB_CUDA(cudaMemcpy2DAsync(d_loss.p, (size_t)LP * 8, loss, (size_t)Ng * 8, (size_t)Ng * 8, (size_t)Ng,
                               cudaMemcpyHostToDevice, h->stream));
    for (int c = 1; c < ncopy; ++c)
        FZB_CUDA(cudaMemcpyAsync(static_cast<char*>(d_loss.p) + c * one, d_loss.p, one, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

}  // namespace

int fzb_summarize_impl(fzb_context* h, const double* pdfs, const double* pgrid, const double* loss, const double* urand,
                       int64_t No, int32_t Ng, int32_t renormalize, double* rowsum, double* est, double* sd, double* risk,
                       double* quant, double* mc) {
    FZB_CHECK(Ng >= 2 && Ng <= NGMAX, "pdfs_summarize: grid of %d points (supported: 2..%d)", Ng, NGMAX);
    h->stats = FzbStats{};
    DevBuf& d_grid = h->summ[0];
    DevBuf& d_loss = h->summ[1];
    DevBuf& d_cdf = h->summ[2];
    DevBuf& d_out = h->summ[3];
    DevBuf& d_in = h->summ[4];
    DevBuf& d_u = h->summ[5];
    if (d_grid.reserve((size_t)Ng * 8) || upload_loss(h, d_loss, loss, Ng) || d_cdf.reserve((size_t)No * Ng * 8 + 64) ||
        d_out.reserve((size_t)No * 18 * 8 + 64) || d_u.reserve((size_t)No * 8 + 64))
        return 1;
    FZB_CUDA(cudaMemcpyAsync(d_grid.p, pgrid, (size_t)Ng * 8, cudaMemcpyHostToDevice, h->stream));
    FZB_CUDA(cudaMemcpyAsync(d_u.p, urand, (size_t)No * 8, cudaMemcpyHostToDevice, h->stream));
    double* o_est = d_out.as<double>();
    double* o_sd = o_est + 4 * No;
    double* o_risk = o_sd + 4 * No;
    double* o_quant = o_risk + 4 * No;
    double* o_mc = o_quant + 4 * No;
    double* o_sum = o_mc + No;
    int64_t chunk = std::max<int64_t>(32, std::min<int64_t>(No, ((int64_t)256 << 20) / ((int64_t)Ng * 8) / 32 * 32));
    const int64_t wave = (int64_t)std::max(1, h->sm_count) * 32;          // whole waves of one 32-object CTA per SM
    if (chunk > wave) chunk = chunk / wave * wave;
    if (d_in.reserve((size_t)chunk * Ng * 8)) return 1;
    FZB_CUDA(cudaEventRecord(h->ev[0], h->stream));
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        const int64_t nc = std::min(chunk, No - o0);
        FZB_CUDA(cudaMemcpyAsync(d_in.p, pdfs + (size_t)o0 * Ng, (size_t)nc * Ng * 8, cudaMemcpyHostToDevice, h->stream));
        SummParams P = {};
        P.pdfs = d_in.as<double>(); P.pgrid = d_grid.as<double>(); P.loss = d_loss.as<double>(); P.urand = d_u.as<double>();
        P.No = nc; P.Ng = Ng; P.LP = summ_lp(Ng); P.ncopy = summ_copies(); P.copy_stride = summ_copy_stride(Ng); P.nleaf = 0; summ_leaves(0, Ng, P); P.renorm = renormalize; P.rowsum = o_sum; P.cdf = d_cdf.as<double>();
        P.est = o_est; P.sd = o_sd; P.risk = o_risk; P.quant = o_quant; P.mc = o_mc; P.Ntot = No; P.o0 = o0;
        const int rc = launch_summ(h, P, nc);
        if (rc) return rc;
    }
    FZB_CUDA(cudaEventRecord(h->ev[1], h->stream));
    FZB_CUDA(cudaMemcpyAsync(est, o_est, (size_t)4 * No * 8, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(sd, o_sd, (size_t)4 * No * 8, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(risk, o_risk, (size_t)4 * No * 8, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(quant, o_quant, (size_t)4 * No * 8, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(mc, o_mc, (size_t)No * 8, cudaMemcpyDeviceToHost, h->stream));
    if (rowsum && renormalize) FZB_CUDA(cudaMemcpyAsync(rowsum, o_sum, (size_t)No * 8, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->stats.ms_total = ms;
    h->summ_No = No;
    h->summ_Ng = Ng;
    return 0;
}

// ---- summaries of PDF rows that are already on the device (fused behind fit_predict) ------------------------------
// The tables (grid, loss matrix, random numbers) are uploaded once per call by fzb_summarize_tables; every chunk of PDFs
// is then summarised where it lies, so that only 21 doubles per object cross PCIe instead of Ngrid.
int fzb_summarize_tables(fzb_context* h, const double* pgrid, const double* loss, const double* urand, int64_t No,
                         int32_t Ng) {
    FZB_CHECK(Ng >= 2 && Ng <= NGMAX, "pdfs_summarize: grid of %d points (supported: 2..%d)", Ng, NGMAX);
    if (h->summ[0].reserve((size_t)Ng * 8) || upload_loss(h, h->summ[1], loss, Ng) ||
        h->summ[3].reserve((size_t)No * 22 * 8 + 64) || h->summ[5].reserve((size_t)No * 8 + 64))
        return 1;
    FZB_CUDA(cudaMemcpyAsync(h->summ[0].p, pgrid, (size_t)Ng * 8, cudaMemcpyHostToDevice, h->stream));
    FZB_CUDA(cudaMemcpyAsync(h->summ[5].p, urand, (size_t)No * 8, cudaMemcpyHostToDevice, h->stream));
    h->summ_No = 0;      // no CDFs are kept: fzb_pdfs_conf does not apply to a fused run
    return 0;
}

// rows [o0, o0 + n) of the batch of Ntot objects; outputs in h->summ[3]: est, sd, risk, quant, conf [4][Ntot], mc [Ntot]
int fzb_summarize_rows_dev(fzb_context* h, const double* d_pdfs, int64_t n, int64_t o0, int64_t Ntot, int32_t Ng,
                           int32_t renormalize, double wfac) {
    double* o_est = h->summ[3].as<double>();
    SummParams P = {};
    P.pdfs = d_pdfs; P.pgrid = h->summ[0].as<double>(); P.loss = h->summ[1].as<double>(); P.urand = h->summ[5].as<double>();
    P.No = n; P.Ng = Ng; P.LP = summ_lp(Ng); P.ncopy = summ_copies(); P.copy_stride = summ_copy_stride(Ng); P.nleaf = 0; summ_leaves(0, Ng, P); P.renorm = renormalize; P.rowsum = nullptr; P.cdf = nullptr;
    P.est = o_est; P.sd = o_est + 4 * Ntot; P.risk = o_est + 8 * Ntot; P.quant = o_est + 12 * Ntot;
    P.conf = o_est + 16 * Ntot; P.mc = o_est + 20 * Ntot; P.wfac = wfac; P.Ntot = Ntot; P.o0 = o0;
    return launch_summ(h, P, n);
}

int fzb_summarize_download(fzb_context* h, int64_t No, double* est, double* sd, double* conf, double* risk, double* quant,
                           double* mc) {
    const double* o = h->summ[3].as<double>();
    const size_t b4 = (size_t)4 * No * 8;
    FZB_CUDA(cudaMemcpyAsync(est, o, b4, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(sd, o + 4 * No, b4, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(risk, o + 8 * No, b4, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(quant, o + 12 * No, b4, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(conf, o + 16 * No, b4, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaMemcpyAsync(mc, o + 20 * No, (size_t)No * 8, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

int fzb_conf_impl(fzb_context* h, const double* points, const double* widths, int64_t No, double* conf) {
    FZB_CHECK(h->summ_No == No && h->summ_Ng > 0, "fzb_pdfs_conf: call fzb_pdfs_summarize on the same %lld PDFs first",
              (long long)No);
    DevBuf& d_pw = h->summ[4];
    if (d_pw.reserve((size_t)No * 12 * 8 + 64)) return 1;
    double* d_points = d_pw.as<double>();
    double* d_widths = d_points + 4 * No;
    double* d_conf = d_widths + 4 * No;
    FZB_CUDA(cudaMemcpyAsync(d_points, points, (size_t)4 * No * 8, cudaMemcpyHostToDevice, h->stream));
    FZB_CUDA(cudaMemcpyAsync(d_widths, widths, (size_t)4 * No * 8, cudaMemcpyHostToDevice, h->stream));
    ConfParams P = {};
    P.pgrid = h->summ[0].as<double>(); P.cdf = h->summ[2].as<double>(); P.points = d_points; P.widths = d_widths;
    P.conf = d_conf; P.No = No; P.Ng = h->summ_Ng;
    k_conf<<<(unsigned)((4 * No + 255) / 256), 256, 0, h->stream>>>(P);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    FZB_CUDA(cudaMemcpyAsync(conf, d_conf, (size_t)4 * No * 8, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}
