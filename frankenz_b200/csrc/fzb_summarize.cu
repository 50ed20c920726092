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
// k_summarize: one CTA = 8 objects x 256 threads.  PDFs, CDFs and risk rows of the 8 objects live in shared memory;
// thread t owns grid columns t, t+256, ... of the GEMM with the 8 x NCOL accumulators in registers, the loss matrix
// streams from L2 (coalesced rows), the PDF values are shared-memory broadcasts; the sequential row sums / CDFs of the
// 8 objects run on 8 lanes of warp 0.
#include <algorithm>
#include <cfloat>

#include "fzb_common.cuh"

namespace {

constexpr int SB = 8;          // objects per CTA
constexpr int ST = 256;        // threads per CTA
constexpr int SMAXCOL = 4;     // grid columns per thread: Ng <= 1024

// numpy's pairwise summation of a contiguous float64 row (numpy/core/src/umath/loops_utils.h, DOUBLE_pairwise_sum)
__device__ double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const double left = np_pairwise_sum(a, n2);          // depth <= 4 for the supported grids (n <= 1024)
    const double right = np_pairwise_sum(a + n2, n - n2);
    return left + right;
}

// numpy.interp(x, xp, fp) for one point, xp non-decreasing (numpy/core/src/multiarray/compiled_base.c, arr_interp):
// j = last index with xp[j] <= x; exact hits return fp[j]; the slope is formed per point
__device__ double np_interp(double x, const double* xp, const double* fp, int n) {
    if (x != x) return x;
    if (x > xp[n - 1]) return fp[n - 1];
    if (x < xp[0]) return fp[0];
    int lo = 0, hi = n;                 // first index with xp[idx] > x
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid + 1;
        else hi = mid;
    }
    const int j = lo - 1;
    if (j == n - 1) return fp[j];
    if (xp[j] == x) return fp[j];
    const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    double r = slope * (x - xp[j]) + fp[j];
    if (r != r) {
        r = slope * (x - xp[j + 1]) + fp[j + 1];
        if (r != r && fp[j] == fp[j + 1]) r = fp[j];
    }
    return r;
}

struct SummParams {
    const double* pdfs;     // device (No x Ng) rows of this launch
    const double* pgrid;    // [Ng]
    const double* loss;     // (Ng x Ng): 1 - kernel[truth, guess]
    const double* urand;    // [No]
    int64_t No;
    int Ng, renorm;
    double* rowsum;         // [No]
    double* cdf;            // (No x Ng) out, kept for stage 2
    double *est, *sd, *risk, *quant, *mc;   // [4][Ntot] each (mc: [Ntot]); column offset o0
    double* conf;           // nullable [4][Ntot]: probability within +-wfac (1 + estimator), the default `wconf_func`
    double wfac;            // (pdf.py:1038-1062), formed while the CDF is still in shared memory
    int64_t Ntot, o0;
};

template <int NCOL>
__global__ void __launch_bounds__(ST, 1) k_summarize(SummParams P) {
    extern __shared__ __align__(16) double sm_s[];
    const int Ng = P.Ng;
    double* spdf = sm_s;                     // [SB][Ng]
    double* scdf = spdf + (size_t)SB * Ng;   // [SB][Ng]
    double* srsk = scdf + (size_t)SB * Ng;   // [SB][Ng]
    __shared__ double ssum[SB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ob = (int64_t)blockIdx.x * SB;

    for (int i = tid; i < SB * Ng; i += ST) {
        const int o = i / Ng, t = i - o * Ng;
        spdf[i] = (ob + o < P.No) ? P.pdfs[(ob + o) * Ng + t] : 0.0;
    }
    __syncthreads();
    if (P.renorm) {
        if (tid < SB) {
            const double s = np_pairwise_sum(spdf + (size_t)tid * Ng, Ng);
            ssum[tid] = s;
            if (ob + tid < P.No && P.rowsum) P.rowsum[P.o0 + ob + tid] = s;
        }
        __syncthreads();
        for (int i = tid; i < SB * Ng; i += ST) spdf[i] = spdf[i] / ssum[i / Ng];
        __syncthreads();
    }
    // sequential CDFs on 8 lanes of warp 0 (numpy.cumsum order); the other warps go straight to the GEMM
    if (tid < SB) {
        const double* p = spdf + (size_t)tid * Ng;
        double* c = scdf + (size_t)tid * Ng;
        double run = p[0];
        c[0] = run;
        for (int t = 1; t < Ng; ++t) {
            run = run + p[t];
            c[t] = run;
        }
    }
    // ---- risk curves: acc[o][c] = sum_t pdf[o][t] * loss[t][col_c] -------------------------------------------
    double acc[SB][NCOL];
#pragma unroll
    for (int o = 0; o < SB; ++o)
#pragma unroll
        for (int c = 0; c < NCOL; ++c) acc[o][c] = 0.0;
    int col[NCOL];
#pragma unroll
    for (int c = 0; c < NCOL; ++c) col[c] = min(tid + c * ST, Ng - 1);
    for (int t = 0; t < Ng; ++t) {
        double k[NCOL];
        const double* lrow = P.loss + (size_t)t * Ng;
#pragma unroll
        for (int c = 0; c < NCOL; ++c) k[c] = __ldg(lrow + col[c]);
#pragma unroll
        for (int o = 0; o < SB; ++o) {
            const double pv = spdf[(size_t)o * Ng + t];
#pragma unroll
            for (int c = 0; c < NCOL; ++c) acc[o][c] = fma(pv, k[c], acc[o][c]);
        }
    }
#pragma unroll
    for (int c = 0; c < NCOL; ++c)
        if (tid + c * ST < Ng)
#pragma unroll
            for (int o = 0; o < SB; ++o) srsk[(size_t)o * Ng + tid + c * ST] = acc[o][c];
    __syncthreads();

    // ---- per-object reductions: warp w <-> object w -------------------------------------------------------------
    const int o = warp;
    const int64_t og = ob + o;
    const double* p = spdf + (size_t)o * Ng;
    const double* c = scdf + (size_t)o * Ng;
    const double* r = srsk + (size_t)o * Ng;
    double mean = 0.0, pmax = -DBL_MAX, rmin = DBL_MAX;
    int imax = 0x7fffffff, imin = 0x7fffffff;
    for (int t = lane; t < Ng; t += 32) {
        mean = fma(p[t], P.pgrid[t], mean);
        if (p[t] > pmax) { pmax = p[t]; imax = t; }
        if (r[t] < rmin) { rmin = r[t]; imin = t; }
    }
    for (int s = 16; s > 0; s >>= 1) {
        mean += __shfl_xor_sync(0xffffffffu, mean, s);
        const double om = __shfl_xor_sync(0xffffffffu, pmax, s);
        const int oi = __shfl_xor_sync(0xffffffffu, imax, s);
        if (om > pmax || (om == pmax && oi < imax)) { pmax = om; imax = oi; }      // first maximum (numpy.argmax)
        const double orr = __shfl_xor_sync(0xffffffffu, rmin, s);
        const int oj = __shfl_xor_sync(0xffffffffu, imin, s);
        if (orr < rmin || (orr == rmin && oj < imin)) { rmin = orr; imin = oj; }   // first minimum (numpy.argmin)
    }
    if (imax == 0x7fffffff) imax = 0;      // all-NaN row: numpy returns the first NaN; not reproduced
    if (imin == 0x7fffffff) imin = 0;
    // quantiles: lanes 0..5
    const double qs[6] = {0.025, 0.16, 0.5, 0.84, 0.975, 0.0};
    double qv = 0.0;
    if (lane < 6) {
        const double q = (lane == 5) ? ((og < P.No) ? P.urand[P.o0 + og] : 0.5) : qs[lane];
        qv = np_interp(q, c, P.pgrid, Ng);
    }
    const double med = __shfl_sync(0xffffffffu, qv, 2);
    const double pts[4] = {mean, med, P.pgrid[imax], P.pgrid[imin]};
    double sd[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = lane; t < Ng; t += 32) {
        const double g = P.pgrid[t];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double d = g - pts[e];
            sd[e] = fma(d * d, p[t], sd[e]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
        for (int s = 16; s > 0; s >>= 1) sd[e] += __shfl_xor_sync(0xffffffffu, sd[e], s);
    if (og < P.No) {
        const int64_t q = P.o0 + og;
        if (lane < 4) {
            P.est[(size_t)lane * P.Ntot + q] = pts[lane];
            P.sd[(size_t)lane * P.Ntot + q] = sqrt(sd[lane]);
            P.risk[(size_t)lane * P.Ntot + q] = np_interp(pts[lane], P.pgrid, r, Ng);
        }
        if (lane < 5 && lane != 2) P.quant[(size_t)(lane < 2 ? lane : lane - 1) * P.Ntot + q] = qv;
        if (lane == 5) P.mc[q] = qv;
        if (P.conf && lane < 4) {
            const double w = (1. + pts[lane]) * P.wfac;
            const double lo = np_interp(pts[lane] - w, P.pgrid, c, Ng);
            const double hi = np_interp(pts[lane] + w, P.pgrid, c, Ng);
            P.conf[(size_t)lane * P.Ntot + q] = hi - lo;
        }
    }
    if (P.cdf == nullptr) return;
    __syncthreads();
    for (int i = tid; i < SB * Ng; i += ST) {
        const int oo = i / Ng;
        if (ob + oo < P.No) P.cdf[(size_t)(P.o0 + ob + oo) * Ng + (i - oo * Ng)] = scdf[i];
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

template <int NCOL>
int launch_summ(fzb_context* h, const SummParams& P, int64_t nobj) {
    const size_t smem = (size_t)3 * SB * P.Ng * sizeof(double);
    FZB_CUDA(cudaFuncSetAttribute(k_summarize<NCOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_summarize<NCOL><<<(unsigned)((nobj + SB - 1) / SB), ST, smem, h->stream>>>(P);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int fzb_summarize_impl(fzb_context* h, const double* pdfs, const double* pgrid, const double* loss, const double* urand,
                       int64_t No, int32_t Ng, int32_t renormalize, double* rowsum, double* est, double* sd, double* risk,
                       double* quant, double* mc) {
    FZB_CHECK(Ng >= 2 && Ng <= ST * SMAXCOL, "pdfs_summarize: grid of %d points (supported: 2..%d)", Ng, ST * SMAXCOL);
    h->stats = FzbStats{};
    DevBuf& d_grid = h->summ[0];
    DevBuf& d_loss = h->summ[1];
    DevBuf& d_cdf = h->summ[2];
    DevBuf& d_out = h->summ[3];
    DevBuf& d_in = h->summ[4];
    DevBuf& d_u = h->summ[5];
    if (d_grid.reserve((size_t)Ng * 8) || d_loss.reserve((size_t)Ng * Ng * 8) || d_cdf.reserve((size_t)No * Ng * 8 + 64) ||
        d_out.reserve((size_t)No * 18 * 8 + 64) || d_u.reserve((size_t)No * 8 + 64))
        return 1;
    FZB_CUDA(cudaMemcpyAsync(d_grid.p, pgrid, (size_t)Ng * 8, cudaMemcpyHostToDevice, h->stream));
    FZB_CUDA(cudaMemcpyAsync(d_loss.p, loss, (size_t)Ng * Ng * 8, cudaMemcpyHostToDevice, h->stream));
    FZB_CUDA(cudaMemcpyAsync(d_u.p, urand, (size_t)No * 8, cudaMemcpyHostToDevice, h->stream));
    double* o_est = d_out.as<double>();
    double* o_sd = o_est + 4 * No;
    double* o_risk = o_sd + 4 * No;
    double* o_quant = o_risk + 4 * No;
    double* o_mc = o_quant + 4 * No;
    double* o_sum = o_mc + No;
    const int64_t chunk = std::max<int64_t>(SB, std::min<int64_t>(No, ((int64_t)256 << 20) / ((int64_t)Ng * 8) / SB * SB));
    if (d_in.reserve((size_t)chunk * Ng * 8)) return 1;
    FZB_CUDA(cudaEventRecord(h->ev[0], h->stream));
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        const int64_t nc = std::min(chunk, No - o0);
        FZB_CUDA(cudaMemcpyAsync(d_in.p, pdfs + (size_t)o0 * Ng, (size_t)nc * Ng * 8, cudaMemcpyHostToDevice, h->stream));
        SummParams P = {};
        P.pdfs = d_in.as<double>(); P.pgrid = d_grid.as<double>(); P.loss = d_loss.as<double>(); P.urand = d_u.as<double>();
        P.No = nc; P.Ng = Ng; P.renorm = renormalize; P.rowsum = o_sum; P.cdf = d_cdf.as<double>();
        P.est = o_est; P.sd = o_sd; P.risk = o_risk; P.quant = o_quant; P.mc = o_mc; P.Ntot = No; P.o0 = o0;
        int rc;
        if (Ng <= ST) rc = launch_summ<1>(h, P, nc);
        else if (Ng <= 2 * ST) rc = launch_summ<2>(h, P, nc);
        else if (Ng <= 3 * ST) rc = launch_summ<3>(h, P, nc);
        else rc = launch_summ<4>(h, P, nc);
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
    FZB_CHECK(Ng >= 2 && Ng <= ST * SMAXCOL, "pdfs_summarize: grid of %d points (supported: 2..%d)", Ng, ST * SMAXCOL);
    if (h->summ[0].reserve((size_t)Ng * 8) || h->summ[1].reserve((size_t)Ng * Ng * 8) ||
        h->summ[3].reserve((size_t)No * 22 * 8 + 64) || h->summ[5].reserve((size_t)No * 8 + 64))
        return 1;
    FZB_CUDA(cudaMemcpyAsync(h->summ[0].p, pgrid, (size_t)Ng * 8, cudaMemcpyHostToDevice, h->stream));
    FZB_CUDA(cudaMemcpyAsync(h->summ[1].p, loss, (size_t)Ng * Ng * 8, cudaMemcpyHostToDevice, h->stream));
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
    P.No = n; P.Ng = Ng; P.renorm = renormalize; P.rowsum = nullptr; P.cdf = nullptr;
    P.est = o_est; P.sd = o_est + 4 * Ntot; P.risk = o_est + 8 * Ntot; P.quant = o_est + 12 * Ntot;
    P.conf = o_est + 16 * Ntot; P.mc = o_est + 20 * Ntot; P.wfac = wfac; P.Ntot = Ntot; P.o0 = o0;
    if (Ng <= ST) return launch_summ<1>(h, P, n);
    if (Ng <= 2 * ST) return launch_summ<2>(h, P, n);
    if (Ng <= 3 * ST) return launch_summ<3>(h, P, n);
    return launch_summ<4>(h, P, n);
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
