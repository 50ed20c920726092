// Generic float64 path: one CTA per object, reference operation order.
//
// This is the "everything, exactly" implementation of the path: every likelihood flavour
// (frankenz/pdf.py:27-235), arbitrary multiplicative masks, the batch-coupled do-while of the
// iterated free-scale mode (pdf.py:197-223), both KDE flavours (pdf.py:444-622) with both
// thresholding rules, the kNN gather form (knn.py:375-386, :541-555) and the model-sharded
// partial passes.  It is bound by L2 bandwidth (each pair re-reads its model row), so the
// fp32 register-tiled kernels in fzb_fast.cu take over for the large reduce-only shapes and
// route objects back here when their fp32 error bound is too large.
#include <math_constants.h>

#include "fzb_common.cuh"
#include "fzb_pair64.cuh"

namespace {

using namespace fzb64;

constexpr int GT = 256;  // threads per CTA
constexpr double kSqrt2Pi = 2.50662827463100050242;

enum Stage { ST_FIT = 0, ST_FIT_PREDICT = 1, ST_PASS1 = 2, ST_PASS2 = 3, ST_LOGWT = 4 };

struct KdeDev {
    int mode;  // FZB_KDE_*
    int Ng;
    const int32_t* widths;
    const int64_t* koff;
    const double* kernels;
    const double* kcdf;
    const int64_t* yidx;
    const int64_t* ysidx;
    const double* grid;
    const double* y;
    const double* ystd;
    const int64_t* lowers;
    const int64_t* uppers;
    int use_wt, use_cdf;
    double wt_thresh, cdf_thresh;
    long long* err;            // nullable: {flag, model} set when a selected label cannot be placed on the grid
    int raw;                   // 1: leave the stack of kernels un-normalised (module-level gauss_kde / gauss_kde_dict)
};

struct GenParams {
    int stage;
    const double *x, *xe, *xm;        // objects (No x Nf)
    const double *m, *me, *mm;        // models (Nm x Nf)
    const double* lnprior;            // per model, nullable
    const double* prior_table;        // [nbins][Nm], nullable: row prior_bins[o - prior_o0] replaces lnprior
    const int32_t* prior_bins;
    int64_t No, Nm;
    int Nf;
    int free_scale, ime, iterate, dim_prior, track_scale;
    double ltol;
    // kNN gather (nullable): object i uses models nbr[i*W + c], c < nnbr[i]
    const int64_t* nbr;
    const int64_t* nnbr;
    int64_t W;                        // row width of outputs / logwt
    const int32_t* objsel;            // nullable list of object indices to process
    int64_t Nsel;
    // full outputs (nullable), row-major (No x W)
    double *o_lnprior, *o_lnlike, *o_lnprob, *o_chi2, *o_scale, *o_scale_err;
    int64_t* o_ndim;
    // per-CTA scratch rows: 5 x W doubles per CTA
    double* rows;
    // caller supplied log-weights (ST_LOGWT)
    const double* logwt;
    // per-object outputs
    double *pdfs, *lmap, *levid, *best_chi2, *best_scale;
    float* pdfs32;                    // ST_PASS2: fp32 partial rows instead of `pdfs`
    int64_t* best_idx;
    // sharded passes
    double *pmax, *psum;
    int64_t* pbest;
    const double *g_lmap, *g_levid;
    KdeDev kde;
};

// ---- block reductions ------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < GT / 32; ++w) t += red[w];
    return t;
}
// max with '>' comparisons only: NaNs never win (Python builtin max, elements after the first)
__device__ __forceinline__ double block_max_gt(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        if (u > v) v = u;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int w = 1; w < GT / 32; ++w)
        if (red[w] > t) t = red[w];
    return t;
}
__device__ __forceinline__ int block_or(int v, int* red) {
    v = __any_sync(0xffffffffu, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < GT / 32; ++w) t |= red[w];
    return t;
}
// (value, index) arg-max with '>' and lowest index on ties; NaN never wins
__device__ __forceinline__ void block_argmax(double& v, long long& i, double* red, long long* redi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        long long ui = __shfl_xor_sync(0xffffffffu, i, o);
        if (u > v || (u == v && ui < i)) { v = u; i = ui; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = v; redi[threadIdx.x >> 5] = i; }
    __syncthreads();
    v = red[0];
    i = redi[0];
#pragma unroll
    for (int w = 1; w < GT / 32; ++w)
        if (red[w] > v || (red[w] == v && redi[w] < i)) { v = red[w]; i = redi[w]; }
}

// ---- KDE scatter of one selected model by one warp ----------------------------------------------
__device__ __forceinline__ void kde_add_dict(const KdeDev& k, int64_t mj, double wt, double* s_pdf, int lane) {
    long long pos = k.yidx[mj];
    int si = (int)k.ysidx[mj];
    long long w = k.widths[si];
    const double* kern = k.kernels + k.koff[si];
    const double* cdf = k.kcdf + k.koff[si];
    long long len = 2 * w + 1;
    if (k.err && (k.koff[si + 1] - k.koff[si] != len || pos + w < 0 || pos - w > k.Ng - 1)) {
        if (lane == 0 && atomicCAS(reinterpret_cast<unsigned long long*>(k.err), 0ull, 1ull) == 0ull) k.err[1] = mj;
        return;
    }
    long long low = pos - w > 0 ? pos - w : 0;
    long long high = pos + w + 1 < k.Ng ? pos + w + 1 : k.Ng;
    long long lpad = low - (pos - w), hpad = high - (pos + w + 1);
    long long hi_i = hpad - 1;              // <= -1: Python negative index from the end
    double norm = cdf[len + hi_i];
    if (lpad != 0) norm -= cdf[lpad - 1];
    double coef = wt / norm;
    for (long long t = lane; t < high - low; t += 32) atomicAdd(&s_pdf[low + t], coef * kern[lpad + t]);
}
__device__ __forceinline__ void kde_add_grid(const KdeDev& k, int64_t mj, double wt, double* s_pdf, int lane) {
    long long lo = k.lowers[mj], up = k.uppers[mj];
    double mu = k.y[mj], sd = k.ystd[mj];
    double nrm = kSqrt2Pi * sd;
    double part = 0.0;
    for (long long t = lo + lane; t < up; t += 32) {
        double d = (k.grid[t] - mu) / sd;
        part += exp(-0.5 * (d * d)) / nrm;
    }
    double tot = warp_sum(part);
    if (tot != 0.0) {   // pdf.py:523 (NaN != 0 is true there as well)
        double coef = wt / tot;
        for (long long t = lo + lane; t < up; t += 32) {
            double d = (k.grid[t] - mu) / sd;
            atomicAdd(&s_pdf[t], coef * (exp(-0.5 * (d * d)) / nrm));
        }
    }
}

// Row of log-weights -> lmap, levid, (un)normalised PDF.   bruteforce.py:358-372
//   row[c], c < n ; model index of column c is map ? map[c] : c
//   given_lmap/levid: non-null for the model-sharded pass 2 (global values, no normalisation)
__device__ void row_to_pdf(const double* row, int64_t n, const int64_t* map, const KdeDev& k, double* s_pdf,
                           double* red, int* redi, double* out_pdf, double* out_lmap, double* out_levid,
                           const double* given_lmap, const double* given_levid, float* out_pdf32 = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double lmap, levid, amax;
    int has_nan = 0;
    if (given_lmap == nullptr) {
        // lmap = Python builtin max (first element wins if NaN); amax = numpy max (NaN-propagating)
        double v = -CUDART_INF;
        for (int64_t c = tid; c < n; c += GT) {
            double a = row[c];
            if (isnan(a)) has_nan = 1;
            if (a > v) v = a;
        }
        v = block_max_gt(v, red);
        has_nan = block_or(has_nan, redi);
        double first = n > 0 ? row[0] : CUDART_NAN;
        lmap = isnan(first) ? first : v;
        amax = has_nan ? CUDART_NAN : v;
        if (has_nan) {
            levid = CUDART_NAN;
        } else if (isinf(amax)) {
            levid = amax;   // all -inf (or a +inf entry)
        } else {
            // scipy >= 1.15 logsumexp: log1p(sum_{a != amax} exp(a - amax) / m) + log(m) + amax
            double s = 0.0, cnt = 0.0;
            for (int64_t c = tid; c < n; c += GT) {
                double a = row[c];
                if (a == amax) cnt += 1.0;
                else s += exp(a - amax);
            }
            s = block_sum(s, red);
            cnt = block_sum(cnt, red);
            if (s != 0.0) s = s / cnt;
            levid = log1p(s) + log(cnt) + amax;
        }
        if (tid == 0) {
            if (out_lmap) *out_lmap = lmap;
            if (out_levid) *out_levid = levid;
        }
    } else {
        // model-sharded pass 2: weights are exp(l - GLOBAL lmap) on every rank and in every kernel (the common
        // factor cancels when the summed partials are normalised), selection is wt > wt_thresh
        lmap = *given_lmap;
        levid = lmap;
        amax = lmap;
    }
    if (out_pdf == nullptr && out_pdf32 == nullptr) return;

    for (int g = tid; g < k.Ng; g += GT) s_pdf[g] = 0.0;
    __syncthreads();

    // selection threshold (pdf.py:589-597 / :508-516)
    double thr_lo = -CUDART_INF;   // select wt > thr_lo
    double thr_hi = CUDART_INF;    // and wt <= thr_hi (CDF rule keeps the LOW end of the sorted weights)
    bool none_selected = false;
    if (k.use_wt) {
        double wmax = exp(amax - levid);          // np.max(y_wt): NaN propagates
        thr_lo = k.wt_thresh * wmax;
        if (isnan(thr_lo)) none_selected = true;
    } else if (k.use_cdf) {
        double tot = 0.0;
        for (int64_t c = tid; c < n; c += GT) tot += exp(row[c] - levid);
        tot = block_sum(tot, red);
        if (isnan(tot)) {
            none_selected = true;
        } else {
            // largest weight value t with sum_{wt <= t} wt / tot <= 1 - cdf_thresh (bisection on bit patterns)
            unsigned long long lo_b = 0ull, hi_b = 0x7ff0000000000000ull;  // [0, +inf]
            const double target = 1.0 - k.cdf_thresh;
            bool any_ok = false;
            for (int it = 0; it < 64 && lo_b < hi_b; ++it) {
                unsigned long long mid = lo_b + (hi_b - lo_b + 1) / 2;
                double tv = __longlong_as_double((long long)mid);
                double part = 0.0;
                for (int64_t c = tid; c < n; c += GT) {
                    double w = exp(row[c] - levid);
                    if (w <= tv) part += w;
                }
                part = block_sum(part, red);
                if (part / tot <= target) { lo_b = mid; any_ok = true; }
                else hi_b = mid - 1;
            }
            if (!any_ok) {
                // check t = 0 itself
                double part = 0.0;
                for (int64_t c = tid; c < n; c += GT) {
                    double w = exp(row[c] - levid);
                    if (w <= 0.0) part += w;
                }
                part = block_sum(part, red);
                if (!(part / tot <= target)) none_selected = true;
            }
            thr_hi = __longlong_as_double((long long)lo_b);
        }
    }

    if (!none_selected) {
        const int64_t nround = (n + GT - 1) / GT * GT;
        for (int64_t c0 = warp * 32; c0 < nround; c0 += GT) {
            int64_t c = c0 + lane;
            double wt = 0.0;
            bool sel = false;
            if (c < n) {
                wt = exp(row[c] - levid);
                sel = (wt > thr_lo) && (wt <= thr_hi);
            }
            unsigned bal = __ballot_sync(0xffffffffu, sel);
            while (bal) {
                int src = __ffs(bal) - 1;
                bal &= bal - 1;
                double w = __shfl_sync(0xffffffffu, wt, src);
                int64_t cc = c0 + src;
                int64_t mj = map ? map[cc] : cc;
                if (k.mode == FZB_KDE_DICT) kde_add_dict(k, mj, w, s_pdf, lane);
                else kde_add_grid(k, mj, w, s_pdf, lane);
            }
        }
    }
    __syncthreads();
    if (given_lmap != nullptr) {
        if (out_pdf32) { for (int g = tid; g < k.Ng; g += GT) out_pdf32[g] = (float)s_pdf[g]; }
        else { for (int g = tid; g < k.Ng; g += GT) out_pdf[g] = s_pdf[g]; }
        return;
    }
    if (k.raw) {
        for (int g = tid; g < k.Ng; g += GT) out_pdf[g] = s_pdf[g];
        return;
    }
    double tot = 0.0;
    for (int g = tid; g < k.Ng; g += GT) tot += s_pdf[g];
    tot = block_sum(tot, red);
    for (int g = tid; g < k.Ng; g += GT) out_pdf[g] = s_pdf[g] / tot;   // bruteforce.py:370
}

__global__ void __launch_bounds__(GT) k_generic(GenParams P) {
    extern __shared__ double smem[];
    double* sx = smem;
    double* sxe = sx + FZB_MAXF;
    double* sxm = sxe + FZB_MAXF;
    double* red = sxm + FZB_MAXF;                       // 8
    long long* redl = reinterpret_cast<long long*>(red + 8);   // 8
    int* redi = reinterpret_cast<int*>(redl + 8);       // 8 ints (4 doubles reserved)
    double* s_pdf = red + 8 + 8 + 4;
    const int tid = threadIdx.x;
    const int Nf = P.Nf;
    const int64_t W = P.W;
    double* r_lnl = P.rows ? P.rows + (size_t)blockIdx.x * 5 * W : nullptr;
    double* r_scale = r_lnl ? r_lnl + W : nullptr;
    double* r_chi2 = r_lnl ? r_lnl + 2 * W : nullptr;
    double* r_shape = r_lnl ? r_lnl + 3 * W : nullptr;
    double* r_ndim = r_lnl ? r_lnl + 4 * W : nullptr;

    const int64_t count = P.objsel ? P.Nsel : P.No;
    for (int64_t it = blockIdx.x; it < count; it += gridDim.x) {
        const int64_t o = P.objsel ? P.objsel[it] : it;
        __syncthreads();
        if (P.stage == ST_LOGWT) {
            const int64_t n = P.nbr ? P.nnbr[o] : W;
            row_to_pdf(P.logwt + (size_t)o * W, n, P.nbr ? P.nbr + (size_t)o * W : nullptr, P.kde, s_pdf, red, redi,
                       P.pdfs + (size_t)o * P.kde.Ng, P.lmap + o, P.levid + o, nullptr, nullptr);
            continue;
        }
        // load + clean the object (pdf.py:310-311)
        if (tid < Nf) {
            double a = P.x[o * Nf + tid], e = P.xe[o * Nf + tid], k = P.xm[o * Nf + tid];
            bool clean = isfinite(a) && isfinite(e) && (e > 0.0);
            sx[tid] = clean ? a : 0.0;
            sxe[tid] = clean ? e : 1.0;
            sxm[tid] = clean ? k : 0.0;
        }
        __syncthreads();
        const int64_t n = P.nbr ? P.nnbr[o] : P.Nm;
        const int64_t* map = P.nbr ? P.nbr + (size_t)o * W : nullptr;

        for (int64_t c = tid; c < n; c += GT) {
            int64_t mj = map ? map[c] : c;
            PairState st;
            pair_first(sx, sxe, sxm, P.m + mj * Nf, P.me + mj * Nf, P.mm + mj * Nf, Nf, P.free_scale, P.ime, st);
            r_lnl[c] = st.lnl;
            r_scale[c] = st.scale;
            r_chi2[c] = st.chi2;
            r_shape[c] = st.shape;
            r_ndim[c] = st.ndim;
        }
        if (P.iterate) {
            // do-while with the object-wide stopping rule (pdf.py:199-223)
            bool again = true;
            while (again) {
                double worst = -CUDART_INF;
                int nan0 = 0;
                for (int64_t c = tid; c < n; c += GT) {
                    int64_t mj = map ? map[c] : c;
                    double sc, c2, ln, sh;
                    pair_refine(sx, sxe, sxm, P.m + mj * Nf, P.me + mj * Nf, P.mm + mj * Nf, Nf, r_ndim[c], r_scale[c],
                                sc, c2, ln, sh);
                    double d = fabs(ln - r_lnl[c]);
                    if (c == 0 && isnan(d)) nan0 = 1;
                    if (d > worst) worst = d;
                    r_lnl[c] = ln;
                    r_scale[c] = sc;
                    r_chi2[c] = c2;
                    r_shape[c] = sh;
                }
                worst = block_max_gt(worst, red);
                nan0 = block_or(nan0, redi);
                again = (!nan0) && (worst > P.ltol);
            }
        }
        // finalise: dimensionality prior, prior, outputs, arg-max
        const double* lnp = P.prior_table ? P.prior_table + (size_t)P.prior_bins[o] * P.Nm : P.lnprior;
        double bv = -CUDART_INF;
        long long bi = 0x7fffffffffffffffll;
        for (int64_t c = tid; c < n; c += GT) {
            int64_t mj = map ? map[c] : c;
            double ndim = r_ndim[c], chi2 = r_chi2[c], lnl = r_lnl[c];
            if (P.dim_prior) lnl = chi2_logpdf(chi2, P.free_scale ? 0.5 * (ndim - 1.0) : 0.5 * ndim);
            double lp = lnp ? lnp[mj] : 0.0;
            double lpost = lnp ? lnl + lp : lnl;
            r_lnl[c] = lpost;
            if (lpost > bv) { bv = lpost; bi = c; }
            if (P.stage == ST_FIT) {
                size_t q = (size_t)o * W + c;
                if (P.o_lnprior) P.o_lnprior[q] = lp;
                if (P.o_lnlike) P.o_lnlike[q] = lnl;
                if (P.o_lnprob) P.o_lnprob[q] = lpost;
                if (P.o_ndim) P.o_ndim[q] = (long long)ndim;
                if (P.o_chi2) P.o_chi2[q] = chi2;
                if (P.o_scale) P.o_scale[q] = (P.track_scale && P.free_scale) ? r_scale[c] : 1.0;
                if (P.o_scale_err)
                    P.o_scale_err[q] = (P.track_scale && P.free_scale) ? sqrt(1.0 / r_shape[c]) : 0.0;
            }
        }
        if (P.stage == ST_FIT) {
            if (P.nbr) {   // padding of the kNN arrays (knn.py:342-352)
                for (int64_t c = n + tid; c < W; c += GT) {
                    size_t q = (size_t)o * W + c;
                    if (P.o_lnprior) P.o_lnprior[q] = -CUDART_INF;
                    if (P.o_lnlike) P.o_lnlike[q] = -CUDART_INF;
                    if (P.o_lnprob) P.o_lnprob[q] = -CUDART_INF;
                    if (P.o_ndim) P.o_ndim[q] = 0;
                    if (P.o_chi2) P.o_chi2[q] = CUDART_INF;
                    if (P.o_scale) P.o_scale[q] = 1.0;
                    if (P.o_scale_err) P.o_scale_err[q] = 0.0;
                }
            }
            continue;
        }
        __syncthreads();   // row complete
        if (P.stage == ST_FIT_PREDICT || P.stage == ST_PASS1) {
            block_argmax(bv, bi, red, redl);
            if (bi == 0x7fffffffffffffffll) bi = 0;
        }
        if (P.stage == ST_PASS1) {
            // partial (max, sum exp(l - max), argmax) over the local models
            double s = 0.0;
            int has_nan = 0;
            for (int64_t c = tid; c < n; c += GT) {
                double a = r_lnl[c];
                if (isnan(a)) has_nan = 1;
                else if (a != -CUDART_INF || bv != -CUDART_INF) s += exp(a - bv);
            }
            s = block_sum(s, red);
            has_nan = block_or(has_nan, redi);
            if (tid == 0) {
                P.pmax[o] = has_nan ? CUDART_NAN : bv;
                P.psum[o] = has_nan ? CUDART_NAN : (isinf(bv) ? 0.0 : s);
                P.pbest[o] = map ? map[bi] : bi;
            }
            continue;
        }
        if (P.stage == ST_PASS2) {
            row_to_pdf(r_lnl, n, map, P.kde, s_pdf, red, redi, P.pdfs32 ? nullptr : P.pdfs + (size_t)o * P.kde.Ng, nullptr,
                       nullptr, P.g_lmap + o, P.g_levid + o, P.pdfs32 ? P.pdfs32 + (size_t)o * P.kde.Ng : nullptr);
            continue;
        }
        // ST_FIT_PREDICT
        if (tid == 0) {
            if (P.best_idx) P.best_idx[o] = map ? map[bi] : bi;
            if (P.best_chi2) P.best_chi2[o] = r_chi2[bi];
            if (P.best_scale) P.best_scale[o] = r_scale[bi];
        }
        row_to_pdf(r_lnl, n, map, P.kde, s_pdf, red, redi, P.pdfs ? P.pdfs + (size_t)o * P.kde.Ng : nullptr,
                   P.lmap ? P.lmap + o : nullptr, P.levid ? P.levid + o : nullptr, nullptr, nullptr);
    }
}

KdeDev make_kde(const fzb_context* h, const FzbConfig& cfg) {
    KdeDev k = {};
    k.mode = h->kde_mode;
    k.Ng = h->Ng;
    k.widths = h->widths.as<int32_t>();
    k.koff = h->koff.as<int64_t>();
    k.kernels = h->kernels.as<double>();
    k.kcdf = h->kcdf.as<double>();
    k.yidx = h->yidx.as<int64_t>();
    k.ysidx = h->ysidx.as<int64_t>();
    k.grid = h->grid.as<double>();
    k.y = h->y.as<double>();
    k.ystd = h->ystd.as<double>();
    k.lowers = h->lowers.as<int64_t>();
    k.uppers = h->uppers.as<int64_t>();
    k.use_wt = cfg.use_wt_thresh;
    k.use_cdf = cfg.use_wt_thresh ? 0 : cfg.use_cdf_thresh;
    k.wt_thresh = cfg.wt_thresh;
    k.cdf_thresh = cfg.cdf_thresh;
    k.raw = cfg.reserved & 1;
    k.err = (h->labels_bad > 0 && h->kde_mode == FZB_KDE_DICT) ? h->kde_err.as<long long>() : nullptr;
    return k;
}

size_t gen_smem_bytes(int Ng) { return sizeof(double) * (3 * FZB_MAXF + 8 + 8 + 4 + (size_t)Ng + 8); }

int launch_generic(fzb_context* h, GenParams& P, int64_t work_items, bool needs_rows) {
    FZB_CHECK(P.Nf <= FZB_MAXF, "Nf=%d exceeds the supported maximum %d", P.Nf, FZB_MAXF);
    int Ng = (P.stage == ST_FIT || P.stage == ST_PASS1) ? 0 : P.kde.Ng;
    FZB_CHECK(Ng <= FZB_MAX_NGRID, "PDF grid of %d points exceeds the supported maximum %d", Ng, FZB_MAX_NGRID);
    if (work_items <= 0) return 0;
    int64_t grid = (int64_t)h->sm_count * 4;
    if (needs_rows) {
        const size_t budget = (size_t)3 << 30;
        size_t per = (size_t)5 * P.W * sizeof(double);
        int64_t fit = (int64_t)(budget / (per ? per : 1));
        if (fit < 1) fit = 1;
        if (grid > fit) grid = fit;
    }
    if (grid > work_items) grid = work_items;
    if (needs_rows) {
        if (h->rows.reserve((size_t)grid * 5 * P.W * sizeof(double))) return 1;
        P.rows = h->rows.as<double>();
    }
    size_t smem = gen_smem_bytes(Ng);
    FZB_CUDA(cudaFuncSetAttribute(k_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_generic<<<(unsigned)grid, GT, smem, h->stream>>>(P);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

void fill_common(fzb_context* h, GenParams& P, const double* x, const double* xe, const double* xm, int64_t No,
                 const FzbConfig& cfg) {
    P.x = x; P.xe = xe; P.xm = xm;
    P.m = h->models.as<double>();
    P.me = h->models_err.as<double>();
    P.mm = h->models_mask.as<double>();
    P.lnprior = h->has_lnprior ? h->lnprior.as<double>() : nullptr;
    if (h->prior_nbins > 0 && h->prior_bins_n > 0) {
        P.prior_table = h->prior_table.as<double>();
        P.prior_bins = h->prior_bins.as<int32_t>() + h->prior_o0;
    }
    P.No = No; P.Nm = h->Nm; P.Nf = h->Nf;
    P.free_scale = cfg.free_scale; P.ime = cfg.ignore_model_err != 0; P.dim_prior = cfg.dim_prior;
    P.iterate = cfg.free_scale && cfg.ignore_model_err != 1;   // pdf.py:197: `ignore_model_err is not True`
    P.track_scale = cfg.track_scale; P.ltol = cfg.ltol;
    P.W = h->Nm;
}

int check_kde(const fzb_context* h) {
    FZB_CHECK(h->kde_mode != FZB_KDE_NONE, "no KDE configured: call fzb_set_kde_dict or fzb_set_kde_grid first");
    if (h->kde_mode == FZB_KDE_DICT) FZB_CHECK(h->labels_dict_set, "dictionary labels not set");
    if (h->kde_mode == FZB_KDE_GRID) FZB_CHECK(h->labels_grid_set, "grid labels not set");
    return 0;
}

}  // namespace

int fzb_generic_fit_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                        const FzbConfig& cfg, double* d_lnprior, double* d_lnlike, double* d_lnprob, int64_t* d_ndim,
                        double* d_chi2, double* d_scale, double* d_scale_err) {
    GenParams P = {};
    fill_common(h, P, d_x, d_xe, d_xm, No, cfg);
    P.stage = ST_FIT;
    P.o_lnprior = d_lnprior; P.o_lnlike = d_lnlike; P.o_lnprob = d_lnprob; P.o_ndim = d_ndim;
    P.o_chi2 = d_chi2; P.o_scale = d_scale; P.o_scale_err = d_scale_err;
    h->stats.pairs_fp64 += No * h->Nm;
    return launch_generic(h, P, No, true);
}

int fzb_generic_gather_fit_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                               int64_t W, const int64_t* d_neighbors, const int64_t* d_nneighbors,
                               const FzbConfig& cfg, double* d_lnprior, double* d_lnlike, double* d_lnprob,
                               int64_t* d_ndim, double* d_chi2, double* d_scale, double* d_scale_err) {
    GenParams P = {};
    fill_common(h, P, d_x, d_xe, d_xm, No, cfg);
    P.stage = ST_FIT;
    P.W = W; P.nbr = d_neighbors; P.nnbr = d_nneighbors;
    P.o_lnprior = d_lnprior; P.o_lnlike = d_lnlike; P.o_lnprob = d_lnprob; P.o_ndim = d_ndim;
    P.o_chi2 = d_chi2; P.o_scale = d_scale; P.o_scale_err = d_scale_err;
    h->stats.pairs_fp64 += No * W;
    return launch_generic(h, P, No, true);
}

int fzb_generic_fit_predict_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                                const int32_t* d_objsel, int64_t Nsel, const FzbConfig& cfg, double* d_pdfs,
                                double* d_lmap, double* d_levid, int64_t* d_best_idx, double* d_best_chi2,
                                double* d_best_scale) {
    if (d_pdfs && check_kde(h)) return 2;
    GenParams P = {};
    fill_common(h, P, d_x, d_xe, d_xm, No, cfg);
    P.stage = ST_FIT_PREDICT;
    P.objsel = d_objsel; P.Nsel = Nsel;
    P.kde = make_kde(h, cfg);
    P.pdfs = d_pdfs; P.lmap = d_lmap; P.levid = d_levid;
    P.best_idx = d_best_idx; P.best_chi2 = d_best_chi2; P.best_scale = d_best_scale;
    int64_t items = d_objsel ? Nsel : No;
    h->stats.pairs_fp64 += items * h->Nm;
    return launch_generic(h, P, items, true);
}

// fit_predict of every object against ITS OWN list of models (kNN union): likelihood + KDE in one pass, nothing but the
// PDFs (and lmap / levid) leaves the kernel.  knn.py:827-872 with save_fits=False.
int fzb_generic_gather_fit_predict_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                                       int64_t W, const int64_t* d_neighbors, const int64_t* d_nneighbors,
                                       const FzbConfig& cfg, double* d_pdfs, double* d_lmap, double* d_levid) {
    if (check_kde(h)) return 2;
    GenParams P = {};
    fill_common(h, P, d_x, d_xe, d_xm, No, cfg);
    P.stage = ST_FIT_PREDICT;
    P.W = W; P.nbr = d_neighbors; P.nnbr = d_nneighbors;
    P.kde = make_kde(h, cfg);
    P.pdfs = d_pdfs; P.lmap = d_lmap; P.levid = d_levid;
    h->stats.pairs_fp64 += No * W;
    return launch_generic(h, P, No, true);
}

int fzb_generic_predict_logwt_dev(fzb_context* h, const double* d_logwt, int64_t No, int64_t W,
                                  const int64_t* d_neighbors, const int64_t* d_nneighbors, const FzbConfig& cfg,
                                  double* d_pdfs, double* d_lmap, double* d_levid) {
    if (check_kde(h)) return 2;
    GenParams P = {};
    P.stage = ST_LOGWT;
    P.No = No; P.Nm = h->Nm; P.Nf = h->Nf; P.W = W;
    P.logwt = d_logwt; P.nbr = d_neighbors; P.nnbr = d_nneighbors;
    P.kde = make_kde(h, cfg);
    P.pdfs = d_pdfs; P.lmap = d_lmap; P.levid = d_levid;
    return launch_generic(h, P, No, false);
}

int fzb_generic_shard_pass1_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                                const int32_t* d_objsel, int64_t Nsel, const FzbConfig& cfg, double* d_pmax,
                                double* d_psum, int64_t* d_pbest) {
    GenParams P = {};
    fill_common(h, P, d_x, d_xe, d_xm, No, cfg);
    P.stage = ST_PASS1;
    P.objsel = d_objsel; P.Nsel = Nsel;
    P.pmax = d_pmax; P.psum = d_psum; P.pbest = d_pbest;
    int64_t items = d_objsel ? Nsel : No;
    h->stats.pairs_fp64 += items * h->Nm;
    return launch_generic(h, P, items, true);
}

int fzb_generic_shard_pass2_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                                const int32_t* d_objsel, int64_t Nsel, const FzbConfig& cfg, const double* d_lmap,
                                const double* d_levid, double* d_pdf_partial) {
    if (check_kde(h)) return 2;
    FZB_CHECK(cfg.use_wt_thresh || !cfg.use_cdf_thresh, "the CDF threshold rule needs all models on one device");
    GenParams P = {};
    fill_common(h, P, d_x, d_xe, d_xm, No, cfg);
    P.stage = ST_PASS2;
    P.objsel = d_objsel; P.Nsel = Nsel;
    P.kde = make_kde(h, cfg);
    P.pdfs = d_pdf_partial; P.pdfs32 = h->shard_out32; P.g_lmap = d_lmap; P.g_levid = d_levid;
    int64_t items = d_objsel ? Nsel : No;
    h->stats.pairs_fp64 += items * h->Nm;
    return launch_generic(h, P, items, true);
}
