// Float64 per-pair likelihood arithmetic in the reference's operation order
// (frankenz/pdf.py:27-235).  Shared by the generic path and the fp32 path's exact re-evaluation
// of each object's best model.
#pragma once
#include <math_constants.h>

namespace fzb64 {

constexpr double kLn2 = 0.69314718055994530942;
constexpr double kLn2Pi = 1.83787706640934548356;  // ln(2*pi)

__device__ __forceinline__ double xlogy_d(double a, double c) {
    // scipy.special.xlogy: 0 where a == 0 and c is not NaN
    if (a == 0.0 && !isnan(c)) return 0.0;
    return a * log(c);
}
__device__ __forceinline__ double chi2_logpdf(double chi2, double a) {
    // pdf.py:93 / :229
    return xlogy_d(a - 1.0, chi2) - (chi2 / 2.0) - lgamma(a) - (kLn2 * a);
}

struct PairState {
    double ndim, chi2, lnl, scale, shape;
};

// first evaluation of a pair: pdf.py:76-98 (fixed scale) or :171-194 (free scale)
__device__ __forceinline__ void pair_first(const double* sx, const double* sxe, const double* sxm,
                                           const double* __restrict__ m, const double* __restrict__ me,
                                           const double* __restrict__ mm, int Nf, int free_scale, int ime,
                                           PairState& st) {
    double ndim = 0.0, slv = 0.0;
    if (!free_scale) {
        double chi2 = 0.0;
        for (int b = 0; b < Nf; ++b) {
            double e = me[b];
            double var = sxe[b] * sxe[b] + (ime ? 0.0 : e * e);
            double msk = sxm[b] * mm[b];
            ndim += msk;
            double r = sx[b] - m[b];
            chi2 += msk * (r * r) / var;
            slv += log(var);
        }
        st.ndim = ndim;
        st.chi2 = chi2;
        st.scale = 1.0;
        st.shape = CUDART_NAN;
        double l = -0.5 * chi2;
        l += -0.5 * (ndim * kLn2Pi + slv);
        st.lnl = l;
        return;
    }
    double inter = 0.0, shape = 0.0;
    for (int b = 0; b < Nf; ++b) {
        double e = me[b];
        double var = sxe[b] * sxe[b] + (ime ? 0.0 : e * e);
        double msk = sxm[b] * mm[b];
        ndim += msk;
        inter += (msk * m[b] * sx[b]) / var;
        shape += (msk * (m[b] * m[b])) / var;
        slv += log(var);
    }
    double scale = inter / shape;
    double chi2 = 0.0;
    for (int b = 0; b < Nf; ++b) {
        double e = me[b];
        double var = sxe[b] * sxe[b] + (ime ? 0.0 : e * e);
        double msk = sxm[b] * mm[b];
        double r = sx[b] - scale * m[b];
        chi2 += msk * (r * r) / var;
    }
    st.ndim = ndim;
    st.chi2 = chi2;
    st.scale = scale;
    st.shape = shape;
    double l = -0.5 * chi2;
    l += -0.5 * (ndim * kLn2Pi + slv);
    st.lnl = l;
}

// one refinement of the iterated free-scale mode: pdf.py:200-216
__device__ __forceinline__ void pair_refine(const double* sx, const double* sxe, const double* sxm,
                                            const double* __restrict__ m, const double* __restrict__ me,
                                            const double* __restrict__ mm, int Nf, double ndim, double scale_prev,
                                            double& scale_new, double& chi2_new, double& lnl_new, double& shape_new) {
    double inter = 0.0, shape = 0.0, slv = 0.0;
    for (int b = 0; b < Nf; ++b) {
        double se = scale_prev * me[b];
        double var = sxe[b] * sxe[b] + se * se;
        double msk = sxm[b] * mm[b];
        inter += (msk * m[b] * sx[b]) / var;
        shape += (msk * (m[b] * m[b])) / var;
        slv += log(var);
    }
    double sc = inter / shape;
    double chi2 = 0.0;
    for (int b = 0; b < Nf; ++b) {
        double se = scale_prev * me[b];
        double var = sxe[b] * sxe[b] + se * se;
        double msk = sxm[b] * mm[b];
        double r = sx[b] - sc * m[b];
        chi2 += msk * (r * r) / var;
    }
    double l = -0.5 * chi2;
    l += -0.5 * (ndim * kLn2Pi + slv);
    scale_new = sc;
    chi2_new = chi2;
    lnl_new = l;
    shape_new = shape;
}


}  // namespace fzb64
