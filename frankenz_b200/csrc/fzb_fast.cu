// fp32 register-tiled path (placeholder until the kernels land: reports "unsupported").
#include "fzb_common.cuh"

bool fzb_fast_supported(const fzb_context*, const FzbConfig&) { return false; }
int fzb_fast_prepare(fzb_context*) { return 0; }
int fzb_fast_fit_predict_dev(fzb_context*, const double*, const double*, const double*, int64_t, const FzbConfig&,
                             double*, double*, double*, int64_t*, double*, double*) {
    fzb_set_error("fp32 path not built");
    return 2;
}
