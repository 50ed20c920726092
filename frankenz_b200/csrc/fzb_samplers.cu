// Population-level likelihood of a redshift distribution given the per-object PDFs the path produces:
// frankenz/samplers.py:24-76 (`loglike_nz`), SURVEY.md section 8f rank 4.
//     overlap_i = sum_g pdfs[i, g] nz[g]  (+ pair_step (pdfs[i, a] - pdfs[i, b]))      a GEMV over (Nobs x Nbins)
//     lnlike    = sum_i log(overlap_i)
// It sits inside the MCMC loops of population_sampler / hierarchical_sampler (samplers.py:196-199, 460-470), which call
// it thousands of times on the SAME PDFs: the PDFs stay resident in HBM (they can be handed over as a device pointer
// straight from fzb_fit_predict_dev), every call streams them once.  HBM-bound: Nobs x Nbins x 8 bytes per call.
#include <algorithm>

#include <math_constants.h>

#include "fzb_common.cuh"

namespace {

// one warp per row; nz in shared memory
__global__ void __launch_bounds__(256) k_nz_overlap(const double* __restrict__ pdfs, const double* __restrict__ nz, int64_t No,
                                                    int Ng, int pa, int pb, double step, double* __restrict__ overlap) {
    extern __shared__ double s_nz[];
    for (int g = threadIdx.x; g < Ng; g += blockDim.x) s_nz[g] = nz[g];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = wid; i < No; i += nw) {
        const double* row = pdfs + (size_t)i * Ng;
        double acc = 0.0;
        for (int g = lane; g < Ng; g += 32) acc = fma(row[g], s_nz[g], acc);
        for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
        if (lane == 0) {
            if (pa >= 0) acc += step * (row[pa] - row[pb]);
            overlap[i] = acc;
        }
    }
}

// deterministic sum of log(overlap): fixed assignment of rows to threads, tree over the block, then over blocks
__global__ void __launch_bounds__(1024) k_nz_logsum(const double* __restrict__ overlap, int64_t No, double* __restrict__ part) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < No; i += (int64_t)gridDim.x * blockDim.x)
        acc += log(overlap[i]);
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = red[threadIdx.x];
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (threadIdx.x == 0) part[blockIdx.x] = v;
    }
}

}  // namespace

int fzb_nz_loglike_impl(fzb_context* h, const double* d_pdfs, int64_t No, int Ng, const double* nz_host, int pa, int pb,
                        double step, double* lnlike, double* overlap_host) {
    const int nblk = 64;
    if (h->nz_buf.reserve((size_t)(Ng + No + nblk) * 8 + 64)) return 1;
    double* d_nz = h->nz_buf.as<double>();
    double* d_ov = d_nz + Ng;
    double* d_part = d_ov + No;
    FZB_CUDA(cudaMemcpyAsync(d_nz, nz_host, (size_t)Ng * 8, cudaMemcpyHostToDevice, h->stream));
    FZB_CUDA(cudaEventRecord(h->ev[0], h->stream));
    const int64_t rows_per_block = 8;
    int64_t grid = std::min<int64_t>((No + rows_per_block - 1) / rows_per_block, (int64_t)h->sm_count * 16);
    k_nz_overlap<<<(unsigned)std::max<int64_t>(1, grid), 256, (size_t)Ng * 8, h->stream>>>(d_pdfs, d_nz, No, Ng, pa, pb, step, d_ov);
    k_nz_logsum<<<nblk, 1024, 0, h->stream>>>(d_ov, No, d_part);
    fzb_count_launch(h, 2);
    FZB_CUDA(cudaGetLastError());
    FZB_CUDA(cudaEventRecord(h->ev[1], h->stream));
    double part[64];
    FZB_CUDA(cudaMemcpyAsync(part, d_part, sizeof(part), cudaMemcpyDeviceToHost, h->stream));
    if (overlap_host) FZB_CUDA(cudaMemcpyAsync(overlap_host, d_ov, (size_t)No * 8, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    for (int i = 0; i < nblk; ++i) tot += part[i];
    *lnlike = tot;
    float ms = 0.f;
    FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->stats.ms_total = ms;
    return 0;
}
