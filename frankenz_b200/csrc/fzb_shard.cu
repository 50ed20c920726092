// Model-sharded mode: the three per-object merges that sit between the sharded passes and
// the NCCL collectives.  Each rank scores the objects against ITS slice of the models; what crosses NVLink is
//   after pass 1: one all-gather of the packed partials (max lnprob, sum exp(lnprob - max), arg-max): 24 B / object / rank
//   after pass 2: one reduce-scatter (sum) of the un-normalised PDF partials in fp32: Ngrid x 4 B / object / rank
// k_shard_merge turns the gathered partials into the global (lmap, levid, best) every rank needs for pass 2
// (bruteforce.py:359: max and logsumexp over ALL models); k_shard_normalise finishes the rows a rank owns after the
// reduce-scatter (bruteforce.py:370).
#include <math_constants.h>

#include "fzb_common.cuh"

namespace {

__global__ void k_shard_add_offset(int64_t* best, int64_t No, int64_t offset) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < No) best[o] += offset;
}

// gathered: [world][3][No] doubles = (pmax, psum, bit pattern of the int64 global arg-max) of every rank.
// NaN partials poison the object (numpy max / logsumexp of bruteforce.py:359); a rank whose partial maximum is -inf
// contributes nothing; exact ties of the maximum go to the lowest global model index.
__global__ void k_shard_merge(const double* __restrict__ g, int world, int64_t No, double* __restrict__ lmap,
                              double* __restrict__ levid, int64_t* __restrict__ best) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= No) return;
    bool poisoned = false;
    double gmax = -CUDART_INF;
    for (int r = 0; r < world; ++r) {
        const double m = g[((size_t)r * 3 + 0) * No + o], s = g[((size_t)r * 3 + 1) * No + o];
        if (isnan(m) || isnan(s)) poisoned = true;
        else if (m > gmax) gmax = m;
    }
    double S = 0.0;
    long long b = 0x7fffffffffffffffll;
    for (int r = 0; r < world; ++r) {
        double m = g[((size_t)r * 3 + 0) * No + o];
        const double s = g[((size_t)r * 3 + 1) * No + o];
        if (isnan(m) || isnan(s)) m = -CUDART_INF;
        if (isfinite(m)) S += s * exp(m - gmax);
        if (m == gmax) {
            const long long cand = __double_as_longlong(g[((size_t)r * 3 + 2) * No + o]);
            if (cand < b) b = cand;
        }
    }
    double le = isinf(gmax) ? gmax : gmax + log(S);
    double lm = gmax;
    if (poisoned) { lm = CUDART_NAN; le = CUDART_NAN; }
    lmap[o] = lm;
    levid[o] = le;
    if (best) best[o] = b;
}

// one warp per row: float64 sum of the reduced fp32 partials, then the division of bruteforce.py:370
__global__ void k_shard_normalise(const float* __restrict__ rows, int64_t n, int Ng, double* __restrict__ pdfs) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const float* in = rows + (size_t)r * Ng;
    double tot = 0.0;
    for (int g = lane; g < Ng; g += 32) tot += (double)in[g];
    for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
    double* out = pdfs + (size_t)r * Ng;
    for (int g = lane; g < Ng; g += 32) out[g] = (double)in[g] / tot;
}

}  // namespace

int fzb_shard_add_offset_launch(fzb_context* h, int64_t* d_best, int64_t No, int64_t offset) {
    if (No <= 0 || offset == 0) return 0;
    k_shard_add_offset<<<(unsigned)((No + 255) / 256), 256, 0, h->stream>>>(d_best, No, offset);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

int fzb_shard_merge_launch(fzb_context* h, const double* d_gathered, int world, int64_t No, double* d_lmap,
                           double* d_levid, int64_t* d_best) {
    if (No <= 0) return 0;
    k_shard_merge<<<(unsigned)((No + 255) / 256), 256, 0, h->stream>>>(d_gathered, world, No, d_lmap, d_levid, d_best);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

int fzb_shard_normalise_launch(fzb_context* h, const float* d_rows, int64_t n, int Ng, double* d_pdfs) {
    if (n <= 0) return 0;
    const int wpb = 8;
    k_shard_normalise<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, h->stream>>>(d_rows, n, Ng, d_pdfs);
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}
