// Host side of the tensor-core sweep (kernel: fzb_sweep_tc.cuh): tile building and launch dispatch.  Its own
// translation unit so that the tcgen05 kernels compile beside the packed-FP32 / float64 sweeps of fzb_fast.cu.
#include <type_traits>

#include "fzb_sweep_common.cuh"

namespace {
using namespace fzbsweep;
#include "fzb_sweep_tc.cuh"

template <int NF, bool DP, int PASS, bool LIN, bool MLO, int FUSE = 0>
int launch_tc_t(fzb_context* h, const SweepParams& P, dim3 grid, bool prior, const unsigned char* tiles) {
    if (!tiles) tiles = h->fast.tiles_tc.as<unsigned char>();
    if (prior) {
        auto kern = k_sweep_tc<NF, DP, true, PASS, LIN, MLO, FUSE>;
        FZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem(NF, MLO)));
        kern<<<grid, TC_THREADS, tc_smem(NF, MLO), h->stream>>>(P, tiles, 128u, 256u);
    } else {
        auto kern = k_sweep_tc<NF, DP, false, PASS, LIN, MLO, FUSE>;
        FZB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem(NF, MLO)));
        kern<<<grid, TC_THREADS, tc_smem(NF, MLO), h->stream>>>(P, tiles, 128u, 256u);
    }
    fzb_count_launch(h);
    FZB_CUDA(cudaGetLastError());
    return 0;
}

template <int NF, bool DP, bool LIN, bool MLO>
int launch_tc_p(fzb_context* h, const SweepParams& P, dim3 grid, bool prior, int pass, const unsigned char* tiles) {
    return pass == 1 ? launch_tc_t<NF, DP, 1, LIN, MLO>(h, P, grid, prior, tiles) : launch_tc_t<NF, DP, 2, LIN, MLO>(h, P, grid, prior, tiles);
}

template <bool MLO>
int launch_tc_m(fzb_context* h, const SweepParams& P, dim3 grid, int nf, bool dp, int pass, bool lin, const unsigned char* tiles,
                int fuse) {
    const bool prior = P.has_prior != 0;
    if (fuse) {
        if (nf == 5 && lin && dp && pass == 1)
            return fuse == 1 ? launch_tc_t<5, true, 1, true, MLO, 1>(h, P, grid, prior, tiles)
                             : launch_tc_t<5, true, 1, true, MLO, 2>(h, P, grid, prior, tiles);
        fzb_set_error("tensor-core sweep: the fused pass needs the linear-domain form");
        return 2;
    }
    if (nf == 5) {
        if (lin && dp) return launch_tc_p<5, true, true, MLO>(h, P, grid, prior, pass, tiles);
        if (dp) return launch_tc_p<5, true, false, MLO>(h, P, grid, prior, pass, tiles);
        return launch_tc_p<5, false, false, MLO>(h, P, grid, prior, pass, tiles);
    }
    if (nf == 4) return dp ? launch_tc_p<4, true, false, MLO>(h, P, grid, prior, pass, tiles) : launch_tc_p<4, false, false, MLO>(h, P, grid, prior, pass, tiles);
    if (nf == 6) return dp ? launch_tc_p<6, true, false, MLO>(h, P, grid, prior, pass, tiles) : launch_tc_p<6, false, false, MLO>(h, P, grid, prior, pass, tiles);
    fzb_set_error("tensor-core sweep: unsupported filter count %d", nf);
    return 2;
}

}  // namespace

int fzb_tc_tile_objects() { return TC_OBJS; }
int fzb_tc_split() { return TC_SPLIT; }

// lin: linear-domain form, valid when every object handled by the fp32 pass has (dof/2 - 1) = 1 (Nf = 5, dim_prior)
int fzb_launch_sweep_tc(fzb_context* h, const SweepParams& P, dim3 grid, int nf, bool dp, int pass, bool lin, bool mlo,
                        const unsigned char* tiles, int fuse) {
    return mlo ? launch_tc_m<true>(h, P, grid, nf, dp, pass, lin, tiles, fuse) : launch_tc_m<false>(h, P, grid, nf, dp, pass, lin, tiles, fuse);
}

// Tile sets of the sorted model order: the full one and a coarse one (every FZB_TC_COARSE-th model: the pre-pass of the
// fused sweep, which only needs a lower bound of every object's maximum).  A model grid that is not fp32-representable
// (mlo) gets both sets twice: with the float64 remainder of the fluxes and without it (the faint objects' sweep, whose
// likelihoods do not feel the 2^-25 relative rounding of the models: fzb_fast.cu, FZB_TC_MLO_SNR).
int fzb_build_tiles_tc(fzb_context* h, const double* lnprior, const int32_t* bins, const float* invnorm, bool mlo) {
    FastModels& F = h->fast;
    const int nf = F.nf;
    for (int v = 0; v < (mlo ? 4 : 2); ++v) {
        const bool coarse = (v & 1) != 0, with_lo = mlo && v < 2;
        const int stride = coarse ? FZB_TC_COARSE : 1;
        const int64_t nm = (F.nm + stride - 1) / stride;
        const int64_t ntile = (nm + TC_TM - 1) / TC_TM;
        const size_t bytes = (size_t)ntile * tc_tile_bytes(nf, with_lo);
        DevBuf& dst = v == 0 ? F.tiles_tc : v == 1 ? F.tiles_tc_coarse : v == 2 ? F.tiles_tc_f32 : F.tiles_tc_f32_coarse;
        if (dst.reserve(bytes + 64)) return 1;
        FZB_CUDA(cudaMemsetAsync(dst.p, 0, bytes, h->stream));
        TcRecParams T = {};
        T.m = h->models.as<double>(); T.lnprior = lnprior; T.perm = F.perm.as<int32_t>(); T.bins = bins; T.invnorm = invnorm;
        T.nm = nm; T.Nf = nf; T.mlo = with_lo ? 1 : 0; T.tiles = dst.as<unsigned char>(); T.stride = stride;
        k_build_tiles_tc<<<(unsigned)((nm + 255) / 256), 256, 0, h->stream>>>(T);
        fzb_count_launch(h);
        FZB_CUDA(cudaGetLastError());
    }
    F.nm_coarse = (F.nm + FZB_TC_COARSE - 1) / FZB_TC_COARSE;
    F.tc_f32_valid = mlo;
    F.tc_valid = true;
    return 0;
}
