// Pieces shared by the sweep kernels (fzb_fast.cu: packed FP32 and float64 sweeps; fzb_sweep_tc.cu: tensor-core sweep):
// mbarrier / TMA bulk-copy PTX helpers, packed f32x2 arithmetic, kernel parameter block.
#pragma once
#include <cfloat>

#include "fzb_common.cuh"

namespace fzbsweep {

// ---- weights at the wt_thresh cut -------------------------------------------------------------------------------------
// The selection wt > wt_thresh max(wt) (pdf.py:589-591) is a hard cut: a model whose fp32 weight lies within the fp32
// error of the cut could fall on the other side than in float64 and move wt_thresh / sum(wt) of the PDF.  Pass 2 of the
// sweeps therefore RECORDS every weight within ex_tol (relative) of the cut - (object, model, weight, fp32 decision), a
// few per thousand sub-batches - and k_exact_cut_fix (fzb_fast.cu) re-decides them with the float64 pair arithmetic of
// the reference against the exact maximum, correcting the histogram where the decision changes.  The hot loop only pays
// for the detection (one packed add and one 3-input minimum per two weights).
struct CutRecord {
    int obj;          // chunk-local object
    int model;        // sorted model position
    float weight;     // the fp32 weight (relative to the pass-2 maximum)
    int selected;     // the fp32 decision
};

constexpr int TM = 256;          // models per shared-memory tile
constexpr int NSTAGE = 2;
constexpr float kHalfLog2e = 0.7213475204444817f;

enum FastMode { FM_FS0 = 0, FM_FX0 = 1, FM_FX1 = 2 };

// ---- PTX helpers: mbarrier + TMA bulk copy ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- kernel parameters -------------------------------------------------------------------------
struct SweepParams {
    // objects, SoA [field][band][No_pad]
    const float* od;      // d_hi
    const float* ow;      // FS0/FX0: mask/err^2        FX1: err^2 (+inf where masked)
    const float* ox;      // FS0: d*w
    const float* odl;     // 2 * d_lo
    const float* oA;      // [No_pad] (dof/2 - 1) or 0
    const int32_t* obits; // [No_pad] band-mask bits of the object (model-mask variant)
    float Atab[16];       // model-mask variant: (dof/2 - 1) by pair dimensionality
    float Ktab[16];       //                     ndim-dependent constant of the ln-likelihood, log2 units
    int64_t No_pad;
    int64_t No;           // objects in this launch (pass 1) / entries of objlist (pass 2)
    // models
    const float* recs;    // [nm][REC]
    int64_t nm;
    int tiles_per_split;
    int has_prior;
    // pass 1 outputs: [nsplit][No_pad]
    double* pM;
    double* pS;
    int32_t* pbest;
    // pass 2
    const int32_t* objlist;
    const float* M2;      // [No_pad] final max (log2 units, without the per-object constant)
    const float* thr2;    // [No_pad] selection cut in the same units
    float* hist;          // [No_pad][hist_stride]
    int64_t hist_stride;
    // pass-2 pruning (tensor-core sweep): one bit per (object, 8-model sub-batch), set by pass 1 when a weight of the
    // sub-batch exceeds live_thr times the running maximum (a superset of the final wt_thresh selection, the running
    // maximum only grows); layout [model tile][half][No_pad] uint16, bit = 2 * chunk + k.  Null: no pruning.
    unsigned short* live;
    float live_thr;
    float live_lthr;                  // the same cut as a log2 difference (packed sweep: one bit per object and tile)
    const unsigned int* tmask;        // pass 2: [model tile][CTA] live bits OR-ed over the 128 objects of each M-tile (16 bits
                                      // each, both halves): chunks and tiles nobody needs are not multiplied / staged at all
    unsigned long long* pairs_done;   // pass 2: object-model pairs actually evaluated (statistics)
    CutRecord* ex_list;               // pass 2: weights within ex_tol of the cut (null: off), see CutRecord
    unsigned int* ex_count;
    unsigned int ex_cap;
    float ex_tol;                     // relative half-width of the band around the cut
    // fused single pass (k_sweep_tc<..., FUSE>): see the kernel's header comment
    const float* fz_M0;               // [No_pad] lower bound of the object's maximum (sweep units), -FLT_MAX: none
    float fz_thr;                     // wt_thresh
    float fz_lofac, fz_gfac;          // recorded band relative to the running cut: (lofac, gfac]
    float fz_mid, fz_half;            // (lofac + gfac) / 2, (gfac - lofac) / 2
    uint4* fz_rec;                    // [part][No_pad][fz_cap] x 3: {first model position, cut, eight weights, -, -}
    int* fz_cnt;                      // [part][No_pad] records of the thread (> fz_cap: overflow; -1: frame changed)
    int fz_cap;
};

__device__ __forceinline__ void record_cut(const SweepParams& P, int obj, int model, float weight, bool selected) {
    const unsigned int at = atomicAdd(P.ex_count, 1u);
    if (at < P.ex_cap) P.ex_list[at] = CutRecord{obj, model, weight, selected ? 1 : 0};
}


typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float a, float b) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo2(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi2(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}


}  // namespace fzbsweep

// ---- tensor-core sweep (fzb_sweep_tc.cu) ----------------------------------------------------------------------------
// lin: linear-domain form ((dof/2 - 1) = 1); mlo: the tiles carry the float64 remainder of the model fluxes
// tiles: null = the full tile set of the context; fuse (pass 1, lin): 1 = the single-pass variant, 2 = seeded pass 1
int fzb_launch_sweep_tc(fzb_context* h, const fzbsweep::SweepParams& P, dim3 grid, int nf, bool dp, int pass, bool lin,
                        bool mlo, const unsigned char* tiles = nullptr, int fuse = 0);
// build the 256-model tiles (MMA operand + packed pairs + KDE tails) from the sorted model order; sets h->fast.tc_valid
int fzb_build_tiles_tc(fzb_context* h, const double* lnprior, const int32_t* bins, const float* invnorm, bool mlo);
constexpr int FZB_TC_COARSE = 16;     // the coarse tile set (h->fast.tiles_tc_coarse) holds every 16th model of the sorted order
int fzb_tc_tile_objects();     // objects per CTA of the tensor-core sweep
int fzb_tc_split();            // partial results per object and model split
