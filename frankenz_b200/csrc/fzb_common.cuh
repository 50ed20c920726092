// Shared declarations of libfzb200: context, error handling, launch bookkeeping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/frankenz_b200.h"

#define FZB_MAXF 64          // generic fp64 path: filters per object
#define FZB_FAST_MAXF 8      // register-tiled fp32 path: filters per object
#define FZB_MAX_NGRID 16384  // PDF grid points held in shared memory (fp64)

void fzb_set_error(const char* fmt, ...);

#define FZB_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            fzb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

#define FZB_CHECK(cond, ...)          \
    do {                              \
        if (!(cond)) {                \
            fzb_set_error(__VA_ARGS__); \
            return 2;                 \
        }                             \
    } while (0)

// Growable device buffer (never shrinks; freed with the context).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            fzb_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return 1;
        }
        cap = bytes;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

enum { FZB_KDE_NONE = 0, FZB_KDE_DICT = 1, FZB_KDE_GRID = 2 };

// Per-model constants of the fp32 fast path, laid out for bulk (TMA) staging into shared memory.
struct FastModels {
    bool valid = false;
    int nf = 0;
    int64_t nm = 0;        // models
    int64_t nm_pad = 0;    // padded to a multiple of the tile
    int rec = 0;           // floats per model record
    DevBuf recs;           // [nm_pad][rec] float: see fzb_fast.cu for the record layout
    DevBuf recs_coarse;    // the same records for every FZB_TC_COARSE-th model of the sorted order (pre-pass of the fused sweep)
    DevBuf recs64;         // [nm][rec64] double records of the float64 sweep
    DevBuf tiles_tc;       // tensor-core sweep: 256-model tiles (MMA operand + packed pairs + tails), fzb_sweep_tc.cuh
    DevBuf tiles_tc_coarse; // the same for every FZB_TC_COARSE-th model of the sorted order (pre-pass of the fused sweep)
    DevBuf tiles_tc_f32, tiles_tc_f32_coarse;   // models not fp32-representable: the same sets without the float64 remainder
    bool tc_f32_valid = false;
    int64_t nm_coarse = 0;
    bool tc_valid = false;
    DevBuf fuse;           // fused sweep: per-thread records, counts, seeds, object flags
    DevBuf aux64;          // per-object float64 pass-2 inputs
    DevBuf perm;           // int32 [nm_pad]: sorted position -> original model index (-1 = padding)
    DevBuf bins;           // int32 [nm_pad]: KDE histogram bin (slot*Ng + pos) of each sorted model, -1 = none
    DevBuf invnorm;        // float [nm_pad]: 1 / kernel normalisation of each sorted model
    DevBuf live;           // pass-2 pruning: live bits of the tensor-core sweep, [model tile x half][object] uint16
    DevBuf live64;         // pass-2 pruning of the float64 sweep: one bit per (model tile, object), uint16 [tile][object]
    DevBuf tmask;          // pass-2 pruning: live bits per (model tile, pass-2 CTA), see fzb_tile_masks
    DevBuf sortbuf;        // keys / values / temporary storage of the sort of the pass-2 object list
    DevBuf cutlist;        // weights recorded at the wt_thresh cut by pass 2 (CutRecord), re-decided in float64
    int nslot = 0;         // distinct dictionary widths in use
    std::vector<int32_t> slot_sidx;  // slot -> dictionary index
    DevBuf d_slot_sidx;
};

struct fzb_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    // pipelined host API: second stream (copy engine), double-buffered device + pinned staging for the PDFs
    cudaStream_t stream2 = nullptr;
    DevBuf pdf_dev[2];
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_cap[2] = {0, 0};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_chunk[2] = {nullptr, nullptr};   // all pieces of chunk c (c & 1) have reached pinned memory

    // model set (fp64 originals, row-major Nm x Nf)
    int64_t Nm = 0;
    int Nf = 0;
    DevBuf models, models_err, models_mask, lnprior;
    bool has_lnprior = false;
    std::vector<double> h_lnprior;   // host copy: detect an unchanged prior
    // object-conditioned tabulated prior: table [nbins][Nm]; bins of the objects of the next call
    DevBuf prior_table, prior_bins;
    int prior_nbins = 0;
    int64_t prior_bins_n = 0;      // 0: no bins pending
    int64_t prior_o0 = 0;          // first object of the chunk being processed (host API chunking)
    bool mask_all_one = false;     // every model mask entry == 1
    bool mask_binary = false;      // every model mask entry is 0 or 1
    bool err_all_zero = false;     // every model error == 0
    bool models_finite = false;
    bool models_f32_exact = false; // every model flux is exactly representable in float32

    // KDE tables
    int kde_mode = FZB_KDE_NONE;
    int Ng = 0;
    int Ndict = 0;
    std::vector<int32_t> h_widths;
    std::vector<int64_t> h_koff;
    std::vector<double> h_kernels, h_kcdf;  // host copies: detect an unchanged dictionary
    DevBuf widths, koff, kernels, kcdf;     // dictionary
    DevBuf yidx, ysidx;                     // int64 per model
    std::vector<int64_t> h_yidx, h_ysidx;
    bool labels_dict_set = false;
    int64_t labels_bad = 0;                 // labels whose kernel misses the grid / is malformed (error only if selected)
    DevBuf kde_err;                         // {flag, model} raised by kde_add_dict when such a label is selected
    DevBuf grid, y, ystd, lowers, uppers;   // exact-Gaussian KDE
    bool labels_grid_set = false;

    // scratch
    DevBuf rows;            // generic kernels: per-CTA row state
    DevBuf obj_in[3];       // staged inputs
    DevBuf out_f64[8];      // staged outputs
    DevBuf out_i64[2];
    DevBuf misc[8];

    // pdfs_summarize (fzb_summarize.cu): grid, loss matrix, CDFs of the last call, outputs, staging
    DevBuf summ[6];
    int64_t summ_No = 0;
    int summ_Ng = 0;

    FastModels fast;
    bool fast_dirty = true;
    int fast_mode = -1;
    bool fast_packed = true;
    int fast_wmax = 0;
    int fast_Ngpad = 0;
    int fast_ns64 = 0;             // model splits of the last float64 sweep
    // state a model-sharded pass 1 leaves for pass 2
    bool shard_valid = false;
    int64_t shard_No = 0;
    int shard_counts[4] = {0, 0, 0, 0};
    int shard_cfg_key = 0;
    bool shard_lin = false;        // pass 1 ran the linear-domain tensor-core sweep
    float* shard_out32 = nullptr;  // pass 2 writes its PDF partials here in fp32 (set for the duration of the call)

    // kNN
    DevBuf knn_feats;       // float32 K x Nm x Nf (+ 64 B pad)
    DevBuf knn_cand, knn_redo;
    DevBuf knn_tiles;       // tensor-core scan: centred, tf32-split row tiles of every tree (fzb_knn_tc.cu)
    DevBuf knn_aux;         // double: centre [FZB_FAST_MAXF], max |f'|^2 per tree [K]
    DevBuf knn_scan;        // filter scan: interleaved, value-duplicated copy of the features (fzb_knn.cu)
    DevBuf knn_buf;         // filter scan: per-(query, tree) row buffers, thresholds, counts
    DevBuf knn_centre;      // filter scan, dot form: mean feature (double[Nf])
    int knn_form = 0;       // fp32 distance form of the scan copy (fzb_knn.cu: KS_DIFF / KS_DOT)
    bool knn_scan_valid = false;
    bool knn_share_off = false;   // filter scan: the threshold of tree 0 does not serve the other trees of this set
    int64_t knn_m = 0, knn_Ns = 0;   // interleave stride / rows of the scan copy
    bool knn_tc_valid = false;
    int64_t knn_ntile = 0;
    int knn_K = 0;
    int64_t knn_stride = 0;  // floats between consecutive trees (16-byte aligned, >= Nm*Nf + 3)
    int64_t knn_Nm = 0;
    int knn_Nf = 0;

    // population likelihood (fzb_samplers.cu): PDFs resident on the device (own copy or a caller's device pointer)
    DevBuf nz_pdfs, nz_buf;
    const double* nz_pdfs_ptr = nullptr;
    int64_t nz_No = 0;
    int nz_Ng = 0;

    FzbStats stats = {};
};

static inline void fzb_count_launch(fzb_context* h, int64_t n = 1) { h->stats.kernel_launches += n; }

// ---- generic fp64 path (fzb_generic.cu) ------------------------------------------------------
int fzb_generic_fit_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                        const FzbConfig& cfg, double* d_lnprior, double* d_lnlike, double* d_lnprob,
                        int64_t* d_ndim, double* d_chi2, double* d_scale, double* d_scale_err);
int fzb_generic_fit_predict_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm,
                                int64_t No, const int32_t* d_objsel, int64_t Nsel, const FzbConfig& cfg,
                                double* d_pdfs, double* d_lmap, double* d_levid, int64_t* d_best_idx,
                                double* d_best_chi2, double* d_best_scale);
int fzb_generic_predict_logwt_dev(fzb_context* h, const double* d_logwt, int64_t No, int64_t W,
                                  const int64_t* d_neighbors, const int64_t* d_nneighbors, const FzbConfig& cfg,
                                  double* d_pdfs, double* d_lmap, double* d_levid);
int fzb_generic_gather_fit_predict_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                                       int64_t W, const int64_t* d_neighbors, const int64_t* d_nneighbors,
                                       const FzbConfig& cfg, double* d_pdfs, double* d_lmap, double* d_levid);
int fzb_generic_shard_pass1_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm,
                                int64_t No, const int32_t* d_objsel, int64_t Nsel, const FzbConfig& cfg,
                                double* d_pmax, double* d_psum, int64_t* d_pbest);
int fzb_generic_shard_pass2_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm,
                                int64_t No, const int32_t* d_objsel, int64_t Nsel, const FzbConfig& cfg,
                                const double* d_lmap, const double* d_levid, double* d_pdf_partial);
int fzb_generic_gather_fit_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm,
                               int64_t No, int64_t W, const int64_t* d_neighbors, const int64_t* d_nneighbors,
                               const FzbConfig& cfg, double* d_lnprior, double* d_lnlike, double* d_lnprob,
                               int64_t* d_ndim, double* d_chi2, double* d_scale, double* d_scale_err);

// ---- fp32 fast path (fzb_fast.cu) ------------------------------------------------------------
bool fzb_fast_supported(const fzb_context* h, const FzbConfig& cfg);
// shard_mode 0: whole fit_predict.  1: model-sharded pass 1 (d_lmap = partial max, d_psum = partial sum, d_best_idx =
// partial arg-max; state for pass 2 stays in the context).  2: model-sharded pass 2 (d_glmap = global lmap in,
// d_pdfs = un-normalised PDF partial out).
int fzb_fast_fit_predict_dev(fzb_context* h, const double* d_x, const double* d_xe, const double* d_xm, int64_t No,
                             const FzbConfig& cfg, double* d_pdfs, double* d_lmap, double* d_levid,
                             int64_t* d_best_idx, double* d_best_chi2, double* d_best_scale, int shard_mode = 0,
                             double* d_psum = nullptr, const double* d_glmap = nullptr);

// pass-2 object order of the tensor-core sweep (fzb_prune.cu)
int fzb_tile_masks(fzb_context* h, const unsigned short* live, int64_t ntiles, int64_t No_pad, const int32_t* list, int64_t n,
                   int tile_objs, unsigned int** out);
int fzb_sort_by_live_bits(fzb_context* h, const unsigned short* live, int64_t nrows, int64_t No_pad, int32_t* list,
                          int64_t n, int bits_per_row = 16);

// ---- PDF summaries (fzb_summarize.cu) ---------------------------------------------------------
int fzb_summarize_impl(fzb_context* h, const double* pdfs, const double* pgrid, const double* loss, const double* urand,
                       int64_t No, int32_t Ng, int32_t renormalize, double* rowsum, double* est, double* sd, double* risk,
                       double* quant, double* mc);
int fzb_conf_impl(fzb_context* h, const double* points, const double* widths, int64_t No, double* conf);
int fzb_summarize_tables(fzb_context* h, const double* pgrid, const double* loss, const double* urand, int64_t No,
                         int32_t Ng);
int fzb_summarize_rows_dev(fzb_context* h, const double* d_pdfs, int64_t n, int64_t o0, int64_t Ntot, int32_t Ng,
                           int32_t renormalize, double wfac);
int fzb_summarize_download(fzb_context* h, int64_t No, double* est, double* sd, double* conf, double* risk, double* quant,
                           double* mc);

// ---- model-sharded merge kernels (fzb_shard.cu) --------------------------------------------------
int fzb_shard_add_offset_launch(fzb_context* h, int64_t* d_best, int64_t No, int64_t offset);
int fzb_shard_merge_launch(fzb_context* h, const double* d_gathered, int world, int64_t No, double* d_lmap,
                           double* d_levid, int64_t* d_best);
int fzb_shard_normalise_launch(fzb_context* h, const float* d_rows, int64_t n, int Ng, double* d_pdfs);

// ---- population likelihood (fzb_samplers.cu) ---------------------------------------------------
int fzb_nz_loglike_impl(fzb_context* h, const double* d_pdfs, int64_t No, int Ng, const double* nz_host, int pa, int pb,
                        double step, double* lnlike, double* overlap_host);

// ---- kNN (fzb_knn.cu) -------------------------------------------------------------------------
int fzb_knn_query_dev(fzb_context* h, const double* d_q, int64_t No, int k, double p, int64_t* d_idx, double* d_dist);
int fzb_knn_scan_build(fzb_context* h);
int fzb_knn_tc_build(fzb_context* h);
int fzb_knn_tc_kcmax();
int fzb_knn_tc_lists();
int fzb_knn_tc_scan(fzb_context* h, const double* d_q, int64_t No, int KC, int nsp, int tiles_per_split, float* cand_d,
                    int* cand_i);
int fzb_knn_union_dev(fzb_context* h, const int64_t* d_idx, int64_t No, int Kk, int64_t* d_neighbors,
                      int64_t* d_nneighbors);
