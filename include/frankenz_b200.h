/*
 * frankenz_b200 -- C ABI of the B200-native brute-force photometric likelihood path.
 *
 * This header is the drop-in boundary: a maintainer of joshspeagle/frankenz would
 * bind exactly these entry points (ctypes stub shown in INTEGRATION.md) to replace
 * the per-object Python loops of
 *     frankenz/pdf.py:238-411          loglike / logprob
 *     frankenz/bruteforce.py:127-205   BruteForce._fit
 *     frankenz/bruteforce.py:303-372   BruteForce._predict
 *     frankenz/bruteforce.py:505-631   BruteForce._fit_predict
 *     frankenz/knn.py:158-188          NearestNeighbors._train_kdtrees
 *     frankenz/knn.py:281-388, 486-558, 722-874   NearestNeighbors._fit/_predict/_fit_predict
 *
 * Conventions
 *   - plain C types only; every array is C-contiguous; "host" pointers are ordinary
 *     (pageable or pinned) CPU memory, "dev" pointers are CUDA device memory on the
 *     handle's device.
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from fzb_last_error() (thread-local).
 *   - a handle owns one CUDA stream; a handle is thread-compatible, not thread-safe.
 *   - there is NO CPU fallback: every compute entry point fails if no CUDA device
 *     is usable.
 */
#ifndef FRANKENZ_B200_H
#define FRANKENZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fzb_context* fzb_handle;

/* Likelihood / KDE options.  Mirrors the keyword arguments of
 * frankenz/pdf.py:238-240 (loglike) and frankenz/pdf.py:444-445, 529-531 (KDE). */
typedef struct FzbConfig {
    int32_t free_scale;        /* pdf.py:313  */
    int32_t ignore_model_err;  /* pdf.py:76, :171, :197 */
    int32_t dim_prior;         /* pdf.py:90, :226 */
    int32_t track_scale;       /* return scale / scale_err (pdf.py:231-233) */
    double  ltol;              /* pdf.py:199, default 1e-4 */
    int32_t use_wt_thresh;     /* 1: select wt > wt_thresh*max(wt) (pdf.py:589-591) */
    int32_t use_cdf_thresh;    /* 1 (and use_wt_thresh==0): CDF rule (pdf.py:592-597) */
    double  wt_thresh;         /* default 1e-3 */
    double  cdf_thresh;        /* default 2e-4 */
    int32_t precision;         /* FZB_PREC_* */
    int32_t reserved;          /* bit 0 (fzb_predict_logwt only): return the un-normalised stack of kernels with weights
                                  exp(logwt - levid), as the module-level gauss_kde / gauss_kde_dict do (pdf.py:519-526) */
} FzbConfig;

enum {
    FZB_PREC_AUTO = 0,   /* fused fp32 scan + fp64 where the error bound requires it */
    FZB_PREC_FP64 = 1,   /* everything in float64, reference operation order          */
    FZB_PREC_FP32 = 2    /* force the fp32 path (benchmarking only)                    */
};

/* Optional (Ndata x Nmodel) outputs of a full fit; any pointer may be NULL.
 * Mirrors the arrays of frankenz/bruteforce.py:182-189. Host pointers. */
typedef struct FzbFitOut {
    double*  lnprior;
    double*  lnlike;
    double*  lnprob;
    int64_t* Ndim;
    double*  chi2;
    double*  scale;
    double*  scale_err;
} FzbFitOut;

/* Device-side counters of the last call (for bench.py's gpu_launches / roofline). */
typedef struct FzbStats {
    int64_t kernel_launches;   /* launches of this library's kernels in the last call   */
    int64_t pairs_fp32;        /* object-model pairs evaluated by fp32 kernels          */
    int64_t pairs_fp64;        /* object-model pairs evaluated by fp64 kernels          */
    int64_t objects_fp64;      /* objects routed to the fp64 path by the error bound    */
    double  ms_scan;           /* CUDA-event time of pass 1 (max / logsumexp / argmax); kNN calls: of the search */
    double  ms_accum;          /* CUDA-event time of pass 2 (weights -> histogram)       */
    double  ms_finish;         /* CUDA-event time of histogram (*) kernel + normalise    */
    double  ms_total;          /* CUDA-event time of all device work of the call        */
    int64_t sweep_kind;        /* fused path of the last call: 0 none, 1 packed FP32 sweep (k_sweep2), 2 tensor-core
                                  sweep (k_sweep_tc), 3 tensor-core sweep in its linear-domain form           */
    int64_t knn_redo;          /* kNN: (query, tree) searches the fp32 scan could not prove exact and the float64
                                  kernel re-did from scratch                                                    */
    int64_t pairs_pass2;       /* object-model pairs pass 2 actually evaluated (after sub-batch pruning)       */
    double  ms_summarize;      /* CUDA-event time of the fused PDF summaries (fzb_fit_predict_summarize)        */
    int64_t knn_tc;            /* 1: the last kNN search generated its candidates on the tensor cores           */
    int64_t cut_recorded;      /* pass 2: weights within the fp32 error of the wt_thresh cut, re-decided in float64   */
    int64_t cut_changed;       /* ... of which the float64 decision differed from the fp32 one                 */
    double  knn_tc_err;        /* largest error of a candidate's fp32 value seen by the float64 re-rank, in units of
                                  the scale of the exactness test: tensor-core scan (|q'|^2 + max |f'|^2), bound 4e-6;
                                  dot-form filter scan (|q'|^2 + |f'|^2), bound 2e-6                              */
    int64_t objects_fused;     /* fused single pass of the tensor-core sweep: objects whose histogram it completed (the others
                                  took the pruned pass 2)                                                         */
    int64_t knn_overflow;      /* kNN filter scan: (query, tree) searches whose row buffer overflowed (re-done by the
                                  float64 kernel, counted in knn_redo as well)                                  */
} FzbStats;

const char* fzb_last_error(void);
int fzb_version(void);
int fzb_device_count(int* count);

int fzb_create(int device, fzb_handle* out);
int fzb_destroy(fzb_handle h);
int fzb_synchronize(fzb_handle h);
int fzb_get_stats(fzb_handle h, FzbStats* out);

/* Roofline denominators measured on the handle's device (MEASURED_PEAKS.json carries only HBM and
 * bf16 tensor peaks): dependency-free FFMA and MUFU.EX2 loops over every SM, best of `reps`.
 * fp32_tflops counts an FMA as 2 flops; mufu_gops is special-function results per second / 1e9. */
int fzb_measure_peaks(fzb_handle h, int reps, double* fp32_tflops, double* mufu_gops);

/* Model set: replaces BruteForce.__init__ (bruteforce.py:36-64) / NearestNeighbors.__init__
 * storage (knn.py:89-101).  models/err/mask: host, (Nm x Nf) float64; mask values are
 * used multiplicatively exactly like the reference (pdf.py:82). */
int fzb_set_models(fzb_handle h, const double* models, const double* models_err,
                   const double* models_mask, int64_t Nm, int32_t Nf);

/* Built-in per-model ln-prior added to lnlike (north_star: replaces custom Python
 * lprob_func).  NULL => zeros (pdf.py:406). host, [Nm]. */
int fzb_set_lnprior(fzb_handle h, const double* lnprior, int64_t Nm);

/* Object-conditioned tabulated prior: lnprior of model j for an object in bin b is
 * table[b*Nm + j] (e.g. ln P(z_j, t_j | magnitude bin), the built-in replacement for demo 2's Python lprob_bpz,
 * demos/2 cell 69).  fzb_set_object_prior_bins gives the bin of every object of the NEXT fit / fit_predict /
 * knn_fit call (host int32 [No], values in [0, nbins)); it is consumed by that call.  NULL table / bins clear it. */
int fzb_set_lnprior_table(fzb_handle h, const double* table, int32_t nbins, int64_t Nm);
int fzb_set_object_prior_bins(fzb_handle h, const int32_t* bins, int64_t No);

/* Dictionary KDE tables: what PDFDict.__init__ tabulates (pdf.py:800-819).
 * widths[Ndict]; koff[Ndict+1] offsets into kernels/kcdf (each kernel has 2*w+1 entries). */
int fzb_set_kde_dict(fzb_handle h, int32_t Ngrid, int32_t Ndict, const int32_t* widths,
                     const int64_t* koff, const double* kernels, const double* kcdf);
/* Per-model dictionary indices: PDFDict.fit output (pdf.py:844-850). host, [Nm]. */
int fzb_set_labels_dict(fzb_handle h, const int64_t* y_idx, const int64_t* y_std_idx, int64_t Nm);

/* Exact-Gaussian KDE (pdf.py:444-526): grid[Ngrid], per-model label / sigma and the
 * clipped windows [lower, upper) computed as pdf.py:499-502. host. */
int fzb_set_kde_grid(fzb_handle h, const double* grid, int32_t Ngrid);
int fzb_set_labels_grid(fzb_handle h, const double* y, const double* y_std, const int64_t* lowers,
                        const int64_t* uppers, int64_t Nm);

/* loglike/logprob of No objects against all models, full (No x Nm) outputs.
 * Replaces the loop bruteforce.py:192-205 -> pdf.py:326-411.  data/err/mask host (No x Nf). */
int fzb_fit(fzb_handle h, const double* data, const double* data_err, const double* data_mask,
            int64_t No, const FzbConfig* cfg, const FzbFitOut* out);

/* Fused fit + predict without materialising the (No x Nm) matrix
 * (bruteforce.py:602-631 with save_fits=False).  pdfs: host (No x Ngrid); lmap/levid host [No];
 * best_* (nullable) host [No]: index / chi2 / scale of the max-lnprob model. */
int fzb_fit_predict(fzb_handle h, const double* data, const double* data_err, const double* data_mask,
                    int64_t No, const FzbConfig* cfg, double* pdfs, double* lmap, double* levid,
                    int64_t* best_idx, double* best_chi2, double* best_scale);

/* fzb_fit_predict with pdf.pdfs_summarize (pdf.py:899-1074; demos/3 cell 21, the immediate consumer of the PDFs) applied
 * to every PDF while it is still on the device: SURVEY.md section 8f rank 2.  pgrid [Ngrid], loss (Ngrid x Ngrid) =
 * 1 - kernel[truth, guess], urand [No] as for fzb_pdfs_summarize; the confidence width is the reference's default
 * wconf_func, wconf_frac * (1 + estimator) with wconf_frac = 0.03 (pdf.py:1040-1042).  Outputs (host): est / std / conf /
 * risk [4][No] for (mean, median, mode, best), quant [4][No] (2.5, 16, 84, 97.5 %), mc [No].  pdfs may be NULL: then only
 * ~170 + 40 bytes per object leave the device instead of Ngrid x 8. */
int fzb_fit_predict_summarize(fzb_handle h, const double* data, const double* data_err, const double* data_mask,
                              int64_t No, const FzbConfig* cfg, const double* pgrid, const double* loss,
                              const double* urand, int32_t renormalize, double wconf_frac, double* pdfs, double* lmap,
                              double* levid, int64_t* best_idx, double* best_chi2, double* best_scale, double* est,
                              double* std, double* conf, double* risk, double* quant, double* mc);

/* Same, all pointers on the device (inputs resident in HBM; used by bench.py "value"
 * and by the multi-GPU driver). */
int fzb_fit_predict_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask,
                        int64_t No, const FzbConfig* cfg, double* d_pdfs, double* d_lmap, double* d_levid,
                        int64_t* d_best_idx, double* d_best_chi2, double* d_best_scale);

/* PDFs from a caller-supplied (No x W) log-weight matrix (bruteforce.py:358-372).
 * neighbors == NULL: W == Nm and column j is model j.  Otherwise the kNN form
 * (knn.py:541-555): row i uses its first nneighbors[i] columns, column c is model
 * neighbors[i*W + c]. host pointers. */
int fzb_predict_logwt(fzb_handle h, const double* logwt, int64_t No, int64_t W,
                      const int64_t* neighbors, const int64_t* nneighbors, const FzbConfig* cfg,
                      double* pdfs, double* lmap, double* levid);

/* Model-sharded building blocks (one rank holds a slice of the models).  Device pointers.
 * pass 1: per-object partial (max lnprob, sum exp(lnprob - max), argmax) over the local models;
 * the caller merges partials across ranks (max / logsumexp all-reduce), then
 * pass 2 accumulates the un-normalised PDF of the local models for the GLOBAL lmap, levid;
 * the caller sums PDF partials across ranks and normalises. */
int fzb_shard_pass1_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask,
                        int64_t No, const FzbConfig* cfg, double* d_pmax, double* d_psum, int64_t* d_pbest);
int fzb_shard_pass2_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask,
                        int64_t No, const FzbConfig* cfg, const double* d_lmap, const double* d_levid,
                        double* d_pdf_partial);

/* Model-sharded mode as it runs over NCCL (frankenz_b200/distributed.py): objects go through in chunks and, per chunk,
 *   fzb_shard_pass1_packed_dev   d_packed[3][No] = (max, sum, bit pattern of the int64 GLOBAL arg-max = local + best_offset)
 *                                -> ONE ncclAllGather of 24 B / object / rank
 *   fzb_shard_merge_dev          d_gathered[world][3][No] -> global lmap / levid / best (numpy NaN rules, bruteforce.py:359)
 *   fzb_shard_pass2_f32_dev      un-normalised PDF partial of the local models in fp32, (No x Ngrid)
 *                                -> ONE ncclReduceScatter(sum): every rank receives the rows of the objects it owns
 *   fzb_shard_normalise_dev      float64 normalisation of the n reduced rows a rank owns (bruteforce.py:370)
 * merge / normalise only enqueue work on the handle's stream (no host synchronisation), so that the caller can order
 * them against its communication stream with events: fzb_get_stream returns the handle's cudaStream_t for that. */
int fzb_shard_pass1_packed_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask,
                               int64_t No, const FzbConfig* cfg, int64_t best_offset, double* d_packed);
int fzb_shard_merge_dev(fzb_handle h, const double* d_gathered, int32_t world, int64_t No, double* d_lmap,
                        double* d_levid, int64_t* d_best);
int fzb_shard_pass2_f32_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask,
                            int64_t No, const FzbConfig* cfg, const double* d_lmap, const double* d_levid,
                            float* d_pdf_partial);
int fzb_shard_normalise_dev(fzb_handle h, const float* d_rows, int64_t n, int32_t Ngrid, double* d_pdfs);
int fzb_get_stream(fzb_handle h, void** stream);

/* kNN: brute-force replacement of K cKDTrees (knn.py:186, :362-365).
 * feats: host float32 (K x Nm x Nf), the MC-realised, feature-mapped training sets
 * (knn.py:177-184; generated by the caller's RandomState so draws match the reference). */
int fzb_knn_build(fzb_handle h, const float* feats, int32_t K, int64_t Nm, int32_t Nf);
/* qfeats: host float64 (No x Nf).  idx: host int64 (No x K x k), tree-major, ascending
 * float64 Minkowski-p distance (p = 1, 2, or <=0 for infinity), exact (eps = 0).
 * dist (nullable): host float64 (No x K x k). */
int fzb_knn_query(fzb_handle h, const double* qfeats, int64_t No, int32_t k, double p,
                  int64_t* idx, double* dist);
/* Ordered union of the K*k neighbours (pandas.unique order, knn.py:368) and the likelihood
 * of each object against its union (knn.py:375-386), padded like knn.py:342-352.
 * neighbors: host int64 (No x K*k) filled with -99 beyond nneighbors[i]. out arrays are
 * (No x K*k). */
int fzb_knn_fit(fzb_handle h, const double* qfeats, const double* data, const double* data_err,
                const double* data_mask, int64_t No, int32_t k, double p, const FzbConfig* cfg,
                int64_t* neighbors, int64_t* nneighbors, const FzbFitOut* out);

/* Search + union + fits + KDE in one call, without the (No x K k) fit arrays: NearestNeighbors.fit_predict with
 * save_fits=False (knn.py:722-874).  pdfs host (No x Ngrid); lmap / levid / nneighbors nullable host [No]. */
int fzb_knn_fit_predict(fzb_handle h, const double* qfeats, const double* data, const double* data_err,
                        const double* data_mask, int64_t No, int32_t k, double p, const FzbConfig* cfg, double* pdfs,
                        double* lmap, double* levid, int64_t* nneighbors);

/* Likelihood of every object against its own list of models (the gather half of fzb_knn_fit, lists from the host):
 * SOM / GNG node-fit second stage, networks.py:918-923.  neighbors host int64 (No x W), row i uses its first nneighbors[i]
 * entries; out arrays (No x W), padded like fzb_knn_fit. */
int fzb_fit_gather(fzb_handle h, const double* data, const double* data_err, const double* data_mask, int64_t No,
                   int64_t W, const int64_t* neighbors, const int64_t* nneighbors, const FzbConfig* cfg,
                   const FzbFitOut* out);

/* PDF summary statistics: replaces pdf.pdfs_summarize (pdf.py:899-1074).
 * pdfs: host (No x Ng) float64, not modified; pgrid: host [Ng] (2 <= Ng <= 1024); loss: host (Ng x Ng) float64 =
 * 1 - kernel[truth, guess] as pdf.py:1003-1024 builds it; urand: host [No], the rstate.rand() of each object in order
 * (pdf.py:995).  renormalize != 0: every row is divided by its sum (numpy's pairwise order) first and rowsum (nullable,
 * host [No]) receives the sums, so that the caller can mirror the reference's in-place `pdfs /= sum` (pdf.py:980).
 * Outputs, host: est / std / risk [4][No] for (mean, median, mode, best); quant [4][No] = 2.5, 16, 84, 97.5 %
 * quantiles; mc [No] Monte-Carlo draws.  The CDFs stay on the device for fzb_pdfs_conf. */
int fzb_pdfs_summarize(fzb_handle h, const double* pdfs, const double* pgrid, const double* loss, const double* urand,
                       int64_t No, int32_t Ng, int32_t renormalize, double* rowsum, double* est, double* std,
                       double* risk, double* quant, double* mc);
/* Second stage (pdf.py:1038-1062): conf[e][i] = CDF_i(point + width) - CDF_i(point - width) for points / widths [4][No]
 * (the four estimators and the caller's wconf_func evaluated at them), on the PDFs of the last fzb_pdfs_summarize. */
int fzb_pdfs_conf(fzb_handle h, const double* points, const double* widths, int64_t No, double* conf);

/* Population likelihood of a redshift distribution given the PDFs (samplers.loglike_nz, samplers.py:24-76; SURVEY 8f
 * rank 4): overlap_i = sum_g pdfs[i,g] nz[g] (+ pair_step (pdfs[i,pair_i] - pdfs[i,pair_j]) when pair_i, pair_j >= 0),
 * lnlike = sum_i log(overlap_i); -inf (and zero overlaps) when nz has a negative or non-finite entry.  The PDFs stay
 * resident on the device between calls (the MCMC samplers call this thousands of times on the same PDFs):
 * fzb_nz_set_pdfs uploads a host array (No x Ng), fzb_nz_set_pdfs_dev adopts a device pointer (e.g. the output of
 * fzb_fit_predict_dev; it must stay valid).  overlap: nullable host [No]. */
int fzb_nz_set_pdfs(fzb_handle h, const double* pdfs, int64_t No, int32_t Ngrid);
int fzb_nz_set_pdfs_dev(fzb_handle h, const double* d_pdfs, int64_t No, int32_t Ngrid);
int fzb_nz_loglike(fzb_handle h, const double* nz, int32_t Ngrid, int32_t pair_i, int32_t pair_j, double pair_step,
                   double* lnlike, double* overlap);

/* Page-locked host memory for large outputs (the (Ndata x Ngrid) PDFs): when the `pdfs` argument of fzb_fit_predict
 * points into such a buffer, the device-to-host copies go straight into it, without the staging buffer and the host-side
 * copy a pageable destination needs. */
int fzb_alloc_pinned(size_t bytes, void** out);
int fzb_free_pinned(void* p);

/* Host helper: pdf.loglike's in-place cleaning (pdf.py:310-311) of n = Ndata x Nfilt float64 entries: where data or
 * err is non-finite or err <= 0, data = 0, err = 1, mask = 0.  Multi-threaded; no device involved. */
int fzb_clean_inplace_f64(double* data, double* err, double* mask, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* FRANKENZ_B200_H */
