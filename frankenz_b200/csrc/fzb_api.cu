// C-ABI entry points of libfzb200 (see include/frankenz_b200.h): handle management, host<->device
// staging, and dispatch between the fp32 register-tiled path (fzb_fast.cu) and the generic
// float64 path (fzb_generic.cu).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <mutex>
#include <thread>

#include "fzb_common.cuh"

static thread_local std::string g_err;

void fzb_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

namespace {

struct Timer {
    fzb_context* h;
    explicit Timer(fzb_context* hh) : h(hh) { cudaEventRecord(h->ev[0], h->stream); }
    int stop() {
        FZB_CUDA(cudaEventRecord(h->ev[1], h->stream));
        FZB_CUDA(cudaEventSynchronize(h->ev[1]));
        float ms = 0.f;
        FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        h->stats.ms_total = ms;
        return 0;
    }
};

int use_device(fzb_context* h) {
    FZB_CHECK(h != nullptr, "null handle");
    FZB_CUDA(cudaSetDevice(h->device));
    return 0;
}

template <class T>
int upload(fzb_context* h, DevBuf& b, const T* src, size_t n) {
    if (b.reserve(n * sizeof(T) + 16)) return 1;
    if (n) FZB_CUDA(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return 0;
}
template <class T>
int download(fzb_context* h, T* dst, const void* src, size_t n) {
    if (n && dst) FZB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

void reset_stats(fzb_context* h) { h->stats = FzbStats{}; }

int check_prior_bins(fzb_context* h, int64_t No) {
    if (h->prior_nbins > 0 && h->prior_bins_n > 0)
        FZB_CHECK(h->prior_bins_n == No, "prior bins were set for %lld objects, this call has %lld",
                  (long long)h->prior_bins_n, (long long)No);
    h->prior_o0 = 0;
    return 0;
}
void consume_prior_bins(fzb_context* h) {
    h->prior_bins_n = 0;
    h->prior_o0 = 0;
}

// a model whose label lies outside the PDF grid (or uses a malformed dictionary kernel) was selected by the weight
// threshold: the reference raises there (pdf.py:612-620)
int check_kde_error(fzb_context* h) {
    if (h->labels_bad <= 0 || h->kde_mode != FZB_KDE_DICT || !h->kde_err.p) return 0;
    long long flag[2] = {0, 0};
    FZB_CUDA(cudaMemcpyAsync(flag, h->kde_err.p, 16, cudaMemcpyDeviceToHost, h->stream));
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    if (flag[0] != 0) {
        FZB_CUDA(cudaMemsetAsync(h->kde_err.p, 0, 16, h->stream));
        fzb_set_error("model %lld passed the weight threshold but its label lies further outside the PDF grid than its "
                      "kernel width, or maps to a dictionary kernel wider than the grid (the reference raises here, "
                      "pdf.py:612-620)", flag[1]);
        return 2;
    }
    return 0;
}

int check_models(fzb_context* h) {
    FZB_CHECK(h->Nm > 0 && h->Nf > 0, "no models loaded: call fzb_set_models first");
    return 0;
}


// ---- staged device -> host download ------------------------------------------------------------------
// Large results (PDFs, full fit arrays) go device -> pinned staging (copy engine, second stream) -> caller's
// pageable array (a pool of host threads), in pieces, double buffered, so that PCIe and the host memcpy (incl. the
// first-touch page faults of a freshly allocated numpy array) overlap with each other and with the kernels.
void parallel_memcpy(char* dst, const char* src, size_t bytes, int nthreads) {
    if (bytes < ((size_t)4 << 20) || nthreads <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> pool;
    size_t per = (bytes / nthreads + 4095) & ~(size_t)4095;
    for (int t = 0; t < nthreads; ++t) {
        size_t lo = (size_t)t * per;
        if (lo >= bytes) break;
        size_t n = std::min(per, bytes - lo);
        pool.emplace_back([=] { memcpy(dst + lo, src + lo, n); });
    }
    for (auto& th : pool) th.join();
}

class StagedDownloader {
  public:
    explicit StagedDownloader(fzb_context* h) : h_(h) {}
    ~StagedDownloader() { finish(); }
    // direct: every destination of this downloader is page-locked host memory
    int init(size_t piece_bytes = (size_t)64 << 20, bool direct = false) {
        piece_ = piece_bytes;
        direct_ = direct;
        if (!h_->stream2) FZB_CUDA(cudaStreamCreateWithFlags(&h_->stream2, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            if (h_->pinned_cap[b] < piece_) {
                if (h_->pinned[b]) cudaFreeHost(h_->pinned[b]);
                h_->pinned[b] = nullptr;
                h_->pinned_cap[b] = 0;
                FZB_CUDA(cudaHostAlloc(&h_->pinned[b], piece_, cudaHostAllocDefault));
                h_->pinned_cap[b] = piece_;
            }
            if (!h_->ev_copied[b]) FZB_CUDA(cudaEventCreateWithFlags(&h_->ev_copied[b], cudaEventDisableTiming));
        }
        nthreads_ = (int)std::min<unsigned>(12, std::max(1u, std::thread::hardware_concurrency() / 2));
        started_ = true;
        worker_ = std::thread([this] { work_loop(); });
        return 0;
    }
    // Non-blocking: copy `bytes` from device `src` to host `dst` once everything enqueued so far on the compute
    // stream is done.  Returns the id of the request (for fence_compute) in *id.
    int push(void* dst, const void* src, size_t bytes, int64_t* id = nullptr) {
        if (bytes == 0) return 0;
        Request r;
        r.dst = static_cast<char*>(dst);
        r.src = static_cast<const char*>(src);
        r.bytes = bytes;
        r.direct = direct_;
        FZB_CUDA(cudaEventCreateWithFlags(&r.ready, cudaEventDisableTiming));
        FZB_CUDA(cudaEventCreateWithFlags(&r.left_device, cudaEventDisableTiming));
        FZB_CUDA(cudaEventRecord(r.ready, h_->stream));
        {
            std::lock_guard<std::mutex> lk(mu_);
            reqs_.push_back(r);
            if (id) *id = (int64_t)reqs_.size() - 1;
        }
        cv_.notify_all();
        return 0;
    }
    // Later work on the compute stream may overwrite the device buffers of requests 0..id.
    int fence_compute(int64_t id) {
        if (id < 0) return 0;
        cudaEvent_t ev;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return issued_ > id; });     // its last device->pinned copy has been enqueued
            ev = reqs_[(size_t)id].left_device;
        }
        FZB_CUDA(cudaStreamWaitEvent(h_->stream, ev, 0));
        return 0;
    }
    int finish() {
        if (!started_) return err_;
        {
            std::lock_guard<std::mutex> lk(mu_);
            closing_ = true;
        }
        cv_.notify_all();
        if (worker_.joinable()) worker_.join();
        started_ = false;
        for (auto& r : reqs_) {
            cudaEventDestroy(r.ready);
            cudaEventDestroy(r.left_device);
        }
        reqs_.clear();
        return err_;
    }

  private:
    struct Request {
        char* dst;
        const char* src;
        size_t bytes;
        bool direct;
        cudaEvent_t ready, left_device;
    };
    // one worker: enqueue the device->pinned copy of piece p, then spread piece p-1 into the caller's array while
    // the copy engine moves piece p
    void work_loop() {
        cudaSetDevice(h_->device);
        int64_t piece = 0;
        char* prev_dst = nullptr;
        size_t prev_bytes = 0;
        auto drain_prev = [&] {
            if (!prev_dst) return;
            if (cudaEventSynchronize(h_->ev_copied[(piece - 1) & 1]) != cudaSuccess) err_ = 1;
            parallel_memcpy(prev_dst, static_cast<const char*>(h_->pinned[(piece - 1) & 1]), prev_bytes, nthreads_);
            prev_dst = nullptr;
        };
        for (int64_t q = 0;; ++q) {
            Request r;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return (int64_t)reqs_.size() > q || closing_; });
                if ((int64_t)reqs_.size() <= q) break;
                r = reqs_[(size_t)q];
            }
            if (cudaStreamWaitEvent(h_->stream2, r.ready, 0) != cudaSuccess) err_ = 1;
            if (r.direct) {
                // page-locked destination (fzb_alloc_pinned): the copy engine writes the caller's array itself
                if (cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyDeviceToHost, h_->stream2) != cudaSuccess) err_ = 1;
                if (cudaEventRecord(r.left_device, h_->stream2) != cudaSuccess) err_ = 1;
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    issued_ = q + 1;
                }
                cv_.notify_all();
                continue;
            }
            for (size_t off = 0; off < r.bytes; off += piece_) {
                size_t n = std::min(piece_, r.bytes - off);
                int b = (int)(piece & 1);
                // staging buffer b was drained two pieces ago (this thread did it)
                if (cudaMemcpyAsync(h_->pinned[b], r.src + off, n, cudaMemcpyDeviceToHost, h_->stream2) != cudaSuccess)
                    err_ = 1;
                if (cudaEventRecord(h_->ev_copied[b], h_->stream2) != cudaSuccess) err_ = 1;
                ++piece;
                // while that copy runs, deliver the previous piece
                if (prev_dst) {
                    if (cudaEventSynchronize(h_->ev_copied[(piece - 2) & 1]) != cudaSuccess) err_ = 1;
                    parallel_memcpy(prev_dst, static_cast<const char*>(h_->pinned[(piece - 2) & 1]), prev_bytes, nthreads_);
                }
                prev_dst = r.dst + off;
                prev_bytes = n;
            }
            if (cudaEventRecord(r.left_device, h_->stream2) != cudaSuccess) err_ = 1;
            {
                std::lock_guard<std::mutex> lk(mu_);
                issued_ = q + 1;
            }
            cv_.notify_all();
        }
        drain_prev();
        if (direct_ && cudaStreamSynchronize(h_->stream2) != cudaSuccess) err_ = 1;
    }
    fzb_context* h_;
    bool direct_ = false;
    size_t piece_ = 0;
    int nthreads_ = 1;
    bool started_ = false, closing_ = false;
    int err_ = 0;
    int64_t issued_ = 0;
    std::vector<Request> reqs_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread worker_;
};

// dependency-free FFMA / MUFU loops: 8 independent chains per thread, 2048 resident threads per SM
__global__ void __launch_bounds__(256) k_peak_ffma(float* out, int iters) {
    float a[8];
    float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-9f * blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.1f * i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __fmaf_rn(a[i], x, y);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_peak_mufu(float* out, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (i + 1) - 1e-6f * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;
}

}  // namespace

extern "C" {

int fzb_measure_peaks(fzb_handle h, int reps, double* fp32_tflops, double* mufu_gops) {
    if (use_device(h)) return 2;
    FZB_CHECK(fp32_tflops && mufu_gops, "null pointer");
    if (h->misc[7].reserve(256)) return 1;
    float* out = h->misc[7].as<float>();
    const int blocks = h->sm_count * 8, iters = 4096;
    double best_f = 0.0, best_m = 0.0;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps + 1; ++r) {
        float ms = 0.f;
        FZB_CUDA(cudaEventRecord(h->ev[0], h->stream));
        k_peak_ffma<<<blocks, 256, 0, h->stream>>>(out, iters);
        FZB_CUDA(cudaEventRecord(h->ev[1], h->stream));
        FZB_CUDA(cudaEventSynchronize(h->ev[1]));
        FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        double fl = 2.0 * 64.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (r > 0 && fl > best_f) best_f = fl;
        FZB_CUDA(cudaEventRecord(h->ev[0], h->stream));
        k_peak_mufu<<<blocks, 256, 0, h->stream>>>(out, iters / 4);
        FZB_CUDA(cudaEventRecord(h->ev[1], h->stream));
        FZB_CUDA(cudaEventSynchronize(h->ev[1]));
        FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        double mo = 64.0 * (iters / 4) * 256.0 * blocks / (ms * 1e-3) / 1e9;
        if (r > 0 && mo > best_m) best_m = mo;
    }
    *fp32_tflops = best_f;
    *mufu_gops = best_m;
    return 0;
}

const char* fzb_last_error(void) { return g_err.c_str(); }
int fzb_version(void) { return 100; }

int fzb_device_count(int* count) {
    FZB_CHECK(count != nullptr, "null pointer");
    *count = 0;
    FZB_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int fzb_create(int device, fzb_handle* out) {
    FZB_CHECK(out != nullptr, "null output pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fzb_set_error("no usable CUDA device (%s); frankenz_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 3;
    }
    FZB_CHECK(device >= 0 && device < n, "device %d out of range (found %d)", device, n);
    FZB_CUDA(cudaSetDevice(device));
    fzb_context* h = new fzb_context();
    h->device = device;
    cudaDeviceProp prop;
    FZB_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    FZB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& ev : h->ev) FZB_CUDA(cudaEventCreate(&ev));
    *out = h;
    return 0;
}

int fzb_destroy(fzb_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf* bufs[] = {&h->models, &h->models_err, &h->models_mask, &h->lnprior, &h->prior_table, &h->prior_bins, &h->widths, &h->koff, &h->kernels,
                      &h->kcdf, &h->yidx, &h->ysidx, &h->grid, &h->y, &h->ystd, &h->lowers, &h->uppers, &h->rows,
                      &h->knn_feats, &h->knn_scan, &h->knn_buf, &h->knn_centre, &h->knn_cand, &h->knn_redo, &h->knn_tiles, &h->knn_aux, &h->kde_err, &h->summ[0], &h->summ[1], &h->summ[2], &h->summ[3], &h->summ[4], &h->summ[5], &h->fast.recs, &h->fast.recs_coarse, &h->fast.tiles_tc, &h->fast.tiles_tc_coarse, &h->fast.tiles_tc_f32, &h->fast.tiles_tc_f32_coarse, &h->fast.fuse, &h->fast.recs64, &h->fast.aux64, &h->fast.perm, &h->fast.bins, &h->fast.invnorm,
                      &h->fast.d_slot_sidx, &h->fast.live, &h->fast.live64, &h->fast.sortbuf, &h->fast.cutlist, &h->nz_pdfs, &h->nz_buf};
    for (auto* b : bufs) b->release();
    for (auto& b : h->obj_in) b.release();
    for (auto& b : h->out_f64) b.release();
    for (auto& b : h->out_i64) b.release();
    for (auto& b : h->misc) b.release();
    for (auto& ev : h->ev) cudaEventDestroy(ev);
    for (int b = 0; b < 2; ++b) {
        h->pdf_dev[b].release();
        if (h->pinned[b]) cudaFreeHost(h->pinned[b]);
        if (h->ev_done[b]) cudaEventDestroy(h->ev_done[b]);
        if (h->ev_copied[b]) cudaEventDestroy(h->ev_copied[b]);
        if (h->ev_chunk[b]) cudaEventDestroy(h->ev_chunk[b]);
    }
    if (h->stream2) cudaStreamDestroy(h->stream2);
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int fzb_synchronize(fzb_handle h) {
    if (use_device(h)) return 2;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

int fzb_get_stats(fzb_handle h, FzbStats* out) {
    FZB_CHECK(h && out, "null pointer");
    *out = h->stats;
    return 0;
}

int fzb_set_models(fzb_handle h, const double* models, const double* models_err, const double* models_mask,
                   int64_t Nm, int32_t Nf) {
    if (use_device(h)) return 2;
    FZB_CHECK(models && models_err && models_mask, "null model array");
    FZB_CHECK(Nm > 0 && Nf > 0, "empty model set (Nm=%lld, Nf=%d)", (long long)Nm, Nf);
    FZB_CHECK(Nf <= FZB_MAXF, "Nf=%d exceeds the supported maximum %d", Nf, FZB_MAXF);
    size_t n = (size_t)Nm * Nf;
    bool all_one = true, all_zero = true, finite = true, f32_exact = true, binary = true;
    for (size_t i = 0; i < n; ++i) {
        all_one &= (models_mask[i] == 1.0);
        binary &= (models_mask[i] == 1.0 || models_mask[i] == 0.0);
        all_zero &= (models_err[i] == 0.0);
        finite &= std::isfinite(models[i]) && std::isfinite(models_err[i]);
        f32_exact &= ((double)(float)models[i] == models[i]);
    }
    h->mask_all_one = all_one;
    h->mask_binary = binary;
    h->err_all_zero = all_zero;
    h->models_finite = finite;
    h->models_f32_exact = f32_exact;
    if (upload(h, h->models, models, n) || upload(h, h->models_err, models_err, n) ||
        upload(h, h->models_mask, models_mask, n))
        return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->Nm = Nm;
    h->Nf = Nf;
    h->has_lnprior = false;
    h->labels_dict_set = h->labels_grid_set = false;
    h->fast_dirty = true;
    return 0;
}

int fzb_set_lnprior(fzb_handle h, const double* lnprior, int64_t Nm) {
    if (use_device(h) || check_models(h)) return 2;
    if (lnprior == nullptr) {
        if (h->has_lnprior) h->fast_dirty = true;      // (every fit call passes through here: only a change costs a rebuild)
        h->has_lnprior = false;
        h->h_lnprior.clear();
        return 0;
    }
    FZB_CHECK(Nm == h->Nm, "lnprior has %lld entries, model set has %lld", (long long)Nm, (long long)h->Nm);
    if (h->has_lnprior && h->h_lnprior.size() == (size_t)Nm && memcmp(h->h_lnprior.data(), lnprior, (size_t)Nm * sizeof(double)) == 0)
        return 0;
    h->h_lnprior.assign(lnprior, lnprior + Nm);
    if (upload(h, h->lnprior, lnprior, (size_t)Nm)) return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->has_lnprior = true;
    h->fast_dirty = true;
    return 0;
}

int fzb_set_lnprior_table(fzb_handle h, const double* table, int32_t nbins, int64_t Nm) {
    if (use_device(h) || check_models(h)) return 2;
    if (table == nullptr) {
        h->prior_nbins = 0;
        h->prior_bins_n = 0;
        return 0;
    }
    FZB_CHECK(nbins > 0, "nbins must be positive");
    FZB_CHECK(Nm == h->Nm, "prior table has %lld columns, model set has %lld", (long long)Nm, (long long)h->Nm);
    if (upload(h, h->prior_table, table, (size_t)nbins * Nm)) return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->prior_nbins = nbins;
    h->prior_bins_n = 0;
    return 0;
}

int fzb_set_object_prior_bins(fzb_handle h, const int32_t* bins, int64_t No) {
    if (use_device(h)) return 2;
    if (bins == nullptr) {
        h->prior_bins_n = 0;
        return 0;
    }
    FZB_CHECK(h->prior_nbins > 0, "call fzb_set_lnprior_table first");
    for (int64_t i = 0; i < No; ++i)
        FZB_CHECK(bins[i] >= 0 && bins[i] < h->prior_nbins, "object %lld: prior bin %d outside [0, %d)", (long long)i,
                  bins[i], h->prior_nbins);
    if (upload(h, h->prior_bins, bins, (size_t)No)) return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->prior_bins_n = No;
    return 0;
}

int fzb_set_kde_dict(fzb_handle h, int32_t Ngrid, int32_t Ndict, const int32_t* widths, const int64_t* koff,
                     const double* kernels, const double* kcdf) {
    if (use_device(h)) return 2;
    FZB_CHECK(Ngrid > 0 && Ndict > 0 && widths && koff && kernels && kcdf, "bad dictionary arguments");
    FZB_CHECK(Ngrid <= FZB_MAX_NGRID, "PDF grid of %d points exceeds the supported maximum %d", Ngrid, FZB_MAX_NGRID);
    for (int i = 0; i < Ndict; ++i)
        FZB_CHECK(koff[i + 1] - koff[i] == 2 * (int64_t)widths[i] + 1 || koff[i + 1] - koff[i] >= 0,
                  "malformed kernel table");
    size_t tot = (size_t)koff[Ndict];
    // the same dictionary as last time (fit_predict passes it with every call): keep the tables and whatever was derived
    // from them (sorted model records, tiles)
    if (h->kde_mode == FZB_KDE_DICT && h->Ng == Ngrid && h->Ndict == Ndict && h->h_kernels.size() == tot &&
        h->h_kcdf.size() == tot && h->h_widths.size() == (size_t)Ndict &&
        memcmp(h->h_widths.data(), widths, (size_t)Ndict * sizeof(int32_t)) == 0 &&
        memcmp(h->h_koff.data(), koff, ((size_t)Ndict + 1) * sizeof(int64_t)) == 0 &&
        memcmp(h->h_kernels.data(), kernels, tot * sizeof(double)) == 0 && memcmp(h->h_kcdf.data(), kcdf, tot * sizeof(double)) == 0)
        return 0;
    h->h_widths.assign(widths, widths + Ndict);
    h->h_koff.assign(koff, koff + Ndict + 1);
    h->h_kernels.assign(kernels, kernels + tot);
    h->h_kcdf.assign(kcdf, kcdf + tot);
    if (upload(h, h->widths, widths, (size_t)Ndict) || upload(h, h->koff, koff, (size_t)Ndict + 1) ||
        upload(h, h->kernels, kernels, tot) || upload(h, h->kcdf, kcdf, tot))
        return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->Ng = Ngrid;
    h->Ndict = Ndict;
    h->kde_mode = FZB_KDE_DICT;
    h->labels_dict_set = false;
    h->fast_dirty = true;
    return 0;
}

int fzb_set_labels_dict(fzb_handle h, const int64_t* y_idx, const int64_t* y_std_idx, int64_t Nm) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(h->kde_mode == FZB_KDE_DICT, "call fzb_set_kde_dict first");
    FZB_CHECK(Nm == h->Nm, "labels have %lld entries, model set has %lld", (long long)Nm, (long long)h->Nm);
    // The reference raises (shape mismatch / IndexError) only when a SELECTED model's kernel misses the grid or uses a
    // wrapped (malformed) dictionary entry (pdf.py:603-620, :814-818): such labels are accepted here, counted, and the
    // KDE kernels raise a flag when one of them passes the weight threshold (checked at the end of every call that
    // builds PDFs); the sweep kernels leave a model set with such labels to the float64 kernel.
    int64_t nbad = 0;
    for (int64_t j = 0; j < Nm; ++j) {
        int64_t si = y_std_idx[j];
        FZB_CHECK(si >= 0 && si < h->Ndict, "label %lld: dictionary index %lld out of range", (long long)j,
                  (long long)si);
        int64_t w = h->h_widths[si];
        int64_t pos = y_idx[j];
        if (h->h_koff[si + 1] - h->h_koff[si] != 2 * w + 1 || !(pos + w >= 0 && pos - w <= h->Ng - 1)) ++nbad;
    }
    h->labels_bad = nbad;
    if (nbad > 0) {
        if (h->kde_err.reserve(16)) return 1;
        FZB_CUDA(cudaMemsetAsync(h->kde_err.p, 0, 16, h->stream));
    }
    if (h->labels_dict_set && h->h_yidx.size() == (size_t)Nm && h->h_ysidx.size() == (size_t)Nm &&
        memcmp(h->h_yidx.data(), y_idx, (size_t)Nm * sizeof(int64_t)) == 0 &&
        memcmp(h->h_ysidx.data(), y_std_idx, (size_t)Nm * sizeof(int64_t)) == 0)
        return 0;          // unchanged labels: the sorted records stay valid
    h->h_yidx.assign(y_idx, y_idx + Nm);
    h->h_ysidx.assign(y_std_idx, y_std_idx + Nm);
    if (upload(h, h->yidx, y_idx, (size_t)Nm) || upload(h, h->ysidx, y_std_idx, (size_t)Nm)) return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->labels_dict_set = true;
    h->fast_dirty = true;
    return 0;
}

int fzb_set_kde_grid(fzb_handle h, const double* grid, int32_t Ngrid) {
    if (use_device(h)) return 2;
    FZB_CHECK(grid && Ngrid > 0, "bad grid arguments");
    FZB_CHECK(Ngrid <= FZB_MAX_NGRID, "PDF grid of %d points exceeds the supported maximum %d", Ngrid, FZB_MAX_NGRID);
    if (upload(h, h->grid, grid, (size_t)Ngrid)) return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->Ng = Ngrid;
    h->kde_mode = FZB_KDE_GRID;
    h->labels_grid_set = false;
    h->fast_dirty = true;
    return 0;
}

int fzb_set_labels_grid(fzb_handle h, const double* y, const double* y_std, const int64_t* lowers,
                        const int64_t* uppers, int64_t Nm) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(h->kde_mode == FZB_KDE_GRID, "call fzb_set_kde_grid first");
    FZB_CHECK(Nm == h->Nm, "labels have %lld entries, model set has %lld", (long long)Nm, (long long)h->Nm);
    for (int64_t j = 0; j < Nm; ++j)
        FZB_CHECK(lowers[j] >= 0 && uppers[j] <= h->Ng, "label %lld: window [%lld, %lld) outside the grid",
                  (long long)j, (long long)lowers[j], (long long)uppers[j]);
    if (upload(h, h->y, y, (size_t)Nm) || upload(h, h->ystd, y_std, (size_t)Nm) ||
        upload(h, h->lowers, lowers, (size_t)Nm) || upload(h, h->uppers, uppers, (size_t)Nm))
        return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->labels_grid_set = true;
    h->fast_dirty = true;
    return 0;
}

int fzb_fit(fzb_handle h, const double* data, const double* data_err, const double* data_mask, int64_t No,
            const FzbConfig* cfg, const FzbFitOut* out) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(cfg && out, "null config / output struct");
    FZB_CHECK(No >= 0, "negative object count");
    reset_stats(h);
    if (No == 0) return 0;
    if (check_prior_bins(h, No)) return 2;
    const int64_t Nm = h->Nm;
    const int Nf = h->Nf;
    // chunk the objects so that the staged (chunk x Nm) outputs stay within ~6 GB of HBM
    int nout = (out->lnprior != nullptr) + (out->lnlike != nullptr) + (out->lnprob != nullptr) +
               (out->Ndim != nullptr) + (out->chi2 != nullptr) + (out->scale != nullptr) + (out->scale_err != nullptr);
    size_t per_obj = (size_t)Nm * 8 * (size_t)(nout > 0 ? nout : 1);
    int64_t chunk = (int64_t)(((size_t)6 << 30) / per_obj);
    if (chunk < 1) chunk = 1;
    if (chunk > No) chunk = No;
    size_t cn = (size_t)chunk * Nm;
    double* d_o[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double* hostp[6] = {out->lnprior, out->lnlike, out->lnprob, out->chi2, out->scale, out->scale_err};
    for (int i = 0; i < 6; ++i)
        if (hostp[i]) {
            if (h->out_f64[i].reserve(cn * 8)) return 1;
            d_o[i] = h->out_f64[i].as<double>();
        }
    int64_t* d_nd = nullptr;
    if (out->Ndim) {
        if (h->out_i64[0].reserve(cn * 8)) return 1;
        d_nd = h->out_i64[0].as<int64_t>();
    }
    StagedDownloader dl(h);
    if (dl.init()) return 1;
    int64_t last_push = -1;
    Timer t(h);
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        int64_t nc = No - o0 < chunk ? No - o0 : chunk;
        size_t nin = (size_t)nc * Nf;
        if (upload(h, h->obj_in[0], data + o0 * Nf, nin) || upload(h, h->obj_in[1], data_err + o0 * Nf, nin) ||
            upload(h, h->obj_in[2], data_mask + o0 * Nf, nin))
            return 1;
        if (dl.fence_compute(last_push)) return 1;     // the previous chunk's outputs have left the device buffers
        h->prior_o0 = o0;
        if (fzb_generic_fit_dev(h, h->obj_in[0].as<double>(), h->obj_in[1].as<double>(), h->obj_in[2].as<double>(), nc,
                                *cfg, d_o[0], d_o[1], d_o[2], d_nd, d_o[3], d_o[4], d_o[5]))
            return 1;
        size_t no = (size_t)nc * Nm;
        for (int i = 0; i < 6; ++i)
            if (hostp[i] && dl.push(hostp[i] + (size_t)o0 * Nm, d_o[i], no * sizeof(double), &last_push)) return 1;
        if (out->Ndim && dl.push(out->Ndim + (size_t)o0 * Nm, d_nd, no * sizeof(int64_t), &last_push)) return 1;
    }
    int rc = t.stop();
    consume_prior_bins(h);
    FZB_CHECK(dl.finish() == 0, "device-to-host copy of the fit arrays failed");
    return rc;
}

static int fit_predict_dev_impl(fzb_context* h, const double* d_data, const double* d_err, const double* d_mask,
                                int64_t No, const FzbConfig* cfg, double* d_pdfs, double* d_lmap, double* d_levid,
                                int64_t* d_best_idx, double* d_best_chi2, double* d_best_scale) {
    if (cfg->precision != FZB_PREC_FP64 && fzb_fast_supported(h, *cfg))
        return fzb_fast_fit_predict_dev(h, d_data, d_err, d_mask, No, *cfg, d_pdfs, d_lmap, d_levid, d_best_idx,
                                        d_best_chi2, d_best_scale);
    FZB_CHECK(cfg->precision != FZB_PREC_FP32, "the fp32 path does not support this configuration");
    return fzb_generic_fit_predict_dev(h, d_data, d_err, d_mask, No, nullptr, 0, *cfg, d_pdfs, d_lmap, d_levid,
                                       d_best_idx, d_best_chi2, d_best_scale);
}

int fzb_fit_predict_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask, int64_t No,
                        const FzbConfig* cfg, double* d_pdfs, double* d_lmap, double* d_levid, int64_t* d_best_idx,
                        double* d_best_chi2, double* d_best_scale) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(cfg != nullptr, "null config");
    reset_stats(h);
    if (No == 0) return 0;
    if (check_prior_bins(h, No)) return 2;
    Timer t(h);
    int rc = fit_predict_dev_impl(h, d_data, d_err, d_mask, No, cfg, d_pdfs, d_lmap, d_levid, d_best_idx, d_best_chi2,
                                  d_best_scale);
    consume_prior_bins(h);
    if (rc) return rc;
    if (t.stop()) return 1;
    return check_kde_error(h);
}

// Host-pointer form.  The objects are processed in chunks; the PDFs of chunk c leave through the staged downloader
// while the kernels of chunk c+1 run, so PCIe and the host memcpy are hidden behind the compute.  With `summ` the PDF
// rows of every chunk are summarised where they lie (pdf.pdfs_summarize, pdf.py:899-1074) and `pdfs` may be NULL: then
// 21 doubles per object cross PCIe instead of Ngrid.
struct SummArgs {
    const double *pgrid, *loss, *urand;
    int renormalize;
    double wfac;
    double *est, *sd, *conf, *risk, *quant, *mc;
};

static int fit_predict_host_impl(fzb_context* h, const double* data, const double* data_err, const double* data_mask,
                                 int64_t No, const FzbConfig* cfg, double* pdfs, double* lmap, double* levid,
                                 int64_t* best_idx, double* best_chi2, double* best_scale, const SummArgs* summ) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(cfg != nullptr, "null config");
    FZB_CHECK(No >= 0, "negative object count");
    reset_stats(h);
    if (No == 0) return 0;
    if (check_prior_bins(h, No)) return 2;
    const int Nf = h->Nf;
    const int Ng = h->Ng;
    size_t nin = (size_t)No * Nf;
    if (upload(h, h->obj_in[0], data, nin) || upload(h, h->obj_in[1], data_err, nin) ||
        upload(h, h->obj_in[2], data_mask, nin))
        return 1;
    if (h->out_f64[1].reserve((size_t)No * 8) || h->out_f64[2].reserve((size_t)No * 8) ||
        h->out_f64[3].reserve((size_t)No * 8) || h->out_f64[4].reserve((size_t)No * 8) ||
        h->out_i64[0].reserve((size_t)No * 8))
        return 1;
    if (summ) {
        FZB_CHECK(h->kde_mode != FZB_KDE_NONE && Ng > 0, "summaries need a configured KDE");
        if (fzb_summarize_tables(h, summ->pgrid, summ->loss, summ->urand, No, Ng)) return 1;
    }
    const double* d_x = h->obj_in[0].as<double>();
    const double* d_xe = h->obj_in[1].as<double>();
    const double* d_xm = h->obj_in[2].as<double>();
    double* d_lmap = h->out_f64[1].as<double>();
    double* d_levid = h->out_f64[2].as<double>();
    double* d_bc = h->out_f64[3].as<double>();
    double* d_bs = h->out_f64[4].as<double>();
    int64_t* d_bi = h->out_i64[0].as<int64_t>();
    const bool want_rows = pdfs != nullptr || summ != nullptr;     // PDF rows are needed on the device

    // Chunks pipeline the device-to-host copy of the PDFs behind the next chunk's kernels.  Every chunk costs a few ms
    // (launch tails, host round trips, less efficient small sweeps) and only the download of the last one is exposed.
    // Pageable destination (staged copy, ~19 GB/s measured on a B200 host: 0.29 us / object against 0.41 us / object of
    // sweep): chunks of 262,144 objects halving towards the end, down to 32k.  Page-locked destination (direct DMA,
    // ~50 GB/s): a chunk is 60 % of what remains, i.e. 2.5 times the next one.  FZB_E2E_CHUNK caps the chunk size (floor
    // 8192, multiples of 4096).
    bool pinned_dst = false;
    if (pdfs) {
        cudaPointerAttributes pa = {};
        pinned_dst = cudaPointerGetAttributes(&pa, pdfs) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
    }
    const bool geom = pinned_dst ? getenv("FZB_E2E_HALVING") == nullptr : getenv("FZB_E2E_GEOM") != nullptr;
    int64_t chunk = geom ? 1048576 : 262144;
    if (const char* e = getenv("FZB_E2E_CHUNK")) chunk = std::max<int64_t>(8192, (atoll(e) + 4095) / 4096 * 4096);
    // nothing to overlap when the PDFs stay on the device (summaries only): one chunk, bounded by the row buffer (12 GB)
    if (!pdfs) chunk = std::max<int64_t>(chunk, ((int64_t)12 << 30) / std::max<int64_t>(1, (int64_t)Ng * 8));
    if (!want_rows || (!pdfs && No <= chunk + chunk / 2)) chunk = No;
    auto next_chunk = [&](int64_t rem) -> int64_t {
        if (!pdfs) return std::min(rem, chunk);
        if (rem <= 49152) return rem;
        int64_t nc = chunk;
        if (geom) nc = std::max<int64_t>(32768, (rem * 6 / 10 + 4095) / 4096 * 4096);
        else if (rem < 2 * chunk) nc = std::max<int64_t>(32768, (rem / 2 + 4095) / 4096 * 4096);
        return std::min(std::min(nc, rem), chunk);
    };
    chunk = std::min(chunk, next_chunk(No));          // the largest chunk = the first one
    const size_t chunk_bytes = (size_t)chunk * Ng * sizeof(double);
    StagedDownloader dl(h);
    // a destination from fzb_alloc_pinned takes the DMA directly (no staging buffer, no host-side copy)
    if (pdfs && dl.init((size_t)64 << 20, pinned_dst)) return 1;
    if (want_rows)
        for (int b = 0; b < (pdfs ? 2 : 1); ++b)
            if (h->pdf_dev[b].reserve(chunk_bytes)) return 1;
    FZB_CUDA(cudaEventRecord(h->ev[0], h->stream));
    int64_t c = 0;
    int64_t push_id[2] = {-1, -1};
    int64_t nc = 0;
    for (int64_t o0 = 0; o0 < No; o0 += nc, ++c) {
        nc = next_chunk(No - o0);
        int b = pdfs ? (int)(c & 1) : 0;
        h->prior_o0 = o0;
        // device buffer b was last read by the download of chunk c-2 (issued before that of chunk c-1)
        if (pdfs && c >= 2 && dl.fence_compute(push_id[b])) return 1;
        int rc = fit_predict_dev_impl(h, d_x + o0 * Nf, d_xe + o0 * Nf, d_xm + o0 * Nf, nc, cfg,
                                      want_rows ? h->pdf_dev[b].as<double>() : nullptr, d_lmap + o0, d_levid + o0, d_bi + o0,
                                      d_bc + o0, d_bs + o0);
        if (rc) return rc;
        if (summ) {
            FZB_CUDA(cudaEventRecord(h->ev[6], h->stream));
            if (fzb_summarize_rows_dev(h, h->pdf_dev[b].as<double>(), nc, o0, No, Ng, summ->renormalize, summ->wfac)) return 1;
            FZB_CUDA(cudaEventRecord(h->ev[7], h->stream));
        }
        if (pdfs && dl.push(pdfs + (size_t)o0 * Ng, h->pdf_dev[b].p, (size_t)nc * Ng * sizeof(double), &push_id[b]))
            return 1;
        if (summ) {
            float ms = 0.f;
            FZB_CUDA(cudaEventSynchronize(h->ev[7]));
            FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]));
            h->stats.ms_summarize += ms;
        }
    }
    FZB_CUDA(cudaEventRecord(h->ev[1], h->stream));
    consume_prior_bins(h);
    if (download(h, lmap, d_lmap, (size_t)No) || download(h, levid, d_levid, (size_t)No) ||
        download(h, best_idx, d_bi, (size_t)No) || download(h, best_chi2, d_bc, (size_t)No) ||
        download(h, best_scale, d_bs, (size_t)No))
        return 1;
    if (summ && fzb_summarize_download(h, No, summ->est, summ->sd, summ->conf, summ->risk, summ->quant, summ->mc)) return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    FZB_CHECK(dl.finish() == 0, "device-to-host copy of the PDFs failed");
    float ms = 0.f;
    FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->stats.ms_total = ms;
    return check_kde_error(h);
}

int fzb_fit_predict(fzb_handle h, const double* data, const double* data_err, const double* data_mask, int64_t No,
                    const FzbConfig* cfg, double* pdfs, double* lmap, double* levid, int64_t* best_idx,
                    double* best_chi2, double* best_scale) {
    return fit_predict_host_impl(h, data, data_err, data_mask, No, cfg, pdfs, lmap, levid, best_idx, best_chi2, best_scale,
                                 nullptr);
}

int fzb_fit_predict_summarize(fzb_handle h, const double* data, const double* data_err, const double* data_mask,
                              int64_t No, const FzbConfig* cfg, const double* pgrid, const double* loss,
                              const double* urand, int32_t renormalize, double wconf_frac, double* pdfs, double* lmap,
                              double* levid, int64_t* best_idx, double* best_chi2, double* best_scale, double* est,
                              double* std, double* conf, double* risk, double* quant, double* mc) {
    FZB_CHECK(pgrid && loss && urand && est && std && conf && risk && quant && mc, "null summary argument");
    SummArgs sa = {pgrid, loss, urand, renormalize, wconf_frac, est, std, conf, risk, quant, mc};
    return fit_predict_host_impl(h, data, data_err, data_mask, No, cfg, pdfs, lmap, levid, best_idx, best_chi2, best_scale,
                                 &sa);
}

int fzb_predict_logwt(fzb_handle h, const double* logwt, int64_t No, int64_t W, const int64_t* neighbors,
                      const int64_t* nneighbors, const FzbConfig* cfg, double* pdfs, double* lmap, double* levid) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(cfg && logwt && pdfs, "null argument");
    FZB_CHECK((neighbors == nullptr) == (nneighbors == nullptr), "neighbors and nneighbors go together");
    if (!neighbors) FZB_CHECK(W == h->Nm, "logwt has %lld columns, model set has %lld", (long long)W, (long long)h->Nm);
    reset_stats(h);
    if (No == 0) return 0;
    const int Ng = h->Ng;
    // chunk so that the staged log-weights stay within ~8 GB
    int64_t chunk = (int64_t)(((size_t)8 << 30) / ((size_t)W * 8 * (neighbors ? 2 : 1)));
    if (chunk < 1) chunk = 1;
    if (chunk > No) chunk = No;
    if (h->out_f64[0].reserve((size_t)chunk * Ng * 8) || h->out_f64[1].reserve((size_t)chunk * 8) ||
        h->out_f64[2].reserve((size_t)chunk * 8))
        return 1;
    Timer t(h);
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        int64_t nc = No - o0 < chunk ? No - o0 : chunk;
        if (upload(h, h->misc[0], logwt + (size_t)o0 * W, (size_t)nc * W)) return 1;
        const int64_t *d_nb = nullptr, *d_nn = nullptr;
        if (neighbors) {
            if (upload(h, h->misc[1], neighbors + (size_t)o0 * W, (size_t)nc * W) ||
                upload(h, h->misc[2], nneighbors + o0, (size_t)nc))
                return 1;
            d_nb = h->misc[1].as<int64_t>();
            d_nn = h->misc[2].as<int64_t>();
        }
        if (fzb_generic_predict_logwt_dev(h, h->misc[0].as<double>(), nc, W, d_nb, d_nn, *cfg,
                                          h->out_f64[0].as<double>(), h->out_f64[1].as<double>(),
                                          h->out_f64[2].as<double>()))
            return 1;
        if (download(h, pdfs + (size_t)o0 * Ng, h->out_f64[0].p, (size_t)nc * Ng) ||
            download(h, lmap ? lmap + o0 : nullptr, h->out_f64[1].p, (size_t)nc) ||
            download(h, levid ? levid + o0 : nullptr, h->out_f64[2].p, (size_t)nc))
            return 1;
        FZB_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (t.stop()) return 1;
    return check_kde_error(h);
}

int fzb_shard_pass1_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask, int64_t No,
                        const FzbConfig* cfg, double* d_pmax, double* d_psum, int64_t* d_pbest) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(cfg != nullptr, "null config");
    reset_stats(h);
    if (No == 0) return 0;
    Timer t(h);
    if (cfg->precision != FZB_PREC_FP64 && fzb_fast_supported(h, *cfg)) {
        if (fzb_fast_fit_predict_dev(h, d_data, d_err, d_mask, No, *cfg, nullptr, d_pmax, nullptr, d_pbest, nullptr,
                                     nullptr, 1, d_psum, nullptr))
            return 1;
    } else {
        h->shard_valid = false;
        if (fzb_generic_shard_pass1_dev(h, d_data, d_err, d_mask, No, nullptr, 0, *cfg, d_pmax, d_psum, d_pbest))
            return 1;
    }
    return t.stop();
}

static int shard_pass2_impl(fzb_context* h, const double* d_data, const double* d_err, const double* d_mask, int64_t No,
                            const FzbConfig* cfg, const double* d_lmap, const double* d_levid, double* d_pdf_partial,
                            float* d_partial32) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(cfg != nullptr, "null config");
    reset_stats(h);
    if (No == 0) return 0;
    Timer t(h);
    h->shard_out32 = d_partial32;
    int rc = 0;
    if (cfg->precision != FZB_PREC_FP64 && fzb_fast_supported(h, *cfg) && h->shard_valid && h->shard_No == No) {
        rc = fzb_fast_fit_predict_dev(h, d_data, d_err, d_mask, No, *cfg, d_pdf_partial, nullptr, nullptr, nullptr,
                                      nullptr, nullptr, 2, nullptr, d_lmap);
    } else {
        rc = fzb_generic_shard_pass2_dev(h, d_data, d_err, d_mask, No, nullptr, 0, *cfg, d_lmap, d_levid, d_pdf_partial);
    }
    h->shard_out32 = nullptr;
    if (rc) return 1;
    if (t.stop()) return 1;
    return check_kde_error(h);
}

int fzb_shard_pass2_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask, int64_t No,
                        const FzbConfig* cfg, const double* d_lmap, const double* d_levid, double* d_pdf_partial) {
    FZB_CHECK(d_pdf_partial != nullptr, "null output");
    return shard_pass2_impl(h, d_data, d_err, d_mask, No, cfg, d_lmap, d_levid, d_pdf_partial, nullptr);
}

int fzb_shard_pass2_f32_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask, int64_t No,
                            const FzbConfig* cfg, const double* d_lmap, const double* d_levid, float* d_pdf_partial) {
    FZB_CHECK(d_pdf_partial != nullptr, "null output");
    return shard_pass2_impl(h, d_data, d_err, d_mask, No, cfg, d_lmap, d_levid, nullptr, d_pdf_partial);
}

int fzb_shard_pass1_packed_dev(fzb_handle h, const double* d_data, const double* d_err, const double* d_mask, int64_t No,
                               const FzbConfig* cfg, int64_t best_offset, double* d_packed) {
    FZB_CHECK(d_packed != nullptr, "null output");
    if (No == 0) return 0;
    int64_t* d_best = reinterpret_cast<int64_t*>(d_packed + 2 * No);
    int rc = fzb_shard_pass1_dev(h, d_data, d_err, d_mask, No, cfg, d_packed, d_packed + No, d_best);
    if (rc) return rc;
    return fzb_shard_add_offset_launch(h, d_best, No, best_offset);     // stream-ordered; no host synchronisation
}

int fzb_shard_merge_dev(fzb_handle h, const double* d_gathered, int32_t world, int64_t No, double* d_lmap,
                        double* d_levid, int64_t* d_best) {
    if (use_device(h)) return 2;
    FZB_CHECK(d_gathered && d_lmap && d_levid && world > 0 && No >= 0, "bad arguments");
    return fzb_shard_merge_launch(h, d_gathered, world, No, d_lmap, d_levid, d_best);
}

int fzb_shard_normalise_dev(fzb_handle h, const float* d_rows, int64_t n, int32_t Ng, double* d_pdfs) {
    if (use_device(h)) return 2;
    FZB_CHECK(d_rows && d_pdfs && n >= 0 && Ng > 0, "bad arguments");
    return fzb_shard_normalise_launch(h, d_rows, n, Ng, d_pdfs);
}

int fzb_get_stream(fzb_handle h, void** stream) {
    FZB_CHECK(h && stream, "null pointer");
    *stream = reinterpret_cast<void*>(h->stream);
    return 0;
}

int fzb_knn_build(fzb_handle h, const float* feats, int32_t K, int64_t Nm, int32_t Nf) {
    if (use_device(h)) return 2;
    FZB_CHECK(feats && K > 0 && Nm > 0 && Nf > 0, "bad kNN build arguments");
    FZB_CHECK(Nf <= FZB_MAXF, "Nf=%d exceeds the supported maximum %d", Nf, FZB_MAXF);
    // per-tree stride padded to 16 bytes so that every tile of every tree is a legal TMA bulk-copy source
    const int64_t stride = ((int64_t)Nm * Nf + 3 + 3) / 4 * 4;
    if (h->knn_feats.reserve((size_t)K * stride * sizeof(float) + 64)) return 1;
    FZB_CUDA(cudaMemsetAsync(h->knn_feats.p, 0, (size_t)K * stride * sizeof(float) + 64, h->stream));
    for (int t = 0; t < K; ++t)
        FZB_CUDA(cudaMemcpyAsync(h->knn_feats.as<float>() + (size_t)t * stride, feats + (size_t)t * Nm * Nf,
                                 (size_t)Nm * Nf * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->knn_stride = stride;
    h->knn_K = K;
    h->knn_Nm = Nm;
    h->knn_Nf = Nf;
    h->knn_tc_valid = false;
    if (fzb_knn_scan_build(h)) return 1;
    if (getenv("FZB_KNN_TC") != nullptr && atoi(getenv("FZB_KNN_TC")) != 0 && fzb_knn_tc_build(h)) return 1;
    return 0;
}

int fzb_knn_query(fzb_handle h, const double* qfeats, int64_t No, int32_t k, double p, int64_t* idx, double* dist) {
    if (use_device(h)) return 2;
    FZB_CHECK(h->knn_K > 0, "call fzb_knn_build first");
    FZB_CHECK(qfeats && idx, "null argument");
    FZB_CHECK(k > 0 && k <= h->knn_Nm, "k=%d must be in [1, Nmodel=%lld]", k, (long long)h->knn_Nm);
    reset_stats(h);
    if (No == 0) return 0;
    size_t nout = (size_t)No * h->knn_K * k;
    if (upload(h, h->obj_in[0], qfeats, (size_t)No * h->knn_Nf) || h->out_i64[0].reserve(nout * 8) ||
        h->out_f64[0].reserve(nout * 8))
        return 1;
    Timer t(h);
    if (fzb_knn_query_dev(h, h->obj_in[0].as<double>(), No, k, p, h->out_i64[0].as<int64_t>(),
                          h->out_f64[0].as<double>()))
        return 1;
    if (download(h, idx, h->out_i64[0].p, nout) || download(h, dist, h->out_f64[0].p, nout)) return 1;
    return t.stop();
}

int fzb_knn_fit(fzb_handle h, const double* qfeats, const double* data, const double* data_err,
                const double* data_mask, int64_t No, int32_t k, double p, const FzbConfig* cfg, int64_t* neighbors,
                int64_t* nneighbors, const FzbFitOut* out) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(h->knn_K > 0, "call fzb_knn_build first");
    FZB_CHECK(h->knn_Nm == h->Nm && h->knn_Nf == h->Nf, "kNN features (%lld x %d) do not match the model set (%lld x %d)",
              (long long)h->knn_Nm, h->knn_Nf, (long long)h->Nm, h->Nf);
    FZB_CHECK(qfeats && data && data_err && data_mask && cfg && neighbors && nneighbors && out, "null argument");
    FZB_CHECK(k > 0 && k <= h->knn_Nm, "k=%d must be in [1, Nmodel=%lld]", k, (long long)h->knn_Nm);
    reset_stats(h);
    if (No == 0) return 0;
    if (check_prior_bins(h, No)) return 2;
    const int Nf = h->Nf;
    const int64_t W = (int64_t)h->knn_K * k;
    // chunk the objects: 2 int64 + 7 output arrays of width W
    int64_t chunk = (int64_t)(((size_t)8 << 30) / ((size_t)W * 8 * 10));
    if (chunk < 1) chunk = 1;
    if (chunk > No) chunk = No;
    size_t cn = (size_t)chunk * W;
    double* hostp[6] = {out->lnprior, out->lnlike, out->lnprob, out->chi2, out->scale, out->scale_err};
    double* d_o[6] = {};
    for (int i = 0; i < 6; ++i)
        if (hostp[i]) {
            if (h->out_f64[i].reserve(cn * 8)) return 1;
            d_o[i] = h->out_f64[i].as<double>();
        }
    int64_t* d_nd = nullptr;
    if (out->Ndim) {
        if (h->out_i64[1].reserve(cn * 8)) return 1;
        d_nd = h->out_i64[1].as<int64_t>();
    }
    if (h->out_i64[0].reserve(cn * 8) || h->misc[3].reserve(cn * 8) || h->misc[4].reserve((size_t)chunk * 8)) return 1;
    Timer t(h);
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        int64_t nc = No - o0 < chunk ? No - o0 : chunk;
        size_t nin = (size_t)nc * Nf;
        if (upload(h, h->obj_in[0], data + o0 * Nf, nin) || upload(h, h->obj_in[1], data_err + o0 * Nf, nin) ||
            upload(h, h->obj_in[2], data_mask + o0 * Nf, nin) || upload(h, h->misc[5], qfeats + o0 * Nf, nin))
            return 1;
        h->prior_o0 = o0;
        int64_t* d_idx = h->out_i64[0].as<int64_t>();
        int64_t* d_nb = h->misc[3].as<int64_t>();
        int64_t* d_nn = h->misc[4].as<int64_t>();
        if (fzb_knn_query_dev(h, h->misc[5].as<double>(), nc, k, p, d_idx, nullptr)) return 1;
        if (fzb_knn_union_dev(h, d_idx, nc, (int)W, d_nb, d_nn)) return 1;
        if (fzb_generic_gather_fit_dev(h, h->obj_in[0].as<double>(), h->obj_in[1].as<double>(),
                                       h->obj_in[2].as<double>(), nc, W, d_nb, d_nn, *cfg, d_o[0], d_o[1], d_o[2],
                                       d_nd, d_o[3], d_o[4], d_o[5]))
            return 1;
        size_t no = (size_t)nc * W;
        for (int i = 0; i < 6; ++i)
            if (download(h, hostp[i] ? hostp[i] + (size_t)o0 * W : nullptr, d_o[i], no)) return 1;
        if (download(h, out->Ndim ? out->Ndim + (size_t)o0 * W : nullptr, d_nd, no) ||
            download(h, neighbors + (size_t)o0 * W, d_nb, no) || download(h, nneighbors + o0, d_nn, (size_t)nc))
            return 1;
        FZB_CUDA(cudaStreamSynchronize(h->stream));
    }
    consume_prior_bins(h);
    return t.stop();
}

// kNN search + union + fits + KDE without the (No x K k) fit arrays (knn.py:722-874 with save_fits=False): everything
// stays on the device, the PDFs leave through the staged downloader while the next chunk is searched.
int fzb_knn_fit_predict(fzb_handle h, const double* qfeats, const double* data, const double* data_err,
                        const double* data_mask, int64_t No, int32_t k, double p, const FzbConfig* cfg, double* pdfs,
                        double* lmap, double* levid, int64_t* nneighbors) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(h->knn_K > 0, "call fzb_knn_build first");
    FZB_CHECK(h->knn_Nm == h->Nm && h->knn_Nf == h->Nf, "kNN features (%lld x %d) do not match the model set (%lld x %d)",
              (long long)h->knn_Nm, h->knn_Nf, (long long)h->Nm, h->Nf);
    FZB_CHECK(qfeats && data && data_err && data_mask && cfg && pdfs, "null argument");
    FZB_CHECK(k > 0 && k <= h->knn_Nm, "k=%d must be in [1, Nmodel=%lld]", k, (long long)h->knn_Nm);
    FZB_CHECK(h->kde_mode != FZB_KDE_NONE, "no KDE configured");
    reset_stats(h);
    if (No == 0) return 0;
    if (check_prior_bins(h, No)) return 2;
    const int Nf = h->Nf, Ng = h->Ng;
    const int64_t W = (int64_t)h->knn_K * k;
    int64_t chunk = std::min<int64_t>(No, 65536);
    const size_t cn = (size_t)chunk * W;
    if (h->out_i64[0].reserve(cn * 8) || h->misc[3].reserve(cn * 8) || h->misc[4].reserve((size_t)chunk * 8) ||
        h->out_f64[1].reserve((size_t)No * 8) || h->out_f64[2].reserve((size_t)No * 8) || h->out_i64[1].reserve((size_t)No * 8))
        return 1;
    for (int b = 0; b < 2; ++b)
        if (h->pdf_dev[b].reserve((size_t)chunk * Ng * 8)) return 1;
    StagedDownloader dl(h);
    cudaPointerAttributes pa = {};
    const bool pinned_dst = cudaPointerGetAttributes(&pa, pdfs) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (dl.init((size_t)64 << 20, pinned_dst)) return 1;
    double* d_lmap = h->out_f64[1].as<double>();
    double* d_levid = h->out_f64[2].as<double>();
    int64_t* d_nn_all = h->out_i64[1].as<int64_t>();
    FZB_CUDA(cudaEventRecord(h->ev[0], h->stream));
    int64_t push_id[2] = {-1, -1};
    int64_t c = 0;
    for (int64_t o0 = 0; o0 < No; o0 += chunk, ++c) {
        const int64_t nc = std::min(chunk, No - o0);
        const int b = (int)(c & 1);
        const size_t nin = (size_t)nc * Nf;
        if (upload(h, h->obj_in[0], data + o0 * Nf, nin) || upload(h, h->obj_in[1], data_err + o0 * Nf, nin) ||
            upload(h, h->obj_in[2], data_mask + o0 * Nf, nin) || upload(h, h->misc[5], qfeats + o0 * Nf, nin))
            return 1;
        h->prior_o0 = o0;
        int64_t* d_idx = h->out_i64[0].as<int64_t>();
        int64_t* d_nb = h->misc[3].as<int64_t>();
        int64_t* d_nn = h->misc[4].as<int64_t>();
        if (fzb_knn_query_dev(h, h->misc[5].as<double>(), nc, k, p, d_idx, nullptr)) return 1;
        if (fzb_knn_union_dev(h, d_idx, nc, (int)W, d_nb, d_nn)) return 1;
        if (c >= 2 && dl.fence_compute(push_id[b])) return 1;
        if (fzb_generic_gather_fit_predict_dev(h, h->obj_in[0].as<double>(), h->obj_in[1].as<double>(),
                                               h->obj_in[2].as<double>(), nc, W, d_nb, d_nn, *cfg,
                                               h->pdf_dev[b].as<double>(), d_lmap + o0, d_levid + o0))
            return 1;
        FZB_CUDA(cudaMemcpyAsync(d_nn_all + o0, d_nn, (size_t)nc * 8, cudaMemcpyDeviceToDevice, h->stream));
        if (dl.push(pdfs + (size_t)o0 * Ng, h->pdf_dev[b].p, (size_t)nc * Ng * sizeof(double), &push_id[b])) return 1;
        // the next chunk's uploads reuse obj_in / misc: everything of this chunk must have been consumed
        FZB_CUDA(cudaStreamSynchronize(h->stream));
    }
    FZB_CUDA(cudaEventRecord(h->ev[1], h->stream));
    consume_prior_bins(h);
    if (download(h, lmap, d_lmap, (size_t)No) || download(h, levid, d_levid, (size_t)No) ||
        download(h, nneighbors, d_nn_all, (size_t)No))
        return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    FZB_CHECK(dl.finish() == 0, "device-to-host copy of the PDFs failed");
    float ms = 0.f;
    FZB_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    h->stats.ms_total = ms;
    return check_kde_error(h);
}

// Likelihood of every object against ITS OWN list of models (host lists): the gather half of fzb_knn_fit without the
// search.  Used by the SOM / GNG node-fit (networks.py:918-923: lprob_func(x, models[idxs], ...)).
int fzb_fit_gather(fzb_handle h, const double* data, const double* data_err, const double* data_mask, int64_t No,
                   int64_t W, const int64_t* neighbors, const int64_t* nneighbors, const FzbConfig* cfg,
                   const FzbFitOut* out) {
    if (use_device(h) || check_models(h)) return 2;
    FZB_CHECK(data && data_err && data_mask && cfg && neighbors && nneighbors && out && W > 0, "bad arguments");
    reset_stats(h);
    if (No == 0) return 0;
    if (check_prior_bins(h, No)) return 2;
    for (int64_t i = 0; i < No; ++i) {
        FZB_CHECK(nneighbors[i] >= 0 && nneighbors[i] <= W, "object %lld: %lld neighbours for a row width of %lld",
                  (long long)i, (long long)nneighbors[i], (long long)W);
        for (int64_t c = 0; c < nneighbors[i]; ++c)
            FZB_CHECK(neighbors[i * W + c] >= 0 && neighbors[i * W + c] < h->Nm, "object %lld: model index %lld out of range",
                      (long long)i, (long long)neighbors[i * W + c]);
    }
    const int Nf = h->Nf;
    int64_t chunk = (int64_t)(((size_t)6 << 30) / ((size_t)W * 8 * 9));
    chunk = std::max<int64_t>(1, std::min(chunk, No));
    const size_t cn = (size_t)chunk * W;
    double* hostp[6] = {out->lnprior, out->lnlike, out->lnprob, out->chi2, out->scale, out->scale_err};
    double* d_o[6] = {};
    for (int i = 0; i < 6; ++i)
        if (hostp[i]) {
            if (h->out_f64[i].reserve(cn * 8)) return 1;
            d_o[i] = h->out_f64[i].as<double>();
        }
    int64_t* d_nd = nullptr;
    if (out->Ndim) {
        if (h->out_i64[1].reserve(cn * 8)) return 1;
        d_nd = h->out_i64[1].as<int64_t>();
    }
    Timer t(h);
    for (int64_t o0 = 0; o0 < No; o0 += chunk) {
        const int64_t nc = std::min(chunk, No - o0);
        const size_t nin = (size_t)nc * Nf, no = (size_t)nc * W;
        if (upload(h, h->obj_in[0], data + o0 * Nf, nin) || upload(h, h->obj_in[1], data_err + o0 * Nf, nin) ||
            upload(h, h->obj_in[2], data_mask + o0 * Nf, nin) || upload(h, h->misc[3], neighbors + (size_t)o0 * W, no) ||
            upload(h, h->misc[4], nneighbors + o0, (size_t)nc))
            return 1;
        h->prior_o0 = o0;
        if (fzb_generic_gather_fit_dev(h, h->obj_in[0].as<double>(), h->obj_in[1].as<double>(), h->obj_in[2].as<double>(),
                                       nc, W, h->misc[3].as<int64_t>(), h->misc[4].as<int64_t>(), *cfg, d_o[0], d_o[1],
                                       d_o[2], d_nd, d_o[3], d_o[4], d_o[5]))
            return 1;
        for (int i = 0; i < 6; ++i)
            if (download(h, hostp[i] ? hostp[i] + (size_t)o0 * W : nullptr, d_o[i], no)) return 1;
        if (download(h, out->Ndim ? out->Ndim + (size_t)o0 * W : nullptr, d_nd, no)) return 1;
        FZB_CUDA(cudaStreamSynchronize(h->stream));
    }
    consume_prior_bins(h);
    return t.stop();
}

// ---- population likelihood of an N(z) given the PDFs (samplers.py:24-76) ------------------------------------------------
int fzb_nz_set_pdfs(fzb_handle h, const double* pdfs, int64_t No, int32_t Ng) {
    if (use_device(h)) return 2;
    FZB_CHECK(pdfs && No > 0 && Ng > 0, "bad arguments");
    if (upload(h, h->nz_pdfs, pdfs, (size_t)No * Ng)) return 1;
    FZB_CUDA(cudaStreamSynchronize(h->stream));
    h->nz_pdfs_ptr = h->nz_pdfs.as<double>();
    h->nz_No = No;
    h->nz_Ng = Ng;
    return 0;
}

int fzb_nz_set_pdfs_dev(fzb_handle h, const double* d_pdfs, int64_t No, int32_t Ng) {
    if (use_device(h)) return 2;
    FZB_CHECK(d_pdfs && No > 0 && Ng > 0, "bad arguments");
    h->nz_pdfs_ptr = d_pdfs;
    h->nz_No = No;
    h->nz_Ng = Ng;
    return 0;
}

int fzb_nz_loglike(fzb_handle h, const double* nz, int32_t Ng, int32_t pair_i, int32_t pair_j, double pair_step,
                   double* lnlike, double* overlap) {
    if (use_device(h)) return 2;
    FZB_CHECK(nz && lnlike, "null argument");
    FZB_CHECK(h->nz_pdfs_ptr && h->nz_No > 0, "call fzb_nz_set_pdfs first");
    FZB_CHECK(Ng == h->nz_Ng, "nz has %d bins, the PDFs have %d", Ng, h->nz_Ng);
    const bool pair = pair_i >= 0 && pair_j >= 0;
    if (pair) FZB_CHECK(pair_i < Ng && pair_j < Ng, "pair (%d, %d) outside the %d bins", pair_i, pair_j, Ng);
    reset_stats(h);
    // samplers.py:63-65: a negative or non-finite N(z) has zero likelihood
    for (int g = 0; g < Ng; ++g)
        if (!std::isfinite(nz[g]) || nz[g] < 0.0) {
            *lnlike = -INFINITY;
            if (overlap) memset(overlap, 0, (size_t)h->nz_No * sizeof(double));
            return 0;
        }
    return fzb_nz_loglike_impl(h, h->nz_pdfs_ptr, h->nz_No, Ng, nz, pair ? pair_i : -1, pair ? pair_j : -1,
                               pair ? pair_step : 0.0, lnlike, overlap);
}

}  // extern "C"

// ---- PDF summaries (pdf.py:899-1074) --------------------------------------------------------------------------------
int fzb_pdfs_summarize(fzb_handle h, const double* pdfs, const double* pgrid, const double* loss, const double* urand,
                       int64_t No, int32_t Ng, int32_t renormalize, double* rowsum, double* est, double* sd, double* risk,
                       double* quant, double* mc) {
    if (use_device(h)) return 2;
    FZB_CHECK(pdfs && pgrid && loss && urand && est && sd && risk && quant && mc && No > 0, "null argument");
    return fzb_summarize_impl(h, pdfs, pgrid, loss, urand, No, Ng, renormalize, rowsum, est, sd, risk, quant, mc);
}

int fzb_pdfs_conf(fzb_handle h, const double* points, const double* widths, int64_t No, double* conf) {
    if (use_device(h)) return 2;
    FZB_CHECK(points && widths && conf && No > 0, "null argument");
    return fzb_conf_impl(h, points, widths, No, conf);
}

// ---- page-locked host buffers for large outputs ---------------------------------------------------------------------
int fzb_alloc_pinned(size_t bytes, void** out) {
    FZB_CHECK(out != nullptr && bytes > 0, "bad arguments");
    *out = nullptr;
    FZB_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return 0;
}

int fzb_free_pinned(void* p) {
    if (p) FZB_CUDA(cudaFreeHost(p));
    return 0;
}

// ---- host helper: the reference's in-place cleaning (pdf.py:310-311) for float64 arrays, on several threads ------------
int fzb_clean_inplace_f64(double* data, double* err, double* mask, int64_t n) {
    FZB_CHECK(data && err && mask && n >= 0, "bad arguments");
    const int nt = (int)std::min<int64_t>(std::max<int64_t>(1, n >> 18), std::max(1u, std::thread::hardware_concurrency() / 2));
    auto work = [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const double d = data[i], e = err[i];
            if (!(std::isfinite(d) && std::isfinite(e) && e > 0.0)) { data[i] = 0.0; err[i] = 1.0; mask[i] = 0.0; }
        }
    };
    if (nt <= 1) { work(0, n); return 0; }
    std::vector<std::thread> pool;
    const int64_t per = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) pool.emplace_back(work, std::min<int64_t>(n, t * per), std::min<int64_t>(n, (t + 1) * per));
    for (auto& th : pool) th.join();
    return 0;
}
