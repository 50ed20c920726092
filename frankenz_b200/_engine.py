"""Thin host-side driver of one libfzb200 handle (numpy in / numpy out).

Only marshals arguments across the C ABI; all arithmetic of the path runs in the CUDA
library.  `Engine` is shared by `BruteForce`, `NearestNeighbors` and `pdf.loglike`.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import FzbConfig, FzbFitOut, FzbStats, dptr, f64, iptr

_LPROB_KEYS = ("free_scale", "ignore_model_err", "dim_prior", "ltol", "return_scale", "lnprior", "lnprior_bin",
               "precision")


def default_device():
    for key in ("FZB_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key, "") != "":
            return int(os.environ[key])
    return 0


def make_config(lprob_kwargs=None, kde_kwargs=None, track_scale=False):
    """Translate the reference's keyword arguments (pdf.py:238-240, :444-445, :529-531) into FzbConfig."""
    lk = dict(lprob_kwargs or {})
    kk = dict(kde_kwargs or {})
    unknown = [k for k in lk if k not in _LPROB_KEYS]
    if unknown:
        raise TypeError("unsupported lprob_kwargs for the built-in logprob: %s" % unknown)
    cfg = FzbConfig()
    cfg.free_scale = 1 if lk.get("free_scale", False) else 0
    ime = lk.get("ignore_model_err", False)
    # pdf.py:197 tests `ignore_model_err is not True`: a truthy non-bool still iterates
    cfg.ignore_model_err = 1 if ime is True else (2 if ime else 0)
    cfg.dim_prior = 1 if lk.get("dim_prior", True) else 0
    cfg.track_scale = 1 if track_scale else 0
    cfg.ltol = float(lk.get("ltol", 1e-4))
    wt, cdf = kk.get("wt_thresh", 1e-3), kk.get("cdf_thresh", 2e-4)
    cfg.use_wt_thresh = 0 if wt is None else 1
    cfg.use_cdf_thresh = 0 if cdf is None else 1
    cfg.wt_thresh = 0.0 if wt is None else float(wt)
    cfg.cdf_thresh = 0.0 if cdf is None else float(cdf)
    prec = lk.get("precision", "auto")
    cfg.precision = {"auto": _lib.PREC_AUTO, "fp64": _lib.PREC_FP64, "fp32": _lib.PREC_FP32}[prec]
    return cfg


def clean_inplace(data, data_err, data_mask):
    """Vectorised form of the reference's per-object in-place cleaning (pdf.py:310-311).

    The reference mutates the caller's arrays one row at a time; the drop-in does the same
    for all rows at once so callers observe identical arrays afterwards.
    """
    arrs = (data, data_err, data_mask)
    if (data.size >= (1 << 18) and all(isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous
                                      and a.flags.writeable and a.shape == data.shape for a in arrs)):
        try:      # large float64 batches: the same rule on several host threads
            _lib.check(_lib.load().fzb_clean_inplace_f64(dptr(data), dptr(data_err), dptr(data_mask), data.size))
            return data, data_err, data_mask
        except Exception:
            pass
    with np.errstate(invalid="ignore"):
        bad = ~(np.isfinite(data) & np.isfinite(data_err) & (data_err > 0.))
    if bad.any():
        data[bad], data_err[bad], data_mask[bad] = 0., 1., False
    return data, data_err, data_mask


class _PinnedBlock(object):
    """A page-locked host buffer exposed through the array interface; goes back to the pool when the last numpy view of
    it is gone."""

    def __init__(self, pool, ptr, nbytes, shape):
        self._pool, self._ptr, self._nbytes = pool, ptr, nbytes
        self.__array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def __del__(self):
        try:
            self._pool._release(self._ptr, self._nbytes)
        except Exception:
            pass


class PinnedPool(object):
    """Pool of page-locked output buffers (large PDF arrays): the device-to-host copy engine writes them directly, which
    takes the staging copy (and its host memory traffic, the limit of several ranks on one host) out of `fit_predict`.
    Pinning is slow (~1 s per few GB), so buffers are recycled once the arrays handed out are garbage collected; at most
    `max_bytes` stay pinned, beyond that (or on failure) callers get ordinary pageable arrays."""
    MIN_BYTES = 64 << 20

    def __init__(self, max_bytes=None):
        # measured on B200 hosts: one process per host is ~4 % faster through the staged path (its host-side copy hides
        # behind the kernels), four ranks on one host are 12 % faster with pinned outputs; FZB_PINNED_POOL_BYTES overrides
        multi = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1) > 1
        dflt = (24 << 30) if multi else 0
        self.max_bytes = int(os.environ.get("FZB_PINNED_POOL_BYTES", dflt)) if max_bytes is None else max_bytes
        self.free, self.total = [], 0

    def empty(self, shape):
        nbytes = int(np.prod(shape)) * 8
        if nbytes < self.MIN_BYTES or self.max_bytes <= 0:
            return np.empty(shape)
        for i, (ptr, cap) in enumerate(self.free):
            if cap >= nbytes and cap <= 2 * nbytes:
                self.free.pop(i)
                return np.asarray(_PinnedBlock(self, ptr, cap, shape))
        lib = _lib.load()
        while self.total + nbytes > self.max_bytes and self.free:      # make room: drop idle buffers of other sizes
            ptr, cap = self.free.pop()
            lib.fzb_free_pinned(ptr)
            self.total -= cap
        if self.total + nbytes > self.max_bytes:
            return np.empty(shape)
        p = C.c_void_p()
        if lib.fzb_alloc_pinned(nbytes, C.byref(p)) != 0 or not p.value:
            return np.empty(shape)
        self.total += nbytes
        return np.asarray(_PinnedBlock(self, p.value, nbytes, shape))

    def _release(self, ptr, nbytes):
        self.free.append((ptr, nbytes))


_pinned_pool = PinnedPool()


class _PageBlock(object):
    """A recycled pageable buffer behind the array interface (see PagePool)."""

    def __init__(self, pool, base, shape):
        self._pool, self._base = pool, base
        self.__array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (base.ctypes.data, False), "version": 3}

    def __del__(self):
        try:
            self._pool._release(self._base)
        except Exception:
            pass


class PagePool(object):
    """Recycles the large PAGEABLE output arrays of `fit_predict` between calls.  A freshly allocated numpy array has never
    been touched: the staged device-to-host copy then runs at the kernel's first-touch rate (every 4 KB page is zeroed at its
    first write; ~19 GB/s measured on a B200 host with 12 copy threads) instead of memcpy speed.  Buffers whose arrays have
    been garbage collected are handed out again (already mapped), at most `max_bytes` are kept (FZB_HOST_POOL_BYTES, default
    12 GB; 0 disables the pool).  The contents of a new array are undefined, exactly as with numpy.empty."""
    MIN_BYTES = 64 << 20

    def __init__(self, max_bytes=None):
        self.max_bytes = int(os.environ.get("FZB_HOST_POOL_BYTES", 12 << 30)) if max_bytes is None else max_bytes
        self.free, self.total = [], 0

    def empty(self, shape):
        nbytes = int(np.prod(shape)) * 8
        if nbytes < self.MIN_BYTES or self.max_bytes <= 0:
            return np.empty(shape)
        for i, base in enumerate(self.free):
            if base.nbytes >= nbytes and base.nbytes <= 2 * nbytes:
                self.free.pop(i)
                return np.asarray(_PageBlock(self, base, shape))
        while self.total + nbytes > self.max_bytes and self.free:       # make room: drop idle buffers of other sizes
            self.total -= self.free.pop().nbytes
        if self.total + nbytes > self.max_bytes:
            return np.empty(shape)
        base = np.empty(nbytes, dtype=np.uint8)
        self.total += nbytes
        return np.asarray(_PageBlock(self, base, shape))

    def _release(self, base):
        self.free.append(base)


_page_pool = PagePool()


def _out_array(shape):
    """Large float64 output array: page-locked (several ranks per host), recycled pageable, or plain numpy.empty."""
    a = _pinned_pool.empty(shape)
    if _pinned_pool.max_bytes > 0:
        return a
    return _page_pool.empty(shape)


class Engine(object):
    def __init__(self, models, models_err, models_mask, device=None):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        self.device = default_device() if device is None else int(device)
        _lib.check(self.lib.fzb_create(self.device, C.byref(self.h)))
        m, me, mm = f64(models), f64(models_err), f64(models_mask)
        if m.ndim != 2 or me.shape != m.shape or mm.shape != m.shape:
            raise ValueError("models, models_err and models_mask must share one (Nmodel, Nfilt) shape")
        self.Nm, self.Nf = m.shape
        _lib.check(self.lib.fzb_set_models(self.h, dptr(m), dptr(me), dptr(mm), self.Nm, self.Nf))
        self._kde_key = None
        self.Ng = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            try:
                self.lib.fzb_destroy(self.h)
            except Exception:
                pass
            self.h = None

    __del__ = close

    # ---- configuration -------------------------------------------------------------------------
    def set_lnprior(self, lnprior, bins=None):
        """Per-model ln-prior [Nmodel], or an object-conditioned table [Nbins, Nmodel] with `bins` [Ndata] giving the
        row of every object of the next fit call (the built-in replacement for a prior-carrying Python lprob_func)."""
        _lib.check(self.lib.fzb_set_lnprior_table(self.h, None, 0, 0))
        if lnprior is not None and np.ndim(lnprior) == 2:
            if bins is None:
                raise ValueError("a 2-D `lnprior` table needs `lnprior_bin` (one row index per object)")
            tab = f64(lnprior)
            if tab.shape[1] != self.Nm:
                raise ValueError("lnprior table must have shape (Nbins, Nmodel)")
            b = np.ascontiguousarray(bins, dtype=np.int32)
            _lib.check(self.lib.fzb_set_lnprior(self.h, None, 0))
            _lib.check(self.lib.fzb_set_lnprior_table(self.h, dptr(tab), tab.shape[0], self.Nm))
            _lib.check(self.lib.fzb_set_object_prior_bins(self.h, b.ctypes.data_as(_lib.c_int32_p), len(b)))
            return
        if lnprior is None:
            _lib.check(self.lib.fzb_set_lnprior(self.h, None, 0))
        else:
            lp = f64(lnprior)
            if lp.shape != (self.Nm,):
                raise ValueError("lnprior must have shape (Nmodel,)")
            _lib.check(self.lib.fzb_set_lnprior(self.h, dptr(lp), self.Nm))

    def set_kde_dict_idx(self, label_dict, y_idx, y_std_idx):
        """Dictionary KDE from ready-made indices (the `y_idx` / `y_std_idx` form of pdf.py:529-531)."""
        widths = np.ascontiguousarray(label_dict.sigma_width, dtype=np.int32)
        lens = np.array([len(k) for k in label_dict.sigma_dict], dtype=np.int64)
        koff = np.zeros(len(lens) + 1, dtype=np.int64)
        koff[1:] = np.cumsum(lens)
        kern = f64(np.concatenate(label_dict.sigma_dict)) if koff[-1] else np.zeros(1)
        kcdf = f64(np.concatenate(label_dict.sigma_dict_cdf)) if koff[-1] else np.zeros(1)
        _lib.check(self.lib.fzb_set_kde_dict(self.h, int(label_dict.Ngrid), int(label_dict.Ndict),
                                             widths.ctypes.data_as(_lib.c_int32_p), iptr(koff), dptr(kern),
                                             dptr(kcdf)))
        yi = np.ascontiguousarray(y_idx, dtype=np.int64)
        si = np.ascontiguousarray(y_std_idx, dtype=np.int64)
        _lib.check(self.lib.fzb_set_labels_dict(self.h, iptr(yi), iptr(si), len(yi)))
        self.Ng = int(label_dict.Ngrid)

    def set_kde(self, model_labels, model_label_errs, label_dict=None, label_grid=None, kde_kwargs=None):
        """Upload the KDE tables and per-model label indices (pdf.py:800-852 / :499-502)."""
        if label_dict is None and label_grid is None:
            raise ValueError("`label_dict` or `label_grid` must be specified.")
        kk = kde_kwargs or {}
        y, ye = f64(model_labels), f64(model_label_errs)
        if label_dict is not None:
            yi, si = label_dict.fit(y, ye)
            return self.set_kde_dict_idx(label_dict, yi, si)
        x = f64(label_grid)
        nx = len(x)
        dx = kk.get("dx", None)
        if dx is None:
            dx = x[1] - x[0]
        sig = kk.get("sig_thresh", 5.)
        # pdf.py:499-502: truncation toward zero, upper-exclusive windows clipped to the grid
        centers = np.array((y - x[0]) / dx, dtype="int")
        offsets = np.array(sig * ye / dx, dtype="int")
        uppers, lowers = centers + offsets, centers - offsets
        uppers[uppers > nx], lowers[lowers < 0] = nx, 0
        # python slice semantics of x[lower:upper]: a negative upper bound counts from the end
        lowers = np.minimum(lowers, nx)
        uppers = np.where(uppers < 0, np.maximum(uppers + nx, 0), uppers)
        uppers = np.maximum(uppers, lowers)
        _lib.check(self.lib.fzb_set_kde_grid(self.h, dptr(x), nx))
        lo = np.ascontiguousarray(lowers, dtype=np.int64)
        up = np.ascontiguousarray(uppers, dtype=np.int64)
        _lib.check(self.lib.fzb_set_labels_grid(self.h, dptr(y), dptr(ye), iptr(lo), iptr(up), len(y)))
        self.Ng = nx

    # ---- compute ------------------------------------------------------------------------------
    @staticmethod
    def _objects(data, data_err, data_mask):
        x, xe, xm = f64(data), f64(data_err), f64(data_mask)
        if x.ndim != 2 or xe.shape != x.shape or xm.shape != x.shape:
            raise ValueError("data, data_err and data_mask must share one (Ndata, Nfilt) shape")
        return x, xe, xm

    def fit(self, data, data_err, data_mask, cfg, want=("lnprior", "lnlike", "lnprob", "Ndim", "chi2", "scale",
                                                       "scale_err"), out=None):
        """Full (Ndata x Nmodel) fit arrays (bruteforce.py:182-205)."""
        x, xe, xm = self._objects(data, data_err, data_mask)
        if x.shape[1] != self.Nf:
            raise ValueError("data has %d filters, models have %d" % (x.shape[1], self.Nf))
        no = len(x)
        res = dict(out or {})
        for name in want:
            if name not in res:
                res[name] = np.empty((no, self.Nm), dtype=np.int64 if name == "Ndim" else np.float64)
        o = FzbFitOut()
        for name in ("lnprior", "lnlike", "lnprob", "chi2", "scale", "scale_err"):
            setattr(o, name, dptr(res.get(name)))
        o.Ndim = iptr(res.get("Ndim"))
        _lib.check(self.lib.fzb_fit(self.h, dptr(x), dptr(xe), dptr(xm), no, C.byref(cfg), C.byref(o)))
        return res

    def fit_predict(self, data, data_err, data_mask, cfg, want_pdf=True):
        """Fused fit + PDF (bruteforce.py:602-631, save_fits=False)."""
        x, xe, xm = self._objects(data, data_err, data_mask)
        if x.shape[1] != self.Nf:
            raise ValueError("data has %d filters, models have %d" % (x.shape[1], self.Nf))
        no = len(x)
        pdfs = _out_array((no, self.Ng)) if want_pdf else None
        lmap, levid = np.empty(no), np.empty(no)
        best = np.empty(no, dtype=np.int64)
        bchi2, bscale = np.empty(no), np.empty(no)
        _lib.check(self.lib.fzb_fit_predict(self.h, dptr(x), dptr(xe), dptr(xm), no, C.byref(cfg), dptr(pdfs),
                                            dptr(lmap), dptr(levid), iptr(best), dptr(bchi2), dptr(bscale)))
        return pdfs, lmap, levid, best, bchi2, bscale

    def fit_predict_summarize(self, data, data_err, data_mask, cfg, pgrid, loss, urand, renormalize=True,
                              wconf_frac=0.03, want_pdf=False):
        """Fused fit + PDF + pdfs_summarize: the PDFs are summarised on the device; they are
        downloaded only with want_pdf."""
        x, xe, xm = self._objects(data, data_err, data_mask)
        if x.shape[1] != self.Nf:
            raise ValueError("data has %d filters, models have %d" % (x.shape[1], self.Nf))
        no = len(x)
        pgrid, loss, urand = f64(pgrid), f64(loss), f64(urand)
        if len(pgrid) != self.Ng or loss.shape != (self.Ng, self.Ng) or urand.shape != (no,):
            raise ValueError("summary tables do not match the PDF grid / the number of objects")
        pdfs = _out_array((no, self.Ng)) if want_pdf else None
        # one recycled block for the 25 per-object float64 outputs (200 bytes per object): a fresh numpy.empty per array
        # would be first-touched, page by page, by the download
        blk = _page_pool.empty((25, no))
        est, sd, conf, risk, quant = (blk[4 * i:4 * i + 4] for i in range(5))
        mc, lmap, levid, bchi2, bscale = (blk[20 + i] for i in range(5))
        best = np.empty(no, dtype=np.int64)
        _lib.check(self.lib.fzb_fit_predict_summarize(
            self.h, dptr(x), dptr(xe), dptr(xm), no, C.byref(cfg), dptr(pgrid), dptr(loss), dptr(urand),
            1 if renormalize else 0, float(wconf_frac), dptr(pdfs), dptr(lmap), dptr(levid), iptr(best), dptr(bchi2),
            dptr(bscale), dptr(est), dptr(sd), dptr(conf), dptr(risk), dptr(quant), dptr(mc)))
        summary = tuple((est[k], sd[k], conf[k], risk[k]) for k in range(4)) + ((quant[0], quant[1], quant[2], quant[3]), mc)
        return summary, pdfs, lmap, levid, best, bchi2, bscale

    def predict_logwt(self, logwt, cfg, neighbors=None, nneighbors=None):
        """PDFs from a log-weight matrix (bruteforce.py:358-372; knn.py:541-555 with neighbours)."""
        lw = f64(logwt)
        no, w = lw.shape
        nb = nn = None
        if neighbors is not None:
            nb = np.ascontiguousarray(neighbors, dtype=np.int64)
            nn = np.ascontiguousarray(nneighbors, dtype=np.int64)
        pdfs, lmap, levid = np.empty((no, self.Ng)), np.empty(no), np.empty(no)
        _lib.check(self.lib.fzb_predict_logwt(self.h, dptr(lw), no, w, iptr(nb), iptr(nn), C.byref(cfg), dptr(pdfs),
                                              dptr(lmap), dptr(levid)))
        return pdfs, lmap, levid

    def knn_build(self, feats):
        f = np.ascontiguousarray(feats, dtype=np.float32)
        K, nm, nf = f.shape
        _lib.check(self.lib.fzb_knn_build(self.h, f.ctypes.data_as(_lib.c_float_p), K, nm, nf))
        self.K = K

    def knn_query(self, qfeats, k, p=2, return_dist=True):
        q = f64(np.atleast_2d(qfeats))
        no = len(q)
        idx = np.empty((no, self.K, k), dtype=np.int64)
        dist = np.empty((no, self.K, k)) if return_dist else None
        pp = 0.0 if np.isinf(p) else float(p)
        _lib.check(self.lib.fzb_knn_query(self.h, dptr(q), no, k, pp, iptr(idx), dptr(dist)))
        return idx, dist

    def knn_fit(self, qfeats, data, data_err, data_mask, k, p, cfg):
        x, xe, xm = self._objects(data, data_err, data_mask)
        q = f64(qfeats)
        no = len(x)
        w = self.K * k
        res = dict(lnprior=np.empty((no, w)), lnlike=np.empty((no, w)), lnprob=np.empty((no, w)),
                   Ndim=np.empty((no, w), dtype=np.int64), chi2=np.empty((no, w)), scale=np.empty((no, w)),
                   scale_err=np.empty((no, w)))
        nb = np.empty((no, w), dtype=np.int64)
        nn = np.empty(no, dtype=np.int64)
        o = FzbFitOut()
        for name in ("lnprior", "lnlike", "lnprob", "chi2", "scale", "scale_err"):
            setattr(o, name, dptr(res[name]))
        o.Ndim = iptr(res["Ndim"])
        pp = 0.0 if np.isinf(p) else float(p)
        _lib.check(self.lib.fzb_knn_fit(self.h, dptr(q), dptr(x), dptr(xe), dptr(xm), no, k, pp, C.byref(cfg),
                                        iptr(nb), iptr(nn), C.byref(o)))
        res["neighbors"], res["Nneighbors"] = nb, nn
        return res

    def knn_fit_predict(self, qfeats, data, data_err, data_mask, k, p, cfg):
        """Search, union, fits and KDE on the device; only PDFs / lmap / levid / Nneighbors come back."""
        x, xe, xm = self._objects(data, data_err, data_mask)
        q = f64(qfeats)
        no = len(x)
        pdfs = _out_array((no, self.Ng))
        lmap, levid = np.empty(no), np.empty(no)
        nn = np.empty(no, dtype=np.int64)
        pp = 0.0 if np.isinf(p) else float(p)
        _lib.check(self.lib.fzb_knn_fit_predict(self.h, dptr(q), dptr(x), dptr(xe), dptr(xm), no, k, pp, C.byref(cfg),
                                                dptr(pdfs), dptr(lmap), dptr(levid), iptr(nn)))
        return pdfs, lmap, levid, nn

    def fit_gather(self, data, data_err, data_mask, neighbors, nneighbors, cfg):
        """Fits of every object to its own list of models (networks.py:918-923); arrays padded to the widest list."""
        x, xe, xm = self._objects(data, data_err, data_mask)
        nb = np.ascontiguousarray(neighbors, dtype=np.int64)
        nn = np.ascontiguousarray(nneighbors, dtype=np.int64)
        no, w = nb.shape
        res = dict(lnprior=np.empty((no, w)), lnlike=np.empty((no, w)), lnprob=np.empty((no, w)),
                   Ndim=np.empty((no, w), dtype=np.int64), chi2=np.empty((no, w)), scale=np.empty((no, w)),
                   scale_err=np.empty((no, w)))
        o = FzbFitOut()
        for name in ("lnprior", "lnlike", "lnprob", "chi2", "scale", "scale_err"):
            setattr(o, name, dptr(res[name]))
        o.Ndim = iptr(res["Ndim"])
        _lib.check(self.lib.fzb_fit_gather(self.h, dptr(x), dptr(xe), dptr(xm), no, w, iptr(nb), iptr(nn), C.byref(cfg),
                                           C.byref(o)))
        return res

    def measure_peaks(self, reps=5):
        """FP32-FMA (TFLOP/s) and MUFU (Gop/s) peaks of this device, measured with dependency-free loops."""
        a, b = C.c_double(), C.c_double()
        _lib.check(self.lib.fzb_measure_peaks(self.h, reps, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fit_predict_dev(self, d_x, d_xe, d_xm, no, cfg, d_pdfs, d_lmap, d_levid, d_best, d_bchi2, d_bscale):
        """Device-pointer form of `fit_predict` (integers = CUDA device addresses on this handle's device)."""
        _lib.check(self.lib.fzb_fit_predict_dev(self.h, d_x, d_xe, d_xm, no, C.byref(cfg), d_pdfs, d_lmap, d_levid,
                                                d_best, d_bchi2, d_bscale))

    def stats(self):
        s = FzbStats()
        _lib.check(self.lib.fzb_get_stats(self.h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in FzbStats._fields_}


class SummaryEngine(object):
    """Handle for the model-free entry points (PDF summaries)."""
    _cache = {}

    def __init__(self, device=None):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        self.device = default_device() if device is None else int(device)
        _lib.check(self.lib.fzb_create(self.device, C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            try:
                self.lib.fzb_destroy(self.h)
            except Exception:
                pass
            self.h = None

    @classmethod
    def get(cls, device=None):
        dev = default_device() if device is None else int(device)
        if dev not in cls._cache:
            cls._cache[dev] = cls(dev)
        return cls._cache[dev]

    def summarize(self, pdfs, pgrid, loss, urand, renormalize):
        """Stage 1 of pdfs_summarize (pdf.py:978-1036, :1064-1068): returns est, std, risk, quant ([4, No]), mc, rowsum."""
        no, ng = pdfs.shape
        est, sd, risk, quant = (np.empty((4, no)) for _ in range(4))
        mc, rowsum = np.empty(no), np.empty(no)
        _lib.check(self.lib.fzb_pdfs_summarize(self.h, dptr(pdfs), dptr(pgrid), dptr(loss), dptr(urand), no, ng,
                                               1 if renormalize else 0, dptr(rowsum), dptr(est), dptr(sd), dptr(risk),
                                               dptr(quant), dptr(mc)))
        return est, sd, risk, quant, mc, rowsum

    def conf(self, points, widths):
        """Stage 2 (pdf.py:1038-1062): probability within +-width of each estimator."""
        points, widths = f64(points), f64(widths)
        out = np.empty_like(points)
        _lib.check(self.lib.fzb_pdfs_conf(self.h, dptr(points), dptr(widths), points.shape[1], dptr(out)))
        return out

    def stats(self):
        s = FzbStats()
        _lib.check(self.lib.fzb_get_stats(self.h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in FzbStats._fields_}
