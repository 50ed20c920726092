"""`BruteForce` estimator with the reference's interface (frankenz/bruteforce.py:30-631).

Every object is scored against every model on the GPU.  The per-object Python loop of the
reference is replaced by batched calls into libfzb200; the generator twins (`_fit`,
`_predict`, `_fit_predict`) still yield one object at a time so streaming callers keep
working, but the work is done up front.
"""
import sys

import numpy as np

from . import pdf as _pdf
from ._engine import Engine, clean_inplace, make_config

__all__ = ["BruteForce"]


def _check_lprob_func(lprob_func):
    if lprob_func is None or lprob_func is _pdf.logprob:
        return
    raise NotImplementedError(
        "frankenz_b200 evaluates the likelihood inside CUDA kernels and cannot call a Python `lprob_func`. "
        "Use lprob_func=None with `lprob_kwargs` (free_scale, ignore_model_err, dim_prior, ltol) and pass a "
        "per-model log-prior as lprob_kwargs['lnprior'] (uniform if omitted).")


def _check_args(args, name):
    if args is not None and len(args) > 0:
        raise NotImplementedError("positional `%s` are not supported; use the keyword form" % name)


# object-conditioned prior tables: from this many object-model pairs on, fit_predict(save_fits=False) goes through the
# fused kernels one table row at a time instead of the float64 kernel that reads the table per pair
TABLE_PRIOR_GROUP_MIN_PAIRS = 1e8
# generator twins with save_fits=False: host bytes of fit arrays / PDFs alive at a time
STREAM_BYTES = 1 << 30


class BruteForce(object):
    """Fits data and generates predictions with a brute-force scan of all models."""

    def __init__(self, models, models_err, models_mask):
        # the reference keeps references to the caller's arrays (bruteforce.py:54-56)
        self.models = models
        self.models_err = models_err
        self.models_mask = models_mask
        self.NMODEL, self.NDIM = models.shape
        self.NDATA = None
        self.fit_lnprior = None
        self.fit_lnlike = None
        self.fit_lnprob = None
        self.fit_Ndim = None
        self.fit_chi2 = None
        self.fit_scale = None
        self.fit_scale_err = None
        self._engine = None
        self._lnprior_id = None

    # ---- internals -----------------------------------------------------------------------------
    def _eng(self):
        if self._engine is None:
            self._engine = Engine(self.models, self.models_err, self.models_mask)
        return self._engine

    def _setup(self, lprob_func, lprob_args, lprob_kwargs, track_scale, kde_kwargs=None):
        _check_lprob_func(lprob_func)
        _check_args(lprob_args, "lprob_args")
        lk = dict(lprob_kwargs or {})
        if track_scale and not (lk.get("free_scale", False) and lk.get("return_scale", False)):
            # the reference indexes results[5] of a 5-tuple here (bruteforce.py:200-202)
            raise IndexError("tuple index out of range: `track_scale` needs lprob_kwargs free_scale=True and "
                             "return_scale=True")
        eng = self._eng()
        eng.set_lnprior(lk.get("lnprior", None), lk.get("lnprior_bin", None))
        return eng, make_config(lk, kde_kwargs, track_scale=track_scale)

    def _store(self, res, Ndata):
        self.fit_lnprior, self.fit_lnlike, self.fit_lnprob = res["lnprior"], res["lnlike"], res["lnprob"]
        self.fit_Ndim, self.fit_chi2 = res["Ndim"], res["chi2"]
        self.fit_scale, self.fit_scale_err = res["scale"], res["scale_err"]
        self.NDATA = Ndata

    @staticmethod
    def _rows(res, i, track_scale):
        out = (res["lnprior"][i], res["lnlike"][i], res["lnprob"][i], res["Ndim"][i], res["chi2"][i])
        if track_scale:
            out = out + (res["scale"][i], res["scale_err"][i])
        return out

    # ---- fit -----------------------------------------------------------------------------------
    def fit(self, data, data_err, data_mask, lprob_func=None, lprob_args=None, lprob_kwargs=None,
            track_scale=False, verbose=True):
        """Fit all models to all objects; results land in the `fit_*` attributes (bruteforce.py:66-125)."""
        Ndata = len(data)
        for i, _ in enumerate(self._fit(data, data_err, data_mask, lprob_func=lprob_func, lprob_args=lprob_args,
                                        lprob_kwargs=lprob_kwargs, track_scale=track_scale, save_fits=True)):
            pass
        if verbose:
            sys.stderr.write('\rFitting object {0}/{1}\n'.format(Ndata, Ndata))
            sys.stderr.flush()

    def _fit(self, data, data_err, data_mask, lprob_func=None, lprob_args=None, lprob_kwargs=None,
             track_scale=False, save_fits=True):
        """Generator over objects yielding the `logprob` tuple of each (bruteforce.py:127-205)."""
        eng, cfg = self._setup(lprob_func, lprob_args, lprob_kwargs, track_scale)
        clean_inplace(data, data_err, data_mask)
        Ndata = len(data)
        self.NDATA = Ndata
        if save_fits:
            res = eng.fit(data, data_err, data_mask, cfg)
            self._store(res, Ndata)
            for i in range(Ndata):
                yield self._rows(res, i, track_scale)
            return
        # save_fits=False: the reference streams one object at a time (bruteforce.py:192-205); here the objects go
        # through the device in chunks whose seven (chunk x Nmodel) arrays stay within STREAM_BYTES
        lk = dict(lprob_kwargs or {})
        chunk = max(1, int(STREAM_BYTES // (7 * 8 * self.NMODEL)))
        for o0 in range(0, Ndata, chunk):
            o1 = min(Ndata, o0 + chunk)
            if lk.get("lnprior_bin", None) is not None and (o0 > 0 or o1 < Ndata):
                eng.set_lnprior(lk.get("lnprior", None), np.asarray(lk["lnprior_bin"])[o0:o1])
            res = eng.fit(data[o0:o1], data_err[o0:o1], data_mask[o0:o1], cfg)
            for i in range(o1 - o0):
                yield self._rows(res, i, track_scale)

    # ---- predict -------------------------------------------------------------------------------
    def predict(self, model_labels, model_label_errs, label_dict=None, label_grid=None, logwt=None, kde_args=None,
                kde_kwargs=None, return_gof=False, verbose=True):
        """1-D PDFs from stored fits or supplied log-weights (bruteforce.py:207-301)."""
        pdfs, lmap, levid = self._predict_all(model_labels, model_label_errs, label_dict, label_grid, logwt,
                                              kde_args, kde_kwargs)
        if verbose:
            sys.stderr.write('\rGenerating PDF {0}/{1}\n'.format(len(pdfs), len(pdfs)))
            sys.stderr.flush()
        if return_gof:
            return pdfs, (lmap, levid)
        return pdfs

    def _predict_all(self, model_labels, model_label_errs, label_dict, label_grid, logwt, kde_args, kde_kwargs):
        _check_args(kde_args, "kde_args")
        if logwt is None:
            logwt = self.fit_lnprob
        if label_dict is None and label_grid is None:
            raise ValueError("`label_dict` or `label_grid` must be specified.")
        if logwt is None:
            raise ValueError("Fits have not been computed and weights have not been provided.")
        eng = self._eng()
        eng.set_kde(model_labels, model_label_errs, label_dict=label_dict, label_grid=label_grid,
                    kde_kwargs=kde_kwargs)
        cfg = make_config(None, kde_kwargs)
        return eng.predict_logwt(logwt, cfg)

    def _predict(self, model_labels, model_label_errs, label_dict=None, label_grid=None, logwt=None, kde_args=None,
                 kde_kwargs=None):
        """Generator twin of `predict` (bruteforce.py:303-372): yields (pdf, (lmap, levid))."""
        pdfs, lmap, levid = self._predict_all(model_labels, model_label_errs, label_dict, label_grid, logwt,
                                              kde_args, kde_kwargs)
        for i in range(len(pdfs)):
            yield pdfs[i], (lmap[i], levid[i])

    # ---- fit_predict ---------------------------------------------------------------------------
    def fit_predict(self, data, data_err, data_mask, model_labels, model_label_errs, lprob_func=None,
                    label_dict=None, label_grid=None, kde_args=None, kde_kwargs=None, lprob_args=None,
                    lprob_kwargs=None, return_gof=False, track_scale=False, verbose=True, save_fits=True,
                    summarize=False, return_pdfs=True, summarize_kwargs=None):
        """Fit and predict in one go (bruteforce.py:374-503).  With `save_fits=False` the
        (Ndata x Nmodel) arrays are never formed: the fused kernels reduce on the fly.

        Extension (not in the reference): `summarize=True` also returns what `pdf.pdfs_summarize` would compute from the
        PDFs, evaluated on the device while they are still there (see `fit_predict_summarize`); with
        `return_pdfs=False` the PDFs themselves are not copied to the host."""
        if summarize:
            return self.fit_predict_summarize(data, data_err, data_mask, model_labels, model_label_errs,
                                              lprob_func=lprob_func, label_dict=label_dict, label_grid=label_grid,
                                              kde_args=kde_args, kde_kwargs=kde_kwargs, lprob_args=lprob_args,
                                              lprob_kwargs=lprob_kwargs, return_gof=return_gof, verbose=verbose,
                                              return_pdfs=return_pdfs, **(summarize_kwargs or {}))
        pdfs, lmap, levid = self._fit_predict_all(data, data_err, data_mask, model_labels, model_label_errs,
                                                  lprob_func, label_dict, label_grid, kde_args, kde_kwargs,
                                                  lprob_args, lprob_kwargs, track_scale, save_fits)
        if verbose:
            sys.stderr.write('\rGenerating PDF {0}/{1}\n'.format(len(pdfs), len(pdfs)))
            sys.stderr.flush()
        if return_gof:
            return pdfs, (lmap, levid)
        return pdfs

    def fit_predict_summarize(self, data, data_err, data_mask, model_labels, model_label_errs, lprob_func=None,
                              label_dict=None, label_grid=None, kde_args=None, kde_kwargs=None, lprob_args=None,
                              lprob_kwargs=None, return_gof=False, verbose=True, return_pdfs=False, renormalize=True,
                              rstate=None, pkern='lorentz', pkern_grid=None, wconf_frac=0.03):
        """`fit_predict(save_fits=False)` followed by `pdf.pdfs_summarize(pdfs, grid, ...)` (pdf.py:899-1074; demo 3
        cells 14 + 21) in one call: the point estimates, intervals and risks are computed on the device from PDFs that
        never leave it, so ~200 bytes per object cross PCIe instead of Ngrid x 8.

        Returns `summary` (the 6-tuple of pdfs_summarize: (mean, std, conf, risk), (median, ...), (mode, ...),
        (best, ...), (low95, low68, high68, high95), mc), followed by `pdfs` if `return_pdfs` and by `(lmap, levid)` if
        `return_gof`.  The random draws come from `rstate` one per object in order, like the reference's; the
        confidence width is the reference's default wconf_func, `wconf_frac * (1 + estimate)`."""
        _check_args(kde_args, "kde_args")
        if label_dict is None and label_grid is None:
            raise ValueError("`label_dict` or `label_grid` must be specified.")
        eng, cfg = self._setup(lprob_func, lprob_args, lprob_kwargs, False, kde_kwargs)
        eng.set_kde(model_labels, model_label_errs, label_dict=label_dict, label_grid=label_grid, kde_kwargs=kde_kwargs)
        clean_inplace(data, data_err, data_mask)
        pgrid = np.ascontiguousarray(label_dict.grid if label_dict is not None else label_grid, dtype=np.float64)
        if rstate is None:
            rstate = np.random
        nobj = len(data)
        urand = np.array([rstate.rand() for _ in range(nobj)]) if nobj < 64 else np.ascontiguousarray(rstate.rand(nobj))
        loss = _pdf._loss_matrix(pgrid, pkern, pkern_grid)
        summary, pdfs, lmap, levid, best, bchi2, bscale = eng.fit_predict_summarize(
            data, data_err, data_mask, cfg, pgrid, loss, urand, renormalize=renormalize, wconf_frac=wconf_frac,
            want_pdf=return_pdfs)
        self.best_idx, self.best_chi2, self.best_scale = best, bchi2, bscale
        self.NDATA = nobj
        if verbose:
            sys.stderr.write('\rGenerating PDF {0}/{1}\n'.format(nobj, nobj))
            sys.stderr.flush()
        out = (summary,)
        if return_pdfs:
            out = out + (pdfs,)
        if return_gof:
            out = out + ((lmap, levid),)
        return out[0] if len(out) == 1 else out

    def _fit_predict_all(self, data, data_err, data_mask, model_labels, model_label_errs, lprob_func, label_dict,
                         label_grid, kde_args, kde_kwargs, lprob_args, lprob_kwargs, track_scale, save_fits):
        _check_args(kde_args, "kde_args")
        if label_dict is None and label_grid is None:
            raise ValueError("`label_dict` or `label_grid` must be specified.")
        lk0 = dict(lprob_kwargs or {})
        if (not save_fits and lk0.get("lnprior", None) is not None and np.ndim(lk0["lnprior"]) == 2
                and float(len(data)) * self.NMODEL >= TABLE_PRIOR_GROUP_MIN_PAIRS):
            return self._fit_predict_by_prior_bin(data, data_err, data_mask, model_labels, model_label_errs, lprob_func,
                                                  label_dict, label_grid, kde_kwargs, lprob_args, lk0, track_scale)
        eng, cfg = self._setup(lprob_func, lprob_args, lprob_kwargs, track_scale, kde_kwargs)
        eng.set_kde(model_labels, model_label_errs, label_dict=label_dict, label_grid=label_grid,
                    kde_kwargs=kde_kwargs)
        clean_inplace(data, data_err, data_mask)
        Ndata = len(data)
        if save_fits:
            res = eng.fit(data, data_err, data_mask, cfg)
            self._store(res, Ndata)
            return eng.predict_logwt(res["lnprob"], cfg)
        pdfs, lmap, levid, best, bchi2, bscale = eng.fit_predict(data, data_err, data_mask, cfg)
        self.best_idx, self.best_chi2, self.best_scale = best, bchi2, bscale   # extras of the fused path
        return pdfs, lmap, levid

    def _fit_predict_by_prior_bin(self, data, data_err, data_mask, model_labels, model_label_errs, lprob_func, label_dict,
                                  label_grid, kde_kwargs, lprob_args, lk, track_scale):
        """Object-conditioned tabulated prior on the fused path: the objects of one table row share a
        per-model prior, which the sweep kernels carry in their model records, so the batch is processed row by row of the
        table (the float64 kernel that reads the table per pair remains the path of small problems and of `fit`)."""
        table, bins = np.asarray(lk["lnprior"], dtype=np.float64), lk.get("lnprior_bin", None)
        if bins is None:
            raise ValueError("a 2-D `lnprior` table needs `lnprior_bin` (one row index per object)")
        bins = np.asarray(bins)
        if bins.shape != (len(data),) or bins.min() < 0 or bins.max() >= len(table):
            raise ValueError("`lnprior_bin` must hold one row index of the table per object")
        clean_inplace(data, data_err, data_mask)
        out = None
        for b in np.unique(bins):
            idx = np.nonzero(bins == b)[0]
            lkb = dict(lk, lnprior=table[b])
            lkb.pop("lnprior_bin", None)
            eng, cfg = self._setup(lprob_func, lprob_args, lkb, track_scale, kde_kwargs)
            eng.set_kde(model_labels, model_label_errs, label_dict=label_dict, label_grid=label_grid, kde_kwargs=kde_kwargs)
            res = eng.fit_predict(data[idx], data_err[idx], data_mask[idx], cfg)
            if out is None:
                out = [np.empty((len(data),) + r.shape[1:], dtype=r.dtype) for r in res]
            for o, r in zip(out, res):
                o[idx] = r
        pdfs, lmap, levid, best, bchi2, bscale = out
        self.best_idx, self.best_chi2, self.best_scale = best, bchi2, bscale
        return pdfs, lmap, levid

    def _fit_predict(self, data, data_err, data_mask, model_labels, model_label_errs, lprob_func=None,
                     label_dict=None, label_grid=None, kde_args=None, kde_kwargs=None, lprob_args=None,
                     lprob_kwargs=None, track_scale=False, save_fits=True):
        """Generator twin of `fit_predict` (bruteforce.py:505-631): yields (pdf, (lmap, levid)).  With
        save_fits=False the objects are processed in chunks, so a streaming caller never holds more than STREAM_BYTES
        of PDFs."""
        Ndata = len(data)
        chunk = Ndata
        if not save_fits and Ndata > 0:
            ng = label_dict.Ngrid if label_dict is not None else (len(label_grid) if label_grid is not None else 1)
            chunk = max(16, int(STREAM_BYTES // (8 * max(1, ng))))
        for o0 in range(0, max(Ndata, 1), max(chunk, 1)):
            o1 = min(Ndata, o0 + chunk)
            lk = lprob_kwargs
            if lk is not None and lk.get("lnprior_bin", None) is not None and (o0 > 0 or o1 < Ndata):
                lk = dict(lk, lnprior_bin=np.asarray(lk["lnprior_bin"])[o0:o1])
            pdfs, lmap, levid = self._fit_predict_all(data[o0:o1], data_err[o0:o1], data_mask[o0:o1], model_labels,
                                                      model_label_errs, lprob_func, label_dict, label_grid, kde_args,
                                                      kde_kwargs, lprob_args, lk, track_scale, save_fits)
            for i in range(len(pdfs)):
                yield pdfs[i], (lmap[i], levid[i])
        self.NDATA = Ndata
