"""`NearestNeighbors` (KMCkNN) with the reference's interface (frankenz/knn.py:33-874).

The K KDTrees over Monte-Carlo realisations of the training set are replaced by an exact
brute-force distance + top-k search on the GPU.  Monte-Carlo draws come from the caller's
`numpy.random.RandomState` in the reference's draw order (K blocks of (Nmodel, Nfilt) normals
at construction, then one Nfilt-vector per object at fit time), so that with `eps=0` the
neighbour lists equal the reference's.  `eps > 0` (approximate search in cKDTree) and a finite
`distance_upper_bound` have no brute-force analogue: the search here is always exact.
"""
import sys

import numpy as np

from . import pdf as _pdf
from ._engine import Engine, clean_inplace, make_config
from .bruteforce import _check_args, _check_lprob_func

__all__ = ["NearestNeighbors"]


def _identity(x, xe, *args, **kwargs):
    return x, xe


class NearestNeighbors(object):
    """Bayesian nearest-neighbour fits over K Monte-Carlo realisations of the training set.

    Differences from the reference (frankenz/knn.py:191, :362-365): the neighbour search is ALWAYS exact.  `eps` is
    accepted and stored but ignored, so the reference's default call (`eps=1e-3`, approximate cKDTree search) can
    return slightly different - never worse - neighbours; with `eps=0` the lists are identical to the reference's.
    A finite `distance_upper_bound` raises NotImplementedError."""

    def __init__(self, models, models_err, models_mask, leafsize=50, K=25, feature_map='luptitude',
                 fmap_args=None, fmap_kwargs=None, rstate=None, verbose=True):
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
        self.leafsize = leafsize   # kept for interface compatibility; no tree is built
        self.K = K
        self.KDTrees = None        # no GPU analogue: see `features`
        self.features = None       # float32 (K, Nmodel, Nfilt): what the K trees would index
        self.neighbors = None
        self.Nneighbors = None
        self.k = None
        self.eps = None
        self.lp_norm = None
        self.p = None
        self.dbound = None

        self.fmap_args = [] if fmap_args is None else fmap_args
        self.fmap_kwargs = dict() if fmap_kwargs is None else fmap_kwargs
        if feature_map == 'identity':
            feature_map = _identity
        elif feature_map == 'magnitude':
            feature_map = _pdf.magnitude
        elif feature_map == 'luptitude':
            feature_map = _pdf.luptitude
        else:
            try:
                feature_map(np.atleast_2d(models[0]), np.atleast_2d(models_err[0]), *self.fmap_args,
                            **self.fmap_kwargs)
            except Exception:
                raise ValueError("The provided feature map is not valid.")
        self.feature_map = feature_map

        if rstate is None:
            rstate = np.random
        self._engine = Engine(models, models_err, models_mask)
        feats = np.empty((K,) + tuple(models.shape), dtype=np.float32)
        for i in range(K):
            # knn.py:177-184: float64 draw -> float32 -> feature map -> float32; masks are not consulted
            mt = np.array(rstate.normal(models, models_err), dtype='float32')
            yt, _ = np.array(self.feature_map(mt, models_err, *self.fmap_args, **self.fmap_kwargs), dtype='float32')
            feats[i] = yt
        self.features = feats
        self._engine.knn_build(feats)
        if verbose:
            sys.stderr.write("\r{0}/{1} KDTrees constructed\n".format(K, K))
            sys.stderr.flush()

    # ---- internals -----------------------------------------------------------------------------
    def _query_features(self, data, data_err, rstate):
        # one Monte-Carlo draw per object, in object order (knn.py:358); RandomState fills a 2-D request
        # in C order, so a single vectorised call consumes the stream exactly like the per-object loop
        x_t = rstate.normal(data, data_err)
        try:
            y_t, _ = self.feature_map(x_t, data_err, *self.fmap_args, **self.fmap_kwargs)
            q = np.array(y_t, dtype=np.float64)
            if q.shape != np.shape(data):
                raise ValueError
        except Exception:
            q = np.empty(np.shape(data), dtype=np.float64)
            for i in range(len(q)):
                q[i] = self.feature_map(x_t[i], data_err[i], *self.fmap_args, **self.fmap_kwargs)[0]
        return q

    def _run_fit(self, data, data_err, data_mask, lprob_func, rstate, lprob_args, lprob_kwargs, track_scale,
                 save_fits, k, lp_norm):
        _check_lprob_func(lprob_func)
        _check_args(lprob_args, "lprob_args")
        lk = dict(lprob_kwargs or {})
        if track_scale and not (lk.get("free_scale", False) and lk.get("return_scale", False)):
            raise IndexError("tuple index out of range: `track_scale` needs lprob_kwargs free_scale=True and "
                             "return_scale=True")
        if rstate is None:
            rstate = np.random
        self._engine.set_lnprior(lk.get("lnprior", None), lk.get("lnprior_bin", None))
        cfg = make_config(lk, None, track_scale=track_scale)
        # the reference cleans each object only inside logprob, i.e. AFTER its Monte-Carlo draw
        q = self._query_features(data, data_err, rstate)
        clean_inplace(data, data_err, data_mask)
        res = self._engine.knn_fit(q, data, data_err, data_mask, k, lp_norm, cfg)
        self.NDATA = len(data)
        if save_fits:
            self.neighbors, self.Nneighbors = res["neighbors"], res["Nneighbors"]
            self.fit_lnprior, self.fit_lnlike, self.fit_lnprob = res["lnprior"], res["lnlike"], res["lnprob"]
            self.fit_Ndim, self.fit_chi2 = res["Ndim"], res["chi2"]
            self.fit_scale, self.fit_scale_err = res["scale"], res["scale_err"]
        return res

    def _remember(self, k, eps, lp_norm, distance_upper_bound):
        if np.isfinite(distance_upper_bound):
            raise NotImplementedError("a finite `distance_upper_bound` is not supported by the brute-force search")
        self.k, self.eps, self.lp_norm, self.p, self.dbound = k, eps, lp_norm, lp_norm, distance_upper_bound

    @staticmethod
    def _rows(res, i, track_scale):
        n = res["Nneighbors"][i]
        out = (res["lnprior"][i, :n], res["lnlike"][i, :n], res["lnprob"][i, :n], res["Ndim"][i, :n],
               res["chi2"][i, :n])
        if track_scale:
            out = out + (res["scale"][i, :n], res["scale_err"][i, :n])
        return res["neighbors"][i, :n], n, out

    # ---- fit -----------------------------------------------------------------------------------
    def fit(self, data, data_err, data_mask, lprob_func=None, rstate=None, k=20, eps=1e-3, lp_norm=2,
            distance_upper_bound=np.inf, lprob_args=None, lprob_kwargs=None, track_scale=False, verbose=True):
        """Neighbour search + fits to the union of neighbours (knn.py:190-279)."""
        self._remember(k, eps, lp_norm, distance_upper_bound)
        self._run_fit(data, data_err, data_mask, lprob_func, rstate, lprob_args, lprob_kwargs, track_scale, True,
                      k, lp_norm)
        if verbose:
            sys.stderr.write('\rFitting object {0}/{1}\n'.format(len(data), len(data)))
            sys.stderr.flush()

    def _fit(self, data, data_err, data_mask, lprob_func=None, rstate=None, lprob_args=None, lprob_kwargs=None,
             track_scale=False, save_fits=True):
        """Generator twin (knn.py:281-388): yields (idxs, Nidx, results) per object."""
        res = self._run_fit(data, data_err, data_mask, lprob_func, rstate, lprob_args, lprob_kwargs, track_scale,
                            save_fits, self.k, self.lp_norm)
        for i in range(len(data)):
            yield self._rows(res, i, track_scale)

    # ---- predict -------------------------------------------------------------------------------
    def _predict_all(self, model_labels, model_label_errs, label_dict, label_grid, logwt, kde_args, kde_kwargs,
                     fit=None):
        _check_args(kde_args, "kde_args")
        if fit is None:
            fit = dict(lnprob=self.fit_lnprob, neighbors=self.neighbors, Nneighbors=self.Nneighbors)
        if logwt is None:
            logwt = fit["lnprob"]
        if label_dict is None and label_grid is None:
            raise ValueError("`label_dict` or `label_grid` must be specified.")
        if logwt is None:
            raise ValueError("Fits have not been computed and weights have not been provided.")
        self._engine.set_kde(model_labels, model_label_errs, label_dict=label_dict, label_grid=label_grid,
                             kde_kwargs=kde_kwargs)
        cfg = make_config(None, kde_kwargs)
        return self._engine.predict_logwt(logwt, cfg, neighbors=fit["neighbors"], nneighbors=fit["Nneighbors"])

    def predict(self, model_labels, model_label_errs, label_dict=None, label_grid=None, logwt=None, kde_args=None,
                kde_kwargs=None, return_gof=False, verbose=True):
        """1-D PDFs over each object's neighbour union (knn.py:390-484)."""
        pdfs, lmap, levid = self._predict_all(model_labels, model_label_errs, label_dict, label_grid, logwt,
                                              kde_args, kde_kwargs)
        if verbose:
            sys.stderr.write('\rGenerating PDF {0}/{1}\n'.format(len(pdfs), len(pdfs)))
            sys.stderr.flush()
        if return_gof:
            return pdfs, (lmap, levid)
        return pdfs

    def _predict(self, model_labels, model_label_errs, label_dict=None, label_grid=None, logwt=None, kde_args=None,
                 kde_kwargs=None):
        """Generator twin (knn.py:486-558)."""
        pdfs, lmap, levid = self._predict_all(model_labels, model_label_errs, label_dict, label_grid, logwt,
                                              kde_args, kde_kwargs)
        for i in range(len(pdfs)):
            yield pdfs[i], (lmap[i], levid[i])

    # ---- fit_predict ---------------------------------------------------------------------------
    def _fit_predict_all(self, data, data_err, data_mask, model_labels, model_label_errs, lprob_func, label_dict,
                         label_grid, kde_args, kde_kwargs, lprob_args, lprob_kwargs, track_scale, save_fits, rstate):
        if label_dict is None and label_grid is None:
            raise ValueError("`label_dict` or `label_grid` must be specified.")
        if not save_fits:
            # nothing is kept (knn.py:806-815 skips every store): search, union, fits and KDE stay on the device
            _check_lprob_func(lprob_func)
            _check_args(lprob_args, "lprob_args")
            _check_args(kde_args, "kde_args")
            lk = dict(lprob_kwargs or {})
            if rstate is None:
                rstate = np.random
            self._engine.set_lnprior(lk.get("lnprior", None), lk.get("lnprior_bin", None))
            self._engine.set_kde(model_labels, model_label_errs, label_dict=label_dict, label_grid=label_grid,
                                 kde_kwargs=kde_kwargs)
            cfg = make_config(lk, kde_kwargs, track_scale=False)
            q = self._query_features(data, data_err, rstate)
            clean_inplace(data, data_err, data_mask)
            pdfs, lmap, levid, nn = self._engine.knn_fit_predict(q, data, data_err, data_mask, self.k, self.lp_norm, cfg)
            self.NDATA = len(data)
            self.Nneighbors_last = nn
            return pdfs, lmap, levid
        res = self._run_fit(data, data_err, data_mask, lprob_func, rstate, lprob_args, lprob_kwargs, track_scale,
                            save_fits, self.k, self.lp_norm)
        return self._predict_all(model_labels, model_label_errs, label_dict, label_grid, None, kde_args, kde_kwargs,
                                 fit=res)

    def fit_predict(self, data, data_err, data_mask, model_labels, model_label_errs, lprob_func=None, rstate=None,
                    k=20, eps=1e-3, lp_norm=2, distance_upper_bound=np.inf, label_dict=None, label_grid=None,
                    kde_args=None, kde_kwargs=None, lprob_args=None, lprob_kwargs=None, return_gof=False,
                    track_scale=False, verbose=True, save_fits=True):
        """Neighbour search, fits and PDFs in one call (knn.py:560-720)."""
        self._remember(k, eps, lp_norm, distance_upper_bound)
        pdfs, lmap, levid = self._fit_predict_all(data, data_err, data_mask, model_labels, model_label_errs,
                                                  lprob_func, label_dict, label_grid, kde_args, kde_kwargs,
                                                  lprob_args, lprob_kwargs, track_scale, save_fits, rstate)
        if verbose:
            sys.stderr.write('\rGenerating PDF {0}/{1}\n'.format(len(pdfs), len(pdfs)))
            sys.stderr.flush()
        if return_gof:
            return pdfs, (lmap, levid)
        return pdfs

    def _fit_predict(self, data, data_err, data_mask, model_labels, model_label_errs, lprob_func=None, rstate=None,
                     label_dict=None, label_grid=None, kde_args=None, kde_kwargs=None, lprob_args=None,
                     lprob_kwargs=None, track_scale=False, save_fits=True):
        """Generator twin (knn.py:722-874)."""
        pdfs, lmap, levid = self._fit_predict_all(data, data_err, data_mask, model_labels, model_label_errs,
                                                  lprob_func, label_dict, label_grid, kde_args, kde_kwargs,
                                                  lprob_args, lprob_kwargs, track_scale, save_fits, rstate)
        for i in range(len(pdfs)):
            yield pdfs[i], (lmap[i], levid[i])
