"""Node-fit stages of the reference's manifold learners on the GPU (frankenz/networks.py:246-356, 782-936).

SURVEY.md section 8f rank 3: a trained SOM / growing-neural-gas network is a small set of `nodes` (Nnode, Nfilt); mapping
the training models onto the nodes (`_Network._populate_network`) and fitting observed objects through the nodes
(`_Network._fit`) both score photometry against the nodes through the `lprob_func` protocol - the same brute-force
likelihood path as `BruteForce`, with the nodes as error-free, unmasked "models" (`ye = 0`, `ym = 1`).  Those two stages
run here on the CUDA kernels of the path; the sequential stochastic TRAINING of the networks
(networks.py:1682-1867, 2037-2260) is out of scope: pass the trained `nodes` in.

    net = NetworkFit(models, models_err, models_mask, nodes)      # nodes from SelfOrganizingMap / GrowingNeuralGas
    net.populate_network()                                         # nodes_idxs, nodes_logwts, nodes_bmus, ...
    net.fit(data, data_err, data_mask)                             # neighbors, Nneighbors, fit_lnprob, ...
    pdfs = net.predict(model_labels, model_label_errs, label_dict=rdict)

Attribute names and contents follow the reference so that code written against `_Network` keeps working.
"""
import sys

import numpy as np

from . import pdf as _pdf
from ._engine import Engine, clean_inplace, make_config
from .bruteforce import _check_args, _check_lprob_func

__all__ = ["NetworkFit"]

_LPNET_DEFAULT = {'free_scale': True, 'ignore_model_err': True, 'return_scale': True}    # networks.py:289-291


def _select(lnprob, wt_thresh, cdf_thresh):
    """Row-wise node selection (networks.py:322-331 / 888-897): boolean mask (N, Nnode) and, for the CDF rule, the
    per-row order in which the reference lists the selected nodes (None = ascending node index)."""
    if wt_thresh is None and cdf_thresh is None:
        wt_thresh = -np.inf
    if wt_thresh is not None:
        with np.errstate(divide="ignore"):
            lwt_min = np.log(wt_thresh) + np.max(lnprob, axis=1)
        return lnprob > lwt_min[:, None], None
    order = np.argsort(lnprob, axis=1)
    mx = np.max(lnprob, axis=1, keepdims=True)
    levid = mx + np.log(np.sum(np.exp(lnprob - mx), axis=1, keepdims=True))
    prob = np.exp(lnprob - levid)
    cdf = np.cumsum(np.take_along_axis(prob, order, axis=1), axis=1)
    keep_sorted = cdf <= (1. - cdf_thresh)
    mask = np.zeros(lnprob.shape, dtype=bool)
    np.put_along_axis(mask, order, keep_sorted, axis=1)
    return mask, order


def _logsumexp_rows(a, mask):
    """logsumexp over the selected entries of every row (scipy's max-shifted form)."""
    am = np.where(mask, a, -np.inf)
    mx = np.max(am, axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.sum(np.where(mask, np.exp(a - mx[:, None]), 0.), axis=1)
        return mx, mx + np.log(s)


class NetworkFit(object):
    """`_Network.populate_network` / `fit` / `predict` for a network whose nodes are already trained."""

    def __init__(self, models, models_err, models_mask, nodes):
        self.models, self.models_err, self.models_mask = models, models_err, models_mask
        self.NMODEL, self.NDIM = models.shape
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.NNODE = len(self.nodes)
        self.models_lmap = np.zeros(self.NMODEL) - np.inf          # networks.py:160-161
        self.models_levid = np.zeros(self.NMODEL) - np.inf
        self.nodes_idxs = self.nodes_logwts = self.nodes_bmus = None
        self.nodes_scales = self.nodes_scales_err = self.nodes_Nmatch = None
        self.lpnet_kwargs = dict(_LPNET_DEFAULT)
        self.NDATA = None
        self.neighbors = self.Nneighbors = None
        self.fit_lnprior = self.fit_lnlike = self.fit_lnprob = self.fit_Ndim = self.fit_chi2 = None
        self.fit_scale = self.fit_scale_err = None
        self._node_engine = None
        self._model_engine = None
        self._node_key = None

    # ---- engines ---------------------------------------------------------------------------------------------------
    def _nodes_eng(self, sel=None):
        """Engine holding the (selected) nodes as error-free, unmasked models (networks.py:307-309, 872-874)."""
        key = None if sel is None else tuple(sel.tolist())
        if self._node_engine is None or key != self._node_key:
            if self._node_engine is not None:
                self._node_engine.close()
            y = self.nodes if sel is None else np.ascontiguousarray(self.nodes[sel])
            self._node_engine = Engine(y, np.zeros_like(y), np.ones_like(y))
            self._node_key = key
        return self._node_engine

    def _models_eng(self):
        if self._model_engine is None:
            self._model_engine = Engine(self.models, self.models_err, self.models_mask)
        return self._model_engine

    @staticmethod
    def _node_fit(eng, x, xe, xm, lpnet_kwargs, chunk_bytes=1 << 30):
        """(N x Nnode) lnprob / scale / scale_err / full result arrays of objects against the nodes, in object chunks."""
        lk = dict(lpnet_kwargs)
        track = bool(lk.pop("return_scale", False) and lk.get("free_scale", False))
        cfg = make_config(lk, None, track_scale=track)
        n, nn = len(x), eng.Nm
        chunk = max(1, int(chunk_bytes // (7 * 8 * nn)))
        parts = []
        for o0 in range(0, n, chunk):
            parts.append(eng.fit(x[o0:o0 + chunk], xe[o0:o0 + chunk], xm[o0:o0 + chunk], cfg))
        return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}, track

    # ---- populate ----------------------------------------------------------------------------------------------------
    def populate_network(self, lpnet_func=None, wt_thresh=1e-3, cdf_thresh=2e-4, lpnet_args=None, lpnet_kwargs=None,
                         track_scale=True, verbose=True):
        """Map the models onto the nodes (networks.py:175-356): for every model the ln-posterior over all nodes, its
        best-matching node, the nodes above the weight threshold and the model's normalised ln-weight on each."""
        _check_lprob_func(lpnet_func)
        _check_args(lpnet_args, "lpnet_args")
        self.lpnet_kwargs = dict(_LPNET_DEFAULT if lpnet_kwargs is None else lpnet_kwargs)
        x, xe, xm = (np.array(a, dtype=np.float64) for a in (self.models, self.models_err, self.models_mask))
        clean_inplace(x, xe, xm)          # logprob cleans every row it is given (pdf.py:310-311); on copies here
        res, tracked = self._node_fit(self._nodes_eng(), x, xe, xm, self.lpnet_kwargs)
        lp = res["lnprob"]
        nnode, nmodel = self.NNODE, self.NMODEL
        mask, _ = _select(lp, wt_thresh, cdf_thresh)
        lmap, levid = _logsumexp_rows(lp, mask)
        self.models_lmap, self.models_levid = lmap, levid
        bmu = np.argmax(lp, axis=1)
        if track_scale and not tracked:
            raise IndexError("tuple index out of range: `track_scale` needs lpnet_kwargs free_scale=True and "
                             "return_scale=True")
        scales = res["scale"] if track_scale else np.ones_like(lp)
        serrs = res["scale_err"] if track_scale else np.zeros_like(lp)
        # per node: the models mapped to it, in model order (the reference appends while it walks the models)
        mi, ni = np.nonzero(mask)                       # row-major: model order within every node after the stable sort
        order = np.argsort(ni, kind="stable")
        mi, ni = mi[order], ni[order]
        cuts = np.searchsorted(ni, np.arange(nnode + 1))
        lw = lp[mi, ni] - levid[mi]
        self.nodes_idxs = [mi[cuts[j]:cuts[j + 1]].tolist() for j in range(nnode)]
        self.nodes_logwts = [lw[cuts[j]:cuts[j + 1]].tolist() for j in range(nnode)]
        self.nodes_scales = [scales[mi[cuts[j]:cuts[j + 1]], j].tolist() for j in range(nnode)]
        self.nodes_scales_err = [serrs[mi[cuts[j]:cuts[j + 1]], j].tolist() for j in range(nnode)]
        self.nodes_Nmatch = np.diff(cuts).astype('int')
        bo = np.argsort(bmu, kind="stable")
        bcuts = np.searchsorted(bmu[bo], np.arange(nnode + 1))
        self.nodes_bmus = [bo[bcuts[j]:bcuts[j + 1]].tolist() for j in range(nnode)]
        if verbose:
            sys.stderr.write('\rMapping objects {0}/{1}\n'.format(nmodel, nmodel))
            sys.stderr.flush()

    # ---- fit -----------------------------------------------------------------------------------------------------------
    def fit(self, data, data_err, data_mask, lprob_func=None, nodes_only=False, wt_thresh=1e-3, cdf_thresh=2e-4,
            lprob_args=None, lprob_kwargs=None, track_scale=False, discrete=False, verbose=True, save_fits=True):
        """Fit objects through the network (networks.py:696-936): objects x matched nodes (first stage, the same
        likelihood kernels), node selection, then either the node results themselves (`nodes_only`) or fits to the union
        of the models mapped to the selected nodes."""
        _check_lprob_func(lprob_func)
        _check_args(lprob_args, "lprob_args")
        if self.nodes_Nmatch is None:
            raise ValueError("Models have not been mapped onto the network: call populate_network first.")
        clean_inplace(data, data_err, data_mask)
        ndata = len(data)
        self.NDATA = ndata
        self.nodes_only = nodes_only
        match_sel = np.arange(self.NNODE)[self.nodes_Nmatch > 0]
        res, tracked = self._node_fit(self._nodes_eng(match_sel), np.asarray(data, dtype=np.float64),
                                      np.asarray(data_err, dtype=np.float64), np.asarray(data_mask, dtype=np.float64),
                                      self.lpnet_kwargs)
        mask, order = _select(res["lnprob"], wt_thresh, cdf_thresh)
        names = ["lnprior", "lnlike", "lnprob", "Ndim", "chi2"] + (["scale", "scale_err"] if tracked else [])
        sel_lists = []
        for i in range(ndata):
            if order is None:
                w = np.nonzero(mask[i])[0]
            else:
                w = order[i][mask[i][order[i]]]
            sel_lists.append(w)
        if nodes_only:
            neighbors = [match_sel[w] for w in sel_lists]
            results = [tuple(res[k][i][w] for k in names) for i, w in enumerate(sel_lists)]
        else:
            src = self.nodes_bmus if discrete else self.nodes_idxs
            neighbors = []
            for w in sel_lists:
                chunks = [src[s] for s in match_sel[w]]
                ind = np.fromiter((v for c in chunks for v in c), dtype=np.int64)
                _, first = np.unique(ind, return_index=True)          # pandas.unique: order of first appearance
                neighbors.append(ind[np.sort(first)])
            nn = np.array([len(v) for v in neighbors], dtype=np.int64)
            width = max(1, int(nn.max()) if ndata else 1)
            nb = np.full((ndata, width), 0, dtype=np.int64)
            for i, v in enumerate(neighbors):
                nb[i, :len(v)] = v
            lk = dict(lprob_kwargs or {})
            track2 = bool(track_scale)
            if track2 and not (lk.get("free_scale", False) and lk.get("return_scale", False)):
                raise IndexError("tuple index out of range: `track_scale` needs lprob_kwargs free_scale=True and "
                                 "return_scale=True")
            eng = self._models_eng()
            eng.set_lnprior(lk.get("lnprior", None), lk.get("lnprior_bin", None))
            cfg = make_config(lk, None, track_scale=track2)
            full = eng.fit_gather(data, data_err, data_mask, nb, nn, cfg)
            names = ["lnprior", "lnlike", "lnprob", "Ndim", "chi2"] + (["scale", "scale_err"] if track2 else [])
            results = [tuple(full[k][i, :nn[i]] for k in names) for i in range(ndata)]
        if save_fits:
            self.Nneighbors = np.array([len(v) for v in neighbors], dtype='int')
            self.neighbors = neighbors
            self.fit_lnprior = [r[0] for r in results]
            self.fit_lnlike = [r[1] for r in results]
            self.fit_lnprob = [r[2] for r in results]
            self.fit_Ndim = [r[3] for r in results]
            self.fit_chi2 = [r[4] for r in results]
            self.fit_scale = [r[5] for r in results] if len(names) > 5 else []
            self.fit_scale_err = [r[6] for r in results] if len(names) > 5 else []
        if verbose:
            sys.stderr.write('\rFitting object {0}/{1}\n'.format(ndata, ndata))
            sys.stderr.flush()
        return neighbors, results

    # ---- predict -------------------------------------------------------------------------------------------------------
    def predict(self, model_labels, model_label_errs, label_dict=None, label_grid=None, logwt=None, kde_args=None,
                kde_kwargs=None, return_gof=False, verbose=True):
        """PDFs from the stored model fits (networks.py:938-1128, the branch that is not `nodes_only`): the KDE over
        each object's neighbour list, on the device."""
        _check_args(kde_args, "kde_args")
        if label_dict is None and label_grid is None:
            raise ValueError("`label_dict` or `label_grid` must be specified.")
        if logwt is None:
            logwt = self.fit_lnprob
        if logwt is None or self.neighbors is None:
            raise ValueError("Fits have not been computed and weights have not been provided.")
        if getattr(self, "nodes_only", False):
            raise NotImplementedError("predict for nodes_only fits needs per-node PDFs; fit with nodes_only=False")
        ndata = len(self.neighbors)
        nn = np.array([len(v) for v in self.neighbors], dtype=np.int64)
        width = max(1, int(nn.max()) if ndata else 1)
        nb = np.zeros((ndata, width), dtype=np.int64)
        lw = np.full((ndata, width), -np.inf)
        for i, (v, l) in enumerate(zip(self.neighbors, logwt)):
            nb[i, :len(v)] = v
            lw[i, :len(v)] = l
        eng = self._models_eng()
        eng.set_kde(model_labels, model_label_errs, label_dict=label_dict, label_grid=label_grid, kde_kwargs=kde_kwargs)
        pdfs, lmap, levid = eng.predict_logwt(lw, make_config(None, kde_kwargs), neighbors=nb, nneighbors=nn)
        if verbose:
            sys.stderr.write('\rGenerating PDF {0}/{1}\n'.format(ndata, ndata))
            sys.stderr.flush()
        if return_gof:
            return pdfs, (lmap, levid)
        return pdfs
