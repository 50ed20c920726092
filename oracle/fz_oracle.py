"""CPU oracle for the frankenz brute-force photometric likelihood path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy (float64) restatement of the
reference algorithm.  It is imported by `tests/`, by `__graft_entry__.smoke()`
and by `bench.py`'s cpu_baseline / `--impl reference` legs, and by nothing in
the product package `frankenz_b200/` (which fails loudly without its CUDA
library).

Parity status: PINNED.  `tests/golden/make_golden.py` runs the unmodified
reference (`/root/reference/frankenz`, v0.3.5, with the numpy/scipy/pandas
installed in the build container) and stores its outputs in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function here
against those vectors (bit-exact for everything except sums whose order the
reference leaves to Python's builtin `sum`).

Third-party arithmetic the reference leans on (unpinned in its setup.py:41-42):
scipy.special.xlogy / gammaln / logsumexp, scipy.spatial.cKDTree (replaced here
by an exact brute-force search: float64 distances on float32-rounded training
features, ascending distance, lowest index first on exact ties),
pandas.unique (order of first appearance), numpy.random.RandomState.normal.

Every function cites the reference lines it restates (paths relative to
/root/reference/).  The arithmetic order inside each expression follows the
reference so that float64 results agree to the last bit wherever numpy's own
reduction order allows.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import gammaln, logsumexp, xlogy

LN2 = math.log(2.0)
LN2PI = math.log(2.0 * math.pi)

# ----------------------------------------------------------------------------
# likelihoods  (frankenz/pdf.py:27-411)
# ----------------------------------------------------------------------------


def clean_inplace(x, xe, xm):
    """In-place input cleaning of ONE object (frankenz/pdf.py:310-311).

    Entries that are non-finite or have a non-positive error get
    flux=0, error=1, mask=0.  The reference mutates the caller's rows.
    """
    bad = ~(np.isfinite(x) & np.isfinite(xe) & (xe > 0.0))
    x[bad] = 0.0
    xe[bad] = 1.0
    xm[bad] = False
    return x, xe, xm


def _variance(xe, me, ignore_model_err):
    # frankenz/pdf.py:76-79 and :171-174
    if ignore_model_err:
        return np.square(xe) + np.zeros_like(me)
    return np.square(xe) + np.square(me)


def _chi2_logpdf(chi2, dof_half):
    # ln of the chi2 density with 2*dof_half degrees of freedom
    # frankenz/pdf.py:92-93 and :228-229
    return xlogy(dof_half - 1.0, chi2) - (chi2 / 2.0) - gammaln(dof_half) - (LN2 * dof_half)


def _mvn_logpdf(chi2, ndim, var):
    # frankenz/pdf.py:96-98, :192-194, :214-216 (sum of ln var is NOT masked)
    out = -0.5 * chi2
    out += -0.5 * (ndim * LN2PI + np.sum(np.log(var), axis=1))
    return out


def loglike_fixed(x, xe, xm, m, me, mm, ignore_model_err=False, dim_prior=True):
    """Fixed-scale ln-likelihood of one object against all models.

    Restates frankenz/pdf.py:27-100 (`_loglike`).  Returns (lnl, Ndim, chi2).
    """
    var = _variance(xe, me, ignore_model_err)
    msk = xm * mm
    ndim = np.sum(msk, axis=1)
    r = x - m
    chi2 = np.sum(msk * np.square(r) / var, axis=1)
    if dim_prior:
        lnl = _chi2_logpdf(chi2, 0.5 * ndim)
    else:
        lnl = _mvn_logpdf(chi2, ndim, var)
    return lnl, ndim, chi2


def loglike_scaled(x, xe, xm, m, me, mm, ignore_model_err=False, dim_prior=True,
                   ltol=1e-3, return_scale=False, return_niter=False):
    """Free-scale ln-likelihood of one object against all models.

    Restates frankenz/pdf.py:103-235 (`_loglike_s`), including the do-while
    refinement whose stopping rule is max|dlnl| over ALL models of the object
    (:199-223) and the `is not True` test on ignore_model_err (:197).
    """
    var = _variance(xe, me, ignore_model_err)
    msk = xm * mm
    ndim = np.sum(msk, axis=1)

    num_i = msk * m * x[None, :]
    inter = np.sum(num_i / var, axis=1)
    num_s = msk * np.square(m)
    shape = np.sum(num_s / var, axis=1)
    scale = inter / shape

    r = x - scale[:, None] * m
    chi2 = np.sum(msk * np.square(r) / var, axis=1)
    lnl = _mvn_logpdf(chi2, ndim, var)

    niter = 0
    if ignore_model_err is not True:
        worst = np.inf
        while worst > ltol:
            var = np.square(xe) + np.square(scale[:, None] * me)
            inter = np.sum(num_i / var, axis=1)
            shape = np.sum(num_s / var, axis=1)
            scale_next = inter / shape
            r = x - scale_next[:, None] * m
            chi2 = np.sum(msk * np.square(r) / var, axis=1)
            lnl_next = _mvn_logpdf(chi2, ndim, var)
            worst = max(abs(lnl_next - lnl))  # builtin max: NaN semantics as reference
            lnl, scale = lnl_next, scale_next
            niter += 1

    if dim_prior:
        lnl = _chi2_logpdf(chi2, 0.5 * (ndim - 1))

    out = [lnl, ndim, chi2]
    if return_scale:
        out += [scale, np.sqrt(1.0 / shape)]
    if return_niter:
        out += [niter]
    return tuple(out)


def loglike(x, xe, xm, m, me, mm, free_scale=False, ignore_model_err=False,
            dim_prior=True, ltol=1e-4, return_scale=False):
    """frankenz/pdf.py:238-323 (`loglike`): clean in place, then dispatch."""
    clean_inplace(x, xe, xm)
    if free_scale:
        return loglike_scaled(x, xe, xm, m, me, mm, ignore_model_err=ignore_model_err,
                              dim_prior=dim_prior, ltol=ltol, return_scale=return_scale)
    return loglike_fixed(x, xe, xm, m, me, mm, ignore_model_err=ignore_model_err,
                         dim_prior=dim_prior)


def logprob(x, xe, xm, m, me, mm, lnprior=None, **kw):
    """frankenz/pdf.py:326-411 (`logprob`).

    `lnprior` (per-model, optional) is this build's replacement for a custom
    Python `lprob_func` (north_star): lnprob = lnlike + lnprior.  With
    lnprior=None the reference's zeros / same-buffer behaviour results.
    """
    res = loglike(x, xe, xm, m, me, mm, **kw)
    lnl = res[0]
    if lnprior is None:
        lp, lpost = np.zeros_like(lnl), lnl[:]
    else:
        lp = np.asarray(lnprior, dtype=float)
        lpost = lnl + lp
    return (lp, lnl, lpost) + tuple(res[1:])


# ----------------------------------------------------------------------------
# kernel density estimation  (frankenz/pdf.py:414-425, 444-622, 778-852)
# ----------------------------------------------------------------------------


def gaussian(mu, std, x):
    """frankenz/pdf.py:414-425."""
    d = x - mu
    nrm = np.sqrt(2.0 * np.pi) * std
    return np.exp(-0.5 * np.square(d / std)) / nrm


class KernelDict:
    """Pre-tabulated truncated Gaussian kernels on an even grid.

    Restates `PDFDict` (frankenz/pdf.py:778-852) including the slice that wraps
    when a kernel is wider than half the grid (:814-818).
    """

    def __init__(self, pdf_grid, sigma_grid, sigma_trunc=5.0):
        self.grid = np.array(pdf_grid)
        self.Ngrid = len(pdf_grid)
        self.min, self.max = min(pdf_grid), max(pdf_grid)
        self.delta = pdf_grid[1] - pdf_grid[0]
        self.sigma_grid = np.array(sigma_grid)
        self.Ndict = len(sigma_grid)
        self.dsigma = sigma_grid[1] - sigma_grid[0]
        self.sigma_trunc = sigma_trunc
        self.sigma_width = np.array(np.ceil(sigma_grid * sigma_trunc / self.delta), dtype="int")
        mid = int(self.Ngrid / 2)
        self.sigma_dict = []
        for s, w in zip(self.sigma_grid, self.sigma_width):
            self.sigma_dict.append(gaussian(self.grid[mid], s, self.grid[mid - w:mid + w + 1]))
        self.sigma_dict_cdf = [np.cumsum(k) for k in self.sigma_dict]

    def fit(self, X, Xe):
        """frankenz/pdf.py:821-852: quantise centres (not clipped) and widths (clipped)."""
        xi = ((X - self.grid[0]) / self.delta).round().astype("int")
        si = np.array(np.round((Xe - self.sigma_grid[0]) / self.dsigma), dtype="int")
        si[si >= self.Ndict] = self.Ndict - 1
        si[si < 0] = 0
        return xi, si


def _select(wt, wt_thresh, cdf_thresh):
    # frankenz/pdf.py:508-516 and :589-597
    n = len(wt)
    if wt_thresh is None and cdf_thresh is None:
        wt_thresh = -np.inf
    if wt_thresh is not None:
        return np.arange(n)[wt > (wt_thresh * np.max(wt))]
    order = np.argsort(wt)
    cdf = np.cumsum(wt[order])
    cdf /= cdf[-1]
    return order[cdf <= (1.0 - cdf_thresh)]


def kde_dict(kd, y_idx, y_std_idx, y_wt=None, wt_thresh=1e-3, cdf_thresh=2e-4):
    """Dictionary KDE (frankenz/pdf.py:529-622, `gauss_kde_dict`)."""
    ng = kd.Ngrid
    pdf = np.zeros(ng)
    if y_wt is None:
        y_wt = np.ones(len(y_idx))
    for i in _select(y_wt, wt_thresh, cdf_thresh):
        s, pos = y_std_idx[i], y_idx[i]
        kern, w, cdf = kd.sigma_dict[s], kd.sigma_width[s], kd.sigma_dict_cdf[s]
        lo, hi = max(pos - w, 0), min(pos + w + 1, ng)
        lpad, hpad = lo - (pos - w), hi - (pos + w + 1)
        nrm = cdf[hpad - 1] if lpad == 0 else cdf[hpad - 1] - cdf[lpad - 1]
        pdf[lo:hi] += (y_wt[i] / nrm) * kern[lpad:2 * w + 1 + hpad]
    return pdf


def kde_grid(y, y_std, x, dx=None, y_wt=None, sig_thresh=5.0, wt_thresh=1e-3,
             cdf_thresh=2e-4):
    """Exact-Gaussian KDE (frankenz/pdf.py:444-526, `gauss_kde`).

    Centres truncate toward zero, windows are upper-exclusive, and each kernel
    is normalised by Python's builtin sum over its clipped window.
    """
    nx = len(x)
    if dx is None:
        dx = x[1] - x[0]
    if y_wt is None:
        y_wt = np.ones(len(y))
    c = np.array((y - x[0]) / dx, dtype="int")
    o = np.array(sig_thresh * y_std / dx, dtype="int")
    up, lo = c + o, c - o
    up[up > nx], lo[lo < 0] = nx, 0
    pdf = np.zeros(nx)
    for i in _select(y_wt, wt_thresh, cdf_thresh):
        g = gaussian(y[i], y_std[i], x[lo[i]:up[i]])
        nrm = sum(g)
        if nrm != 0.0:
            pdf[lo[i]:up[i]] += y_wt[i] / nrm * g
    return pdf


# ----------------------------------------------------------------------------
# feature maps  (frankenz/pdf.py:625-657, 695-734; knn.py:121-130)
# ----------------------------------------------------------------------------


def magnitude(phot, err, zeropoints=1.0):
    mag = -2.5 * np.log10(phot / zeropoints)
    mag_err = 2.5 / np.log(10.0) * err / phot
    return mag, mag_err


def luptitude(phot, err, skynoise=1.0, zeropoints=1.0):
    mag = -2.5 / np.log(10.0) * (np.arcsinh(phot / (2.0 * skynoise)) + np.log(skynoise / zeropoints))
    mag_err = np.sqrt(np.square(2.5 * np.log10(np.e) * err) / (np.square(2.0 * skynoise) + np.square(phot)))
    return mag, mag_err


def identity(phot, err):
    return phot, err


FEATURE_MAPS = {"identity": identity, "magnitude": magnitude, "luptitude": luptitude}


# ----------------------------------------------------------------------------
# BruteForce  (frankenz/bruteforce.py:30-631)
# ----------------------------------------------------------------------------


def bruteforce_fit(models, models_err, models_mask, data, data_err, data_mask,
                   lnprior=None, track_scale=False, **lprob_kwargs):
    """All objects x all models (frankenz/bruteforce.py:127-205, `_fit`).

    Returns a dict of the seven (Ndata, Nmodel) arrays the estimator stores
    (:182-189).  Scale arrays are only filled when `track_scale` (:200-202).
    """
    nd, nm = len(data), len(models)
    out = dict(lnprior=np.zeros((nd, nm)), lnlike=np.zeros((nd, nm)), lnprob=np.zeros((nd, nm)),
               Ndim=np.zeros((nd, nm), dtype="int"), chi2=np.zeros((nd, nm)),
               scale=np.ones((nd, nm)), scale_err=np.zeros((nd, nm)))
    for i in range(nd):
        res = logprob(data[i], data_err[i], data_mask[i], models, models_err, models_mask,
                      lnprior=lnprior, **lprob_kwargs)
        out["lnprior"][i], out["lnlike"][i], out["lnprob"][i] = res[0], res[1], res[2]
        out["Ndim"][i], out["chi2"][i] = res[3], res[4]
        if track_scale:
            out["scale"][i], out["scale_err"][i] = res[5], res[6]
    return out


def weights_to_pdf(lwt, labels, label_errs, y_idx=None, y_std_idx=None, label_dict=None,
                   label_grid=None, **kde_kwargs):
    """One object's PDF from its log-weights (frankenz/bruteforce.py:358-372).

    lmap uses Python's builtin max (:359), the PDF is divided by its own sum (:370).
    """
    lmap, levid = max(lwt), logsumexp(lwt)
    wt = np.exp(lwt - levid)
    if label_dict is not None:
        pdf = kde_dict(label_dict, y_idx, y_std_idx, y_wt=wt, **kde_kwargs)
    else:
        pdf = kde_grid(labels, label_errs, label_grid, y_wt=wt, **kde_kwargs)
    pdf /= pdf.sum()
    return pdf, lmap, levid


def bruteforce_predict(logwt, labels, label_errs, label_dict=None, label_grid=None, **kde_kwargs):
    """frankenz/bruteforce.py:303-372 (`_predict`) over a (Ndata, Nmodel) log-weight array."""
    if label_dict is None and label_grid is None:
        raise ValueError("`label_dict` or `label_grid` must be specified.")
    y_idx = y_std_idx = None
    if label_dict is not None:
        y_idx, y_std_idx = label_dict.fit(labels, label_errs)
    nx = label_dict.Ngrid if label_dict is not None else len(label_grid)
    nd = len(logwt)
    pdfs, lmap, levid = np.zeros((nd, nx)), np.zeros(nd), np.zeros(nd)
    for i in range(nd):
        pdfs[i], lmap[i], levid[i] = weights_to_pdf(logwt[i], labels, label_errs, y_idx, y_std_idx,
                                                    label_dict, label_grid, **kde_kwargs)
    return pdfs, lmap, levid


def bruteforce_fit_predict(models, models_err, models_mask, data, data_err, data_mask,
                           labels, label_errs, label_dict=None, label_grid=None, lnprior=None,
                           kde_kwargs=None, return_best=False, **lprob_kwargs):
    """frankenz/bruteforce.py:505-631 (`_fit_predict`, save_fits=False form)."""
    if label_dict is None and label_grid is None:
        raise ValueError("`label_dict` or `label_grid` must be specified.")
    kde_kwargs = kde_kwargs or {}
    y_idx = y_std_idx = None
    if label_dict is not None:
        y_idx, y_std_idx = label_dict.fit(labels, label_errs)
    nx = label_dict.Ngrid if label_dict is not None else len(label_grid)
    nd = len(data)
    pdfs, lmap, levid = np.zeros((nd, nx)), np.zeros(nd), np.zeros(nd)
    best = np.zeros(nd, dtype="int")
    for i in range(nd):
        res = logprob(data[i], data_err[i], data_mask[i], models, models_err, models_mask,
                      lnprior=lnprior, **lprob_kwargs)
        pdfs[i], lmap[i], levid[i] = weights_to_pdf(res[2], labels, label_errs, y_idx, y_std_idx,
                                                    label_dict, label_grid, **kde_kwargs)
        if return_best:
            best[i] = int(np.argmax(res[2])) if np.any(np.isfinite(res[2])) else 0
    if return_best:
        return pdfs, lmap, levid, best
    return pdfs, lmap, levid


# ----------------------------------------------------------------------------
# NearestNeighbors  (frankenz/knn.py:33-874)
# ----------------------------------------------------------------------------


def knn_train_features(models, models_err, K, feature_map="luptitude", fmap_args=(), fmap_kwargs=None,
                       rstate=None):
    """K Monte-Carlo realisations -> feature map -> float32 (frankenz/knn.py:158-188).

    Draw order follows the reference: one (Nmodel, Nfilt) normal block per tree,
    float64 draws cast to float32 BEFORE the feature map, result cast to float32.
    Masks are not consulted.  Returns float32 array (K, Nmodel, Nfilt).
    """
    fmap = FEATURE_MAPS[feature_map] if isinstance(feature_map, str) else feature_map
    fmap_kwargs = fmap_kwargs or {}
    rstate = np.random if rstate is None else rstate
    feats = []
    for _ in range(K):
        mt = np.array(rstate.normal(models, models_err), dtype="float32")
        yt, _ = np.array(fmap(mt, models_err, *fmap_args, **fmap_kwargs), dtype="float32")
        feats.append(yt)
    return np.stack(feats)


def knn_query_exact(feats, y, k, p=2):
    """Exact k nearest rows of each tree (replaces cKDTree.query with eps=0).

    feats: float32 (K, Nm, Nf) promoted to float64 (cKDTree stores doubles);
    y: float64 (Nf,).  Distances: Minkowski-p in float64; ascending; exact ties
    resolved by lowest index (cKDTree leaves tie order unspecified).
    Returns (idx[K, k] int64, dist[K, k]).  frankenz/knn.py:362-365.
    """
    K = len(feats)
    idx = np.zeros((K, k), dtype="int64")
    dist = np.zeros((K, k))
    for t in range(K):
        d = np.asarray(feats[t], dtype=np.float64) - y[None, :]
        if p == 2:
            dd = np.sum(d * d, axis=1)
        elif p == 1:
            dd = np.sum(np.abs(d), axis=1)
        elif np.isinf(p):
            dd = np.max(np.abs(d), axis=1)
        else:
            dd = np.sum(np.abs(d) ** p, axis=1)
        order = np.argsort(dd, kind="stable")[:k]
        idx[t] = order
        if p == 2:
            dist[t] = np.sqrt(dd[order])
        elif p == 1 or np.isinf(p):
            dist[t] = dd[order]
        else:
            dist[t] = dd[order] ** (1.0 / p)
    return idx, dist


def ordered_unique(indices):
    """Order-of-first-appearance de-duplication (pandas.unique; frankenz/knn.py:368)."""
    seen, out = set(), []
    for v in np.asarray(indices).ravel().tolist():
        if v not in seen:
            seen.add(v)
            out.append(v)
    return np.array(out, dtype="int64")


def knn_fit(models, models_err, models_mask, feats, data, data_err, data_mask, k=20, p=2,
            feature_map="luptitude", fmap_args=(), fmap_kwargs=None, rstate=None, lnprior=None,
            track_scale=False, **lprob_kwargs):
    """frankenz/knn.py:281-388 (`_fit`) with exact neighbours.

    One MC draw per object (in object order) from `rstate`, feature map on the
    draw, K exact queries, ordered union, likelihood of the UNPERTURBED object
    against the ORIGINAL models of the union; padded (Ndata, K*k) outputs
    (neighbors -99, lnX -inf, chi2 +inf, scale 1, scale_err 0; :342-352).
    """
    fmap = FEATURE_MAPS[feature_map] if isinstance(feature_map, str) else feature_map
    fmap_kwargs = fmap_kwargs or {}
    rstate = np.random if rstate is None else rstate
    K = len(feats)
    nd, width = len(data), K * k
    out = dict(Nneighbors=np.zeros(nd, dtype="int"), neighbors=np.zeros((nd, width), dtype="int") - 99,
               lnprior=np.zeros((nd, width)) - np.inf, lnlike=np.zeros((nd, width)) - np.inf,
               lnprob=np.zeros((nd, width)) - np.inf, Ndim=np.zeros((nd, width), dtype="int"),
               chi2=np.zeros((nd, width)) + np.inf, scale=np.ones((nd, width)),
               scale_err=np.zeros((nd, width)))
    for i in range(nd):
        xt = rstate.normal(data[i], data_err[i])
        yt, _ = fmap(xt, data_err[i], *fmap_args, **fmap_kwargs)
        idx, _ = knn_query_exact(feats, np.asarray(yt, dtype=float), k, p)
        u = ordered_unique(idx)
        n = len(u)
        out["Nneighbors"][i] = n
        out["neighbors"][i, :n] = u
        res = logprob(data[i], data_err[i], data_mask[i], models[u], models_err[u], models_mask[u],
                      lnprior=None if lnprior is None else np.asarray(lnprior)[u], **lprob_kwargs)
        out["lnprior"][i, :n], out["lnlike"][i, :n], out["lnprob"][i, :n] = res[0], res[1], res[2]
        out["Ndim"][i, :n], out["chi2"][i, :n] = res[3], res[4]
        if track_scale:
            out["scale"][i, :n], out["scale_err"][i, :n] = res[5], res[6]
    return out


def knn_predict(fit, labels, label_errs, label_dict=None, label_grid=None, logwt=None, **kde_kwargs):
    """frankenz/knn.py:486-558 (`_predict`): KDE over each object's neighbour union."""
    if label_dict is None and label_grid is None:
        raise ValueError("`label_dict` or `label_grid` must be specified.")
    logwt = fit["lnprob"] if logwt is None else logwt
    y_idx = y_std_idx = None
    if label_dict is not None:
        y_idx, y_std_idx = label_dict.fit(labels, label_errs)
    nx = label_dict.Ngrid if label_dict is not None else len(label_grid)
    nd = len(logwt)
    pdfs, lmap, levid = np.zeros((nd, nx)), np.zeros(nd), np.zeros(nd)
    for i in range(nd):
        n = fit["Nneighbors"][i]
        u = fit["neighbors"][i, :n]
        pdfs[i], lmap[i], levid[i] = weights_to_pdf(
            logwt[i][:n], labels[u], label_errs[u],
            None if y_idx is None else y_idx[u], None if y_std_idx is None else y_std_idx[u],
            label_dict, label_grid, **kde_kwargs)
    return pdfs, lmap, levid


# ---- PDF summaries ------------------------------------------------------

def pdfs_resample(pdfs, old_grid, new_grid, renormalize=True, left=0., right=0.):
    """frankenz/pdf.py:855-896: linear interpolation of every PDF onto a new grid, then rows summed to one."""
    out = np.array([np.interp(new_grid, old_grid, row, left=left, right=right) for row in pdfs])
    if renormalize:
        out /= out.sum(axis=1)[:, None]
    return out


def loss_kernel(pgrid, pkern="lorentz", pkern_grid=None):
    """frankenz/pdf.py:1003-1023: the (truth x guess) loss kernel of the `best` estimator.  The default
    argument is (truth - guess) / ((1 + truth) * 0.15), a photo-z convention."""
    if pkern_grid is None:
        truth = pgrid.reshape(-1, 1)
        guess = pgrid.reshape(1, -1)
        pkern_grid = (truth - guess) / ((1. + truth) * 0.15)
    if pkern == "tophat":
        return (np.square(pkern_grid) < 1.)
    if pkern == "gaussian":
        return np.exp(-0.5 * np.square(pkern_grid))
    if pkern == "lorentz":
        return 1. / (1. + np.square(pkern_grid))
    try:
        return pkern(pkern_grid)
    except Exception:
        raise RuntimeError("The input kernel does not appear to be valid.")


def pdfs_summarize(pdfs, pgrid, renormalize=True, rstate=None, pkern="lorentz", pkern_grid=None, wconf_func=None):
    """frankenz/pdf.py:899-1074.  Point estimators (mean / median / mode / minimum-risk `best`), for each of them
    the standard deviation about it, the probability within +-wconf_func(point) and the risk at it; the 2.5 / 16 /
    84 / 97.5 % quantiles; one Monte-Carlo draw per object (inverse CDF at rstate.rand(), drawn in object order,
    pdf.py:995).  `renormalize` divides `pdfs` IN PLACE by its row sums (pdf.py:980)."""
    if rstate is None:
        rstate = np.random
    nobj, ng = len(pdfs), len(pgrid)
    if renormalize:
        pdfs /= pdfs.sum(axis=1)[:, None]
    mean = np.dot(pdfs, pgrid)                                  # pdf.py:983
    mode = pgrid[np.argmax(pdfs, axis=1)]                       # pdf.py:986
    cdfs = pdfs.cumsum(axis=1)                                  # pdf.py:989
    quant = np.zeros((nobj, 6))
    for i in range(nobj):                                       # pdf.py:994-997
        quant[i] = np.interp([0.025, 0.16, 0.5, 0.84, 0.975, rstate.rand()], cdfs[i], pgrid)
    med = quant[:, 2].copy()
    risk = np.dot(pdfs, 1.0 - loss_kernel(pgrid, pkern, pkern_grid))   # pdf.py:1024
    best = pgrid[np.argmin(risk, axis=1)]                              # pdf.py:1025
    grid = pgrid.reshape(1, ng)
    points = (mean, med, mode, best)
    stds = [np.sqrt(np.sum(np.square(grid - p.reshape(nobj, 1)) * pdfs, axis=1)) for p in points]   # pdf.py:1028-1036
    if wconf_func is None:
        def wconf_func(point):
            return (1. + point) * 0.03
    conf = np.zeros((4, nobj))
    rsk = np.zeros((4, nobj))
    for i in range(nobj):
        lo_hi = []
        for p in points:                                        # pdf.py:1044-1062
            w = wconf_func(p[i])
            lo_hi += [p[i] - w, p[i] + w]
        c = np.interp(np.array(lo_hi), pgrid, cdfs[i])
        conf[:, i] = c[1::2] - c[0::2]
        rsk[:, i] = np.interp([p[i] for p in points], pgrid, risk[i])   # pdf.py:1066-1068
    est = tuple((points[k], stds[k], conf[k], rsk[k]) for k in range(4))
    return est + ((quant[:, 0].copy(), quant[:, 1].copy(), quant[:, 3].copy(), quant[:, 4].copy()), quant[:, 5].copy())


# ----------------------------------------------------------------------------
# population likelihood  (frankenz/samplers.py:24-76)
# ----------------------------------------------------------------------------
def loglike_nz(nz, pdfs, overlap=None, return_overlap=False, pair=None, pair_step=None):
    """ln-likelihood of a population N(z) given per-object PDFs (samplers.py:62-76)."""
    perturb = 0.0
    if np.any(~np.isfinite(nz) | (nz < 0.0)):
        lnlike, overlap = -np.inf, np.zeros(len(pdfs))
    else:
        if overlap is None:
            overlap = np.dot(pdfs, nz)
        if pair is not None and pair_step is not None:
            perturb = pair_step * (pdfs[:, pair[0]] - pdfs[:, pair[1]])
        lnlike = np.sum(np.log(overlap + perturb))
    if return_overlap:
        return lnlike, overlap + perturb
    return lnlike


# ----------------------------------------------------------------------------
# SOM / GNG node-fit stages  (frankenz/networks.py:246-356, 782-936)
# ----------------------------------------------------------------------------
def _node_select(node_lnprob, wt_thresh, cdf_thresh):
    """networks.py:322-331 / 888-897: indices of the nodes an object maps to, in the reference's order."""
    if wt_thresh is None and cdf_thresh is None:
        wt_thresh = -np.inf
    if wt_thresh is not None:
        lwt_min = np.log(wt_thresh) + np.max(node_lnprob)
        return np.arange(len(node_lnprob))[node_lnprob > lwt_min]
    idx_sort = np.argsort(node_lnprob)
    node_prob = np.exp(node_lnprob - logsumexp(node_lnprob))
    node_cdf = np.cumsum(node_prob[idx_sort])
    return idx_sort[node_cdf <= (1.0 - cdf_thresh)]


def network_populate(models, models_err, models_mask, nodes, wt_thresh=1e-3, cdf_thresh=2e-4, lpnet_kwargs=None):
    """networks.py:246-356 with track_scale=True: per-node lists of models, ln-weights, scales and BMU lists."""
    if lpnet_kwargs is None:
        lpnet_kwargs = dict(free_scale=True, ignore_model_err=True, return_scale=True)
    nnode = len(nodes)
    out = dict(nodes_idxs=[[] for _ in range(nnode)], nodes_logwts=[[] for _ in range(nnode)],
               nodes_bmus=[[] for _ in range(nnode)], nodes_scales=[[] for _ in range(nnode)],
               nodes_scales_err=[[] for _ in range(nnode)], nodes_Nmatch=np.zeros(nnode, dtype=int),
               models_lmap=np.zeros(len(models)), models_levid=np.zeros(len(models)))
    ye, ym = np.zeros_like(nodes), np.ones_like(nodes)
    for i in range(len(models)):
        x, xe, xm = models[i].copy(), models_err[i].copy(), np.array(models_mask[i], dtype=float)
        res = logprob(x, xe, xm, nodes, ye, ym, **lpnet_kwargs)
        lp = res[2]
        out["nodes_bmus"][np.argmax(lp)].append(i)
        n_idxs = _node_select(lp, wt_thresh, cdf_thresh)
        n_lp = lp[n_idxs]
        lmap, levid = np.max(n_lp), logsumexp(n_lp)
        out["models_lmap"][i], out["models_levid"][i] = lmap, levid
        for j, lw, s, se in zip(n_idxs, n_lp - levid, res[5][n_idxs], res[6][n_idxs]):
            out["nodes_idxs"][j].append(i)
            out["nodes_logwts"][j].append(lw)
            out["nodes_scales"][j].append(s)
            out["nodes_scales_err"][j].append(se)
            out["nodes_Nmatch"][j] += 1
    return out


def network_fit(models, models_err, models_mask, nodes, pop, data, data_err, data_mask, nodes_only=False,
                wt_thresh=1e-3, cdf_thresh=2e-4, lpnet_kwargs=None, lprob_kwargs=None):
    """networks.py:782-936: neighbour lists and fits of every object through the populated network `pop`."""
    if lpnet_kwargs is None:
        lpnet_kwargs = dict(free_scale=True, ignore_model_err=True, return_scale=True)
    match_sel = np.arange(len(nodes))[pop["nodes_Nmatch"] > 0]
    y = nodes[match_sel]
    ye, ym = np.zeros_like(y), np.ones_like(y)
    neighbors, results = [], []
    for i in range(len(data)):
        x, xe, xm = data[i], data_err[i], data_mask[i]
        clean_inplace(x, xe, xm)
        node_results = logprob(x, xe, xm, y, ye, ym, **lpnet_kwargs)
        wsel = _node_select(node_results[2], wt_thresh, cdf_thresh)
        sel_arr = match_sel[wsel]
        if nodes_only:
            neighbors.append(sel_arr)
            results.append([nr[wsel] for nr in node_results])
            continue
        indices = np.array([idx for sidx in sel_arr for idx in pop["nodes_idxs"][sidx]])
        idxs = ordered_unique(indices)
        neighbors.append(np.array(idxs))
        results.append(logprob(x, xe, xm, models[idxs], models_err[idxs], models_mask[idxs], **(lprob_kwargs or {})))
    return neighbors, results
