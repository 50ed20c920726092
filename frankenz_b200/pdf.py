"""Likelihood and PDF helpers with the reference's signatures (frankenz/pdf.py:21-24).

`loglike` / `logprob` evaluate on the GPU through the C ABI (one object against all
models).  `PDFDict` tabulates the truncated Gaussian kernels that the CUDA KDE kernels
consume; the magnitude transforms are the feature maps of the kNN estimator.
"""
import numpy as np

from ._engine import Engine, SummaryEngine, clean_inplace, make_config

__all__ = ["_loglike", "_loglike_s", "loglike", "logprob", "gaussian", "gaussian_bin", "gauss_kde", "gauss_kde_dict",
           "magnitude", "inv_magnitude", "luptitude", "inv_luptitude", "PDFDict", "pdfs_resample", "pdfs_summarize"]


def _one_object(data, data_err, data_mask, models, models_err, models_mask, free_scale, ignore_model_err,
                dim_prior, ltol, return_scale, lnprior=None):
    # in-place cleaning of the caller's vectors, as pdf.py:310-311
    clean_inplace(data, data_err, data_mask)
    eng = Engine(models, models_err, models_mask)
    try:
        if lnprior is not None:
            eng.set_lnprior(lnprior)
        cfg = make_config(dict(free_scale=free_scale, ignore_model_err=ignore_model_err, dim_prior=dim_prior,
                               ltol=ltol), track_scale=bool(return_scale and free_scale))
        res = eng.fit(np.atleast_2d(data), np.atleast_2d(data_err), np.atleast_2d(data_mask), cfg)
    finally:
        eng.close()
    return {k: v[0] for k, v in res.items()}


def loglike(data, data_err, data_mask, models, models_err, models_mask, free_scale=False, ignore_model_err=False,
            dim_prior=True, ltol=1e-4, return_scale=False, *args, **kwargs):
    """ln-likelihood of one object against all models (signature of frankenz/pdf.py:238-240).

    Returns (lnlike, Ndim, chi2) or, with `free_scale` and `return_scale`, additionally
    (scale, scale_err).  `Ndim` is returned as float64 like the reference (device value, integral for 0/1 masks).
    """
    r = _one_object(data, data_err, data_mask, models, models_err, models_mask, free_scale, ignore_model_err,
                    dim_prior, ltol, return_scale)
    ndim = r["Ndim"].astype(float)
    if return_scale and free_scale:
        return r["lnlike"], ndim, r["chi2"], r["scale"], r["scale_err"]
    return r["lnlike"], ndim, r["chi2"]


def logprob(data, data_err, data_mask, models, models_err, models_mask, free_scale=False, ignore_model_err=False,
            dim_prior=True, ltol=1e-4, return_scale=False, lnprior=None, *args, **kwargs):
    """Estimator protocol adapter (signature of frankenz/pdf.py:326-328).

    Returns (lnprior, lnlike, lnprob, Ndim, chi2[, scale, scale_err]).  `lnprior` (per model,
    optional) is the built-in replacement for a custom Python `lprob_func`.
    """
    r = _one_object(data, data_err, data_mask, models, models_err, models_mask, free_scale, ignore_model_err,
                    dim_prior, ltol, return_scale, lnprior=lnprior)
    ndim = r["Ndim"].astype(float)
    out = (r["lnprior"], r["lnlike"], r["lnprob"], ndim, r["chi2"])
    if return_scale and free_scale:
        out = out + (r["scale"], r["scale_err"])
    return out


def _loglike(data, data_err, data_mask, models, models_err, models_mask, ignore_model_err=False, dim_prior=True,
             *args, **kwargs):
    """Fixed-scale ln-likelihood (signature of frankenz/pdf.py:27-29) -> (lnlike, Ndim, chi2).  Unlike the reference's
    internal function, non-finite entries are cleaned like `loglike` does (the device kernels always clean)."""
    return loglike(data, data_err, data_mask, models, models_err, models_mask, free_scale=False,
                   ignore_model_err=ignore_model_err, dim_prior=dim_prior)


def _loglike_s(data, data_err, data_mask, models, models_err, models_mask, ignore_model_err=False, dim_prior=True,
               ltol=1e-4, return_scale=False, *args, **kwargs):
    """Free-scale ln-likelihood (signature of frankenz/pdf.py:103-105) -> (lnlike, Ndim, chi2[, scale, scale_err])."""
    return loglike(data, data_err, data_mask, models, models_err, models_mask, free_scale=True,
                   ignore_model_err=ignore_model_err, dim_prior=dim_prior, ltol=ltol, return_scale=return_scale)


def gaussian(mu, std, x):
    """N(x | mu, std) on the grid `x` (frankenz/pdf.py:414-425); used to tabulate `PDFDict`."""
    z = (x - mu) / std
    return np.exp(-0.5 * np.square(z)) / (np.sqrt(2. * np.pi) * std)


def gaussian_bin(mu, std, bins):
    """Gaussian integrated over the bins with edges `bins` (frankenz/pdf.py:428-441; host helper, not on the path)."""
    from scipy.special import erf
    cdf = 0.5 * (1. + erf((bins - mu) / (np.sqrt(2) * std)))
    return cdf[1:] - cdf[:-1]


def _kde_on_device(ny, y_wt, wt_thresh, cdf_thresh, setup):
    """Shared body of gauss_kde / gauss_kde_dict: the device KDE kernels work on the log-weights of one pseudo-object
    (weights exp(logwt - levid), levid = ln sum(y_wt)); the un-normalised stack they return is scaled back by
    exp(levid) = sum(y_wt), which gives the reference's un-normalised PDF."""
    y_wt = np.ones(ny) if y_wt is None else np.asarray(y_wt, dtype=np.float64)
    eng = Engine(np.zeros((ny, 1)), np.zeros((ny, 1)), np.ones((ny, 1)))
    try:
        setup(eng)
        if not np.any(y_wt > 0.) or not np.all(np.isfinite(y_wt)):
            return np.zeros(eng.Ng)         # nothing passes `wt > wt_thresh * max(wt)` (pdf.py:508-516 / :589-597)
        cfg = make_config(None, dict(wt_thresh=wt_thresh, cdf_thresh=cdf_thresh))
        cfg.reserved = 1
        with np.errstate(divide="ignore"):
            pdfs, _, levid = eng.predict_logwt(np.log(y_wt)[None, :], cfg)
    finally:
        eng.close()
    return pdfs[0] * np.exp(levid[0])


def gauss_kde(y, y_std, x, dx=None, y_wt=None, sig_thresh=5., wt_thresh=1e-3, cdf_thresh=2e-4, *args, **kwargs):
    """Exact-Gaussian KDE of weighted labels on the grid `x` (signature of frankenz/pdf.py:444-445), evaluated by the
    device kernel the estimators use (`kde_add_grid`).  Returns the un-normalised PDF like the reference."""
    kk = dict(dx=dx, sig_thresh=sig_thresh)
    return _kde_on_device(len(y), y_wt, wt_thresh, cdf_thresh,
                          lambda eng: eng.set_kde(y, y_std, label_grid=x, kde_kwargs=kk))


def gauss_kde_dict(pdfdict, y=None, y_std=None, y_idx=None, y_std_idx=None, y_wt=None, wt_thresh=1e-3,
                   cdf_thresh=2e-4, *args, **kwargs):
    """Dictionary KDE (signature of frankenz/pdf.py:529-531) through the device kernel `kde_add_dict`; either the
    labels (`y`, `y_std`) or their dictionary indices (`y_idx`, `y_std_idx`) are given.  Un-normalised like the
    reference."""
    if y_idx is None or y_std_idx is None:
        if y is None or y_std is None:
            raise ValueError("At least one pair of (y, y_std) or (y_idx, y_std_idx) must be specified.")
        y_idx, y_std_idx = pdfdict.fit(np.asarray(y, dtype=np.float64), np.asarray(y_std, dtype=np.float64))
    return _kde_on_device(len(y_idx), y_wt, wt_thresh, cdf_thresh,
                          lambda eng: eng.set_kde_dict_idx(pdfdict, y_idx, y_std_idx))


def magnitude(phot, err, zeropoints=1., *args, **kwargs):
    """Flux -> AB magnitude feature map (frankenz/pdf.py:625-657)."""
    return -2.5 * np.log10(phot / zeropoints), 2.5 / np.log(10.) * err / phot


def inv_magnitude(mag, err, zeropoints=1., *args, **kwargs):
    """AB magnitude -> flux (frankenz/pdf.py:660-692)."""
    phot = 10**(-0.4 * mag) * zeropoints
    return phot, err * 0.4 * np.log(10.) * phot


def luptitude(phot, err, skynoise=1., zeropoints=1., *args, **kwargs):
    """Flux -> asinh magnitude feature map (frankenz/pdf.py:695-734)."""
    mag = -2.5 / np.log(10.) * (np.arcsinh(phot / (2. * skynoise)) + np.log(skynoise / zeropoints))
    mag_err = np.sqrt(np.square(2.5 * np.log10(np.e) * err) / (np.square(2. * skynoise) + np.square(phot)))
    return mag, mag_err


def inv_luptitude(mag, err, skynoise=1., zeropoints=1., *args, **kwargs):
    """asinh magnitude -> flux (frankenz/pdf.py:737-775)."""
    phot = (2. * skynoise) * np.sinh(np.log(10.) / -2.5 * mag - np.log(skynoise / zeropoints))
    phot_err = np.sqrt((np.square(2. * skynoise) + np.square(phot)) * np.square(err)) / (2.5 * np.log10(np.e))
    return phot, phot_err


class PDFDict(object):
    """Grid + dictionary of truncated Gaussian kernels (frankenz/pdf.py:778-852).

    Attributes match the reference: Ngrid, min, max, delta, grid, Ndict, sigma_grid, dsigma,
    sigma_width, sigma_dict, sigma_dict_cdf.  The tables are uploaded to the GPU by the
    estimators; kernels wider than half the grid come out truncated exactly as in the reference
    (negative slice start) and are refused when a label maps onto them.
    """

    def __init__(self, pdf_grid, sigma_grid, sigma_trunc=5.):
        self.grid = np.array(pdf_grid)
        self.Ngrid = len(pdf_grid)
        self.min, self.max = min(pdf_grid), max(pdf_grid)
        self.delta = pdf_grid[1] - pdf_grid[0]
        self.sigma_grid = np.array(sigma_grid)
        self.Ndict = len(sigma_grid)
        self.dsigma = sigma_grid[1] - sigma_grid[0]
        self.sigma_trunc = sigma_trunc
        self.sigma_width = np.array(np.ceil(sigma_grid * sigma_trunc / self.delta), dtype='int')
        centre = int(self.Ngrid / 2)
        self.sigma_dict = [gaussian(self.grid[centre], s, self.grid[centre - w:centre + w + 1])
                           for s, w in zip(self.sigma_grid, self.sigma_width)]
        self.sigma_dict_cdf = [np.cumsum(k) for k in self.sigma_dict]

    def fit(self, X, Xe):
        """Nearest grid index of each mean (unclipped) and dictionary index of each width (clipped)."""
        X_idx = ((X - self.grid[0]) / self.delta).round().astype('int')
        Xe_idx = np.array(np.round((Xe - self.sigma_grid[0]) / self.dsigma), dtype='int')
        np.clip(Xe_idx, 0, self.Ndict - 1, out=Xe_idx)
        return X_idx, Xe_idx


def pdfs_resample(pdfs, old_grid, new_grid, renormalize=True, left=0., right=0.):
    """Drop-in for frankenz.pdf.pdfs_resample (pdf.py:855-896): numpy.interp of every PDF onto `new_grid` (host; a
    convenience around the hot path, not part of it)."""
    new_pdfs = np.array([np.interp(new_grid, old_grid, row, left=left, right=right) for row in pdfs])
    if renormalize:
        new_pdfs /= new_pdfs.sum(axis=1)[:, None]
    return new_pdfs


def _loss_kernel(pgrid, pkern, pkern_grid):
    """The (truth x guess) kernel of the `best` estimator (pdf.py:1003-1023), evaluated on the host with numpy like
    the reference so that a user-supplied callable keeps working."""
    ng = len(pgrid)
    if pkern_grid is None:
        truth, guess = pgrid.reshape(ng, 1), pgrid.reshape(1, ng)
        pkern_grid = (truth - guess) / ((1. + truth) * 0.15)
    if pkern == 'tophat':
        return (np.square(pkern_grid) < 1.)
    if pkern == 'gaussian':
        return np.exp(-0.5 * np.square(pkern_grid))
    if pkern == 'lorentz':
        return 1. / (1. + np.square(pkern_grid))
    try:
        return pkern(pkern_grid)
    except Exception:
        raise RuntimeError("The input kernel does not appear to be valid.")


_LOSS_CACHE = []      # [(key, loss)], most recent first


def _loss_matrix(pgrid, pkern, pkern_grid):
    """`1 - kernel` as the device wants it.  For the named kernels on the default (truth - guess) / ((1 + truth) 0.15)
    argument the matrix depends on the grid alone, and a survey is summarised batch after batch on one grid: the last
    few are kept (4 MB each at 701 points) instead of being rebuilt per call (~15 ms of host time).  Callables and
    user-supplied `pkern_grid`s are evaluated every time, like the reference does."""
    cacheable = isinstance(pkern, str) and pkern_grid is None
    if cacheable:
        key = (pkern, pgrid.tobytes())
        for i, (k, v) in enumerate(_LOSS_CACHE):
            if k == key:
                if i:
                    _LOSS_CACHE.insert(0, _LOSS_CACHE.pop(i))
                return v
    loss = np.ascontiguousarray(1.0 - _loss_kernel(pgrid, pkern, pkern_grid), dtype=np.float64)
    if cacheable:
        loss.setflags(write=False)
        _LOSS_CACHE.insert(0, (key, loss))
        del _LOSS_CACHE[4:]
    return loss


def pdfs_summarize(pdfs, pgrid, renormalize=True, rstate=None, pkern='lorentz', pkern_grid=None, wconf_func=None):
    """Drop-in for frankenz.pdf.pdfs_summarize (pdf.py:899-1074): same arguments, same 6-tuple
    ((mean, std, conf, risk), (median, ...), (mode, ...), (best, ...), (low95, low68, high68, high95), mc).

    The per-object work (row sums, CDFs, quantile interpolation, the (Nobj x Ngrid) x (Ngrid x Ngrid) risk product,
    moments) runs in `fzb_pdfs_summarize` / `fzb_pdfs_conf`; the loss kernel, `wconf_func` and the random draws are
    evaluated on the host exactly as the reference does (one `rstate.rand()` per object, in order), so callables keep
    working.  Like the reference, `renormalize=True` divides `pdfs` in place by its row sums."""
    if rstate is None:
        rstate = np.random
    if not (isinstance(pdfs, np.ndarray) and pdfs.dtype == np.float64 and pdfs.flags.c_contiguous):
        if renormalize:
            raise TypeError("pdfs must be a C-contiguous float64 array (it is renormalised in place, pdf.py:980)")
        pdfs = np.ascontiguousarray(pdfs, dtype=np.float64)
    pgrid = np.ascontiguousarray(pgrid, dtype=np.float64)
    nobj = len(pdfs)
    urand = np.array([rstate.rand() for _ in range(nobj)]) if nobj < 64 else np.ascontiguousarray(rstate.rand(nobj))
    loss = _loss_matrix(pgrid, pkern, pkern_grid)
    eng = SummaryEngine.get()
    est, sd, risk, quant, mc, rowsum = eng.summarize(pdfs, pgrid, loss, urand, renormalize)
    if renormalize:
        pdfs /= rowsum[:, None]          # the same sums, the same IEEE division as pdf.py:980
    if wconf_func is None:
        widths = (1. + est) * 0.03
    else:
        try:
            widths = np.asarray(wconf_func(est), dtype=np.float64)
            if widths.shape != est.shape:
                raise ValueError
        except Exception:
            widths = np.array([[wconf_func(v) for v in row] for row in est], dtype=np.float64)
    conf = eng.conf(est, widths)
    out = tuple((est[k], sd[k], conf[k], risk[k]) for k in range(4))
    return out + ((quant[0], quant[1], quant[2], quant[3]), mc)
