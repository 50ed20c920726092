"""Likelihood and PDF helpers with the reference's signatures (frankenz/pdf.py:21-24).

`loglike` / `logprob` evaluate on the GPU through the C ABI (one object against all
models).  `PDFDict` tabulates the truncated Gaussian kernels that the CUDA KDE kernels
consume; the magnitude transforms are the feature maps of the kNN estimator.
"""
import numpy as np

from ._engine import Engine, clean_inplace, make_config

__all__ = ["loglike", "logprob", "gaussian", "magnitude", "inv_magnitude", "luptitude", "inv_luptitude", "PDFDict"]


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


def gaussian(mu, std, x):
    """N(x | mu, std) on the grid `x` (frankenz/pdf.py:414-425); used to tabulate `PDFDict`."""
    z = (x - mu) / std
    return np.exp(-0.5 * np.square(z)) / (np.sqrt(2. * np.pi) * std)


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
