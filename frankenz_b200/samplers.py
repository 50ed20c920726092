"""`loglike_nz` of frankenz/samplers.py:24-76 on the GPU.

The population / hierarchical samplers of the reference evaluate the likelihood of a trial N(z) against the SAME set of
per-object PDFs thousands of times (samplers.py:196-199, 460-470); each evaluation is a (Nobs x Nbins) GEMV plus a sum of
logs.  `NzLikelihood` keeps the PDFs in HBM, so that every call streams them once at HBM bandwidth; the module-level
`loglike_nz` is the drop-in with the reference's signature (it re-uses the resident copy while the caller passes the same
array).  The MCMC loops themselves (sequential Metropolis / Gibbs updates) are out of scope.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._engine import SummaryEngine
from ._lib import dptr, f64

__all__ = ["loglike_nz", "NzLikelihood"]


class NzLikelihood(object):
    """PDFs resident on the device + repeated `loglike_nz` evaluations."""

    def __init__(self, pdfs=None, device=None, device_ptr=None, shape=None):
        self.eng = SummaryEngine(device)
        if device_ptr is not None:
            self.No, self.Ng = int(shape[0]), int(shape[1])
            _lib.check(self.eng.lib.fzb_nz_set_pdfs_dev(self.eng.h, int(device_ptr), self.No, self.Ng))
        else:
            p = f64(pdfs)
            self.No, self.Ng = p.shape
            _lib.check(self.eng.lib.fzb_nz_set_pdfs(self.eng.h, dptr(p), self.No, self.Ng))

    def loglike_nz(self, nz, overlap=None, return_overlap=False, pair=None, pair_step=None):
        nz = f64(nz)
        if overlap is not None:
            # the caller supplies the overlaps (samplers.py:67-68): only the perturbation and the sum remain; numpy on
            # the host is the right tool for an O(Nobs) vector expression
            return _host_from_overlap(nz, overlap, self._columns(pair), pair_step, return_overlap, self.No)
        lnl = C.c_double()
        ov = np.empty(self.No) if return_overlap else None
        pi, pj = (-1, -1)
        step = 0.0
        if pair is not None and pair_step is not None:
            pi, pj = int(pair[0]), int(pair[1])
            step = float(pair_step)
        _lib.check(self.eng.lib.fzb_nz_loglike(self.eng.h, dptr(nz), len(nz), pi, pj, step, C.byref(lnl), dptr(ov)))
        if return_overlap:
            return lnl.value, ov
        return lnl.value

    def _columns(self, pair):
        raise NotImplementedError("a supplied `overlap` with a `pair` needs the PDF columns on the host; call the "
                                  "module-level loglike_nz with the host array instead")

    def stats(self):
        return self.eng.stats()

    def close(self):
        self.eng.close()


def _host_from_overlap(nz, overlap, cols, pair_step, return_overlap, nobs):
    perturb = 0.
    if np.any(~np.isfinite(nz) | (nz < 0.)):
        lnlike, overlap = -np.inf, np.zeros(nobs)
    else:
        if cols is not None and pair_step is not None:
            perturb = pair_step * (cols[0] - cols[1])
        with np.errstate(divide="ignore", invalid="ignore"):
            lnlike = np.sum(np.log(overlap + perturb))
    if return_overlap:
        return lnlike, overlap + perturb
    return lnlike


_cache = {"key": None, "obj": None}


def loglike_nz(nz, pdfs, overlap=None, return_overlap=False, pair=None, pair_step=None):
    """Drop-in for frankenz.samplers.loglike_nz (samplers.py:24-76).  The PDFs are uploaded on the first call and stay
    on the device while the same array object (same buffer, shape) is passed again, as the samplers do."""
    if overlap is not None:
        cols = None if pair is None else (pdfs[:, pair[0]], pdfs[:, pair[1]])
        return _host_from_overlap(f64(nz), overlap, cols, pair_step, return_overlap, len(pdfs))
    p = f64(pdfs)
    key = (p.ctypes.data, p.shape)
    if _cache["key"] != key:
        _cache["obj"] = NzLikelihood(p)
        _cache["key"] = key
    return _cache["obj"].loglike_nz(nz, return_overlap=return_overlap, pair=pair, pair_step=pair_step)
