"""`samplers.loglike_nz` on the GPU against the reference's golden vectors
(tests/golden/make_golden.py nz) and against the oracle at a size where the GEMV is HBM-bound."""
import numpy as np
import pytest

from conftest import golden
from oracle import fz_oracle as fo

pytestmark = pytest.mark.gpu


def test_loglike_nz_matches_reference_golden():
    from frankenz_b200 import samplers
    g = golden("loglike_nz.npz")
    p = g["pdfs"]
    for nz, ref in zip(g["nz"], g["lnlike"]):
        assert abs(samplers.loglike_nz(nz, p) - ref) <= 1e-12 * abs(ref)
    ll, ov = samplers.loglike_nz(g["nz"][1], p, return_overlap=True)
    assert abs(ll - g["lnlike_ov"]) <= 1e-12 * abs(ll) and np.allclose(ov, g["overlap"], rtol=1e-13, atol=0)
    ll, ov = samplers.loglike_nz(g["nz"][2], p, return_overlap=True, pair=(120, 260), pair_step=3e-4)
    assert abs(ll - g["lnlike_pair"]) <= 1e-12 * abs(ll) and np.allclose(ov, g["overlap_pair"], rtol=1e-12, atol=1e-300)
    ll, ov = samplers.loglike_nz(g["nz_bad"], p, return_overlap=True)
    assert ll == -np.inf and not ov.any()
    # overlaps supplied by the caller (samplers.py:67-68): host expression, same numbers as the reference
    ll2 = samplers.loglike_nz(g["nz"][1], p, overlap=g["overlap"])
    assert ll2 == g["lnlike_ov"]


def test_loglike_nz_resident_pdfs_large():
    from frankenz_b200.samplers import NzLikelihood
    rs = np.random.RandomState(5)
    no, ng = 200000, 701
    p = rs.gamma(0.3, size=(no, ng))
    p /= p.sum(axis=1)[:, None]
    nl = NzLikelihood(p)
    for k in range(3):
        nz = rs.dirichlet(np.ones(ng))
        ref = fo.loglike_nz(nz, p)
        got, ov = nl.loglike_nz(nz, return_overlap=True)
        assert abs(got - ref) <= 1e-11 * abs(ref)
        assert np.allclose(ov, p @ nz, rtol=1e-13, atol=0)
    ms = nl.stats()["ms_total"]
    assert ms > 0 and no * ng * 8 / (ms * 1e-3) > 1e12        # streams the PDFs at > 1 TB/s
    nl.close()
