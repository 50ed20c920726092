"""SOM / GNG node-fit stages against the reference: a SelfOrganizingMap trained by the
unmodified reference on the SDSS mock supplies the nodes (tests/golden/make_network_golden.py); `populate_network` and
`fit` run on the GPU and must reproduce the reference's node lists, weights, neighbour lists and fits."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def unragged(g, name):
    off, flat = g[name + "_off"], g[name]
    return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.fixture(scope="module")
def net():
    from frankenz_b200.networks import NetworkFit
    g = golden("som_nodefit.npz")
    n = NetworkFit(g["models"], g["models_err"], g["models_mask"], g["nodes"])
    n.populate_network(verbose=False)
    return g, n


def test_populate_network_matches_reference(net):
    g, n = net
    assert np.array_equal(n.nodes_Nmatch, g["nodes_Nmatch"])
    for name in ("nodes_idxs", "nodes_bmus"):
        ref = unragged(g, name)
        got = getattr(n, name)
        assert len(ref) == len(got)
        for a, b in zip(got, ref):
            assert np.array_equal(np.asarray(a, dtype=np.int64), b)
    for name in ("nodes_logwts", "nodes_scales", "nodes_scales_err"):
        for a, b in zip(getattr(n, name), unragged(g, name)):
            assert np.allclose(np.asarray(a), b, rtol=1e-9, atol=1e-12)
    assert np.allclose(n.models_lmap, g["models_lmap"], rtol=1e-10, atol=0)
    assert np.allclose(n.models_levid, g["models_levid"], rtol=1e-10, atol=0)


@pytest.mark.parametrize("tag,nodes_only", [("nodes", True), ("full", False)])
def test_fit_through_the_network_matches_reference(net, tag, nodes_only):
    g, n = net
    x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
    kw = dict(free_scale=True, ignore_model_err=True, dim_prior=True)
    n.fit(x, xe, xm, nodes_only=nodes_only, lprob_kwargs=kw, verbose=False)
    assert np.array_equal(n.Nneighbors, g[tag + "_Nneighbors"])
    assert not np.isnan(x[7, 4]) and xm[7, 4] == 0          # cleaned in place like logprob does (pdf.py:310-311)
    for a, b in zip(n.neighbors, unragged(g, tag + "_neighbors")):
        assert np.array_equal(np.asarray(a, dtype=np.int64), b)
    for a, b in zip(n.fit_lnprob, unragged(g, tag + "_lnprob")):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-11, equal_nan=True)
    for a, b in zip(n.fit_chi2, unragged(g, tag + "_chi2")):
        assert np.allclose(a, b, rtol=1e-9, atol=1e-11, equal_nan=True)


def test_predict_from_network_fits(net):
    """PDFs over each object's neighbour list: the same KDE kernel as the kNN estimator, against the oracle."""
    import frankenz_b200 as fz
    from oracle import fz_oracle as fo
    g, n = net
    x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
    n.fit(x, xe, xm, nodes_only=False, lprob_kwargs=dict(free_scale=True, ignore_model_err=True), verbose=False)
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    sig = np.linspace(0.005, 2, 500)
    z = golden("sdss_cww_mock.npz")["redshifts"][:len(g["models"])]
    labe = np.full(len(z), 0.05)
    pdfs = n.predict(z, labe, label_dict=fz.pdf.PDFDict(zgrid, sig), verbose=False)
    kd = fo.KernelDict(zgrid, sig)
    yi, si = kd.fit(z, labe)
    for i in (0, 5, 17, 59):
        idx, lw = n.neighbors[i], n.fit_lnprob[i]
        wt = np.exp(lw - lw.max())
        ref = fo.kde_dict(kd, yi[idx], si[idx], y_wt=wt)
        ref /= ref.sum()
        assert np.sum(np.abs(pdfs[i] - ref)) <= 1e-9
