"""pdf.pdfs_summarize / pdfs_resample on the GPU (csrc/fzb_summarize.cu) against the outputs of the unmodified reference
(tests/golden/pdfs_summarize.npz, frankenz/pdf.py:855-1074).  Everything that is selection or interpolation on exactly
reproduced arrays (row sums in numpy's pairwise order, sequential CDFs, numpy.interp) must match bit for bit: mode,
quantiles, median, Monte-Carlo draw, the renormalised PDFs; sums whose order differs from numpy's BLAS / pairwise order
(mean, risk product, standard deviations) and what is interpolated at them are held to 1e-12."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
NAMES = ("mean", "med", "mode", "best")


@pytest.fixture(scope="module")
def fz():
    import frankenz_b200
    return frankenz_b200


def close(a, b, tol=1e-12):
    return np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b)))


def check(res, g, tag):
    for k, n in enumerate(NAMES):
        est, sd, conf, risk = res[k]
        if n in ("mode", "med"):
            assert np.array_equal(est, g["%s_%s" % (tag, n)]), (tag, n)
        else:
            assert close(est, g["%s_%s" % (tag, n)]), (tag, n)
        assert close(sd, g["%s_%s_std" % (tag, n)], 1e-11), (tag, n)
        assert close(conf, g["%s_%s_conf" % (tag, n)], 1e-10), (tag, n)
        assert close(risk, g["%s_%s_risk" % (tag, n)], 1e-11), (tag, n)
    for j, q in enumerate(("low95", "low68", "high68", "high95")):
        assert np.array_equal(res[4][j], g["%s_%s" % (tag, q)]), (tag, q)
    assert np.array_equal(res[5], g[tag + "_mc"]), tag


@pytest.mark.parametrize("tag,kw,seed", [("lorentz", dict(pkern="lorentz"), 5), ("gaussian", dict(pkern="gaussian"), 5),
                                         ("tophat", dict(pkern="tophat"), 5), ("noren", dict(renormalize=False), 6),
                                         ("custom", dict(pkern=lambda x: np.exp(-np.abs(x)),
                                                         wconf_func=lambda z: 0.02 + 0.05 * z * z), 7)])
def test_pdfs_summarize_matches_reference(fz, tag, kw, seed):
    g = golden("pdfs_summarize.npz")
    p = (g["pdfs_normed"] if tag == "noren" else g["pdfs"]).copy()
    res = fz.pdf.pdfs_summarize(p, g["zgrid"], rstate=np.random.RandomState(seed), **kw)
    check(res, g, tag)
    assert np.array_equal(p, g[tag + "_pdfs_after"])          # in-place renormalisation, pdf.py:980


def test_user_kernel_grid_small_grid_and_scalar_only_wconf(fz):
    g = golden("pdfs_summarize.npz")
    q = g["pdfs2"].copy()
    res = fz.pdf.pdfs_summarize(q, g["grid2"], rstate=np.random.RandomState(8), pkern="gaussian", pkern_grid=g["kgrid2"])
    check(res, g, "grid2")
    import math
    p = g["pdfs"].copy()
    res = fz.pdf.pdfs_summarize(p, g["zgrid"], rstate=np.random.RandomState(7), pkern=lambda x: np.exp(-np.abs(x)),
                                wconf_func=lambda z: 0.02 + 0.05 * math.pow(z, 2) if z == z else 0.0)   # scalars only
    assert close(res[0][2], g["custom_mean_conf"], 1e-10)
    with pytest.raises(RuntimeError):
        fz.pdf.pdfs_summarize(g["pdfs"].copy(), g["zgrid"], pkern="no such kernel")


def test_resample_and_batching(fz):
    g = golden("pdfs_summarize.npz")
    assert np.array_equal(fz.pdf.pdfs_resample(g["pdfs"].copy(), g["zgrid"], g["resample_grid"]), g["resampled"])
    assert np.array_equal(fz.pdf.pdfs_resample(g["pdfs"].copy(), g["zgrid"], g["resample_grid"], renormalize=False,
                                               left=0.5, right=0.25), g["resampled_noren"])
    # objects are independent: tiling the batch (several CTAs, a partial last one) repeats the rows
    p = np.tile(g["pdfs"], (13, 1))[:1237].copy()
    res = fz.pdf.pdfs_summarize(p, g["zgrid"], rstate=np.random.RandomState(5))
    one = fz.pdf.pdfs_summarize(g["pdfs"].copy(), g["zgrid"], rstate=np.random.RandomState(5))
    n = len(g["pdfs"])
    for k in range(4):
        for j in range(4):
            assert np.array_equal(res[k][j][n:2 * n], one[k][j]) or j == 2 or k == 3, (k, j)
    assert np.array_equal(res[4][0][:n], one[4][0]) and np.array_equal(res[4][3][n:2 * n], one[4][3])
