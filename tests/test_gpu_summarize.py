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


@pytest.mark.parametrize("ng,nobj", [(40, 37), (150, 70), (300, 45), (640, 33), (705, 50), (900, 61), (1056, 17)])
def test_every_kernel_variant_against_the_oracle(fz, ng, nobj):
    """Grid sizes that select each instantiation of k_summarize (2 / 4 / 8 column blocks per warp with 32 objects per CTA,
    12 with 16), ragged last tiles, grids that are not a multiple of 4: against the oracle's numpy arithmetic
    (oracle/fz_oracle.py, bit-exact against the reference's golden vectors).  Selections and interpolations on exactly
    reproduced arrays (mode, quantiles, median, Monte-Carlo draw, renormalised rows) bit for bit, sums to 1e-12."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import fz_oracle as fo
    rs = np.random.RandomState(ng)
    zg = np.sort(rs.uniform(0., 6., ng))
    mu, sg = rs.uniform(0.3, 5.5, nobj), rs.uniform(0.05, 0.7, nobj)
    p = np.exp(-0.5 * ((zg[None, :] - mu[:, None]) / sg[:, None]) ** 2) + 1e-12 * rs.uniform(size=(nobj, ng))
    p[3] = 0.0
    p[3, ng // 2] = 2.5                       # a single spike: plateaus in the CDF
    a, b = p.copy(), p.copy()
    got = fz.pdf.pdfs_summarize(a, zg, rstate=np.random.RandomState(11))
    ref = fo.pdfs_summarize(b, zg, rstate=np.random.RandomState(11))
    assert np.array_equal(a, b)               # renormalised in place, numpy's pairwise row sums
    for k, name in enumerate(NAMES):
        for j in range(4):
            if name in ("mode", "med") and j == 0:
                assert np.array_equal(got[k][j], ref[k][j]), (name, j)
            else:
                assert close(got[k][j], ref[k][j], 1e-10 if j == 2 else 1e-11), (name, j, np.max(np.abs(got[k][j] - ref[k][j])))
    for j in range(4):
        assert np.array_equal(got[4][j], ref[4][j]), j
    assert np.array_equal(got[5], ref[5])


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


def test_summaries_fused_behind_fit_predict(sdss_mock):
    """SURVEY 8f rank 2 as written: fit_predict(..., summarize=True, return_pdfs=False) computes the summaries on the
    device from PDFs that never reach the host; they must equal pdfs_summarize applied to the PDFs of fit_predict."""
    import frankenz_b200 as fz
    from oracle import fz_oracle as fo
    phot, err, z = sdss_mock
    m, me, mm = phot[:3000].copy(), err[:3000].copy(), np.ones((3000, 5))
    x, xe, xm = phot[3500:3900].copy(), err[3500:3900].copy(), np.ones((400, 5))
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    rdict = fz.pdf.PDFDict(zgrid, np.linspace(0.005, 2, 500))
    labe = np.full(3000, 0.05)
    for kw in (dict(), dict(free_scale=True, ignore_model_err=True)):
        bf = fz.BruteForce(m, me, mm)
        pdfs, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z[:3000], labe, label_dict=rdict, return_gof=True,
                                        verbose=False, save_fits=False, lprob_kwargs=kw)
        ref = fz.pdf.pdfs_summarize(pdfs.copy(), zgrid, rstate=np.random.RandomState(3))
        got, (lm2, le2) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z[:3000], labe, label_dict=rdict,
                                         return_gof=True, verbose=False, save_fits=False, lprob_kwargs=kw,
                                         summarize=True, return_pdfs=False,
                                         summarize_kwargs=dict(rstate=np.random.RandomState(3)))
        assert np.array_equal(lm, lm2) and np.allclose(le, le2, rtol=0, atol=1e-9)
        for a, b in zip(got[:4], ref[:4]):
            for u, v in zip(a, b):          # estimate, std, conf, risk: the PDFs of two runs differ by fp32 atomics order
                assert np.allclose(u, v, rtol=0, atol=2e-5), np.max(np.abs(u - v))
        for u, v in zip(got[4], ref[4]):
            assert np.allclose(u, v, rtol=0, atol=2e-5)
        assert np.allclose(got[5], ref[5], rtol=0, atol=2e-5)
        # with the PDFs returned as well: summaries of exactly those PDFs, so everything but fp rounding is identical
        got2, p2 = bf.fit_predict_summarize(x.copy(), xe.copy(), xm.copy(), z[:3000], labe, label_dict=rdict,
                                            lprob_kwargs=kw, verbose=False, return_pdfs=True,
                                            rstate=np.random.RandomState(3))
        ref2 = fz.pdf.pdfs_summarize(p2.copy(), zgrid, rstate=np.random.RandomState(3))
        for a, b in zip(got2[:4], ref2[:4]):
            assert np.array_equal(a[0], b[0]) and np.allclose(a[1], b[1], rtol=1e-12, atol=0)     # estimate, std
            assert np.array_equal(a[2], b[2]) and np.allclose(a[3], b[3], rtol=1e-12, atol=1e-15)  # conf, risk
        for u, v in zip(got2[4], ref2[4]):
            assert np.array_equal(u, v)
        assert np.array_equal(got2[5], ref2[5])
        # oracle (the reference's arithmetic, bit-exact against its golden vectors) on the same PDFs
        oref = fo.pdfs_summarize(p2.copy(), zgrid, rstate=np.random.RandomState(3))
        for a, b in zip(got2[:4], oref[:4]):
            assert np.allclose(a[0], b[0], rtol=0, atol=1e-12) and np.allclose(a[3], b[3], rtol=1e-10, atol=1e-12)
