"""Round-2 API additions, on the GPU through the C ABI: the module-level KDE functions (pdf.py:444-622), the internal
likelihood names, label checks that only fire for selected models (pdf.py:603-620) and the streaming generator twins."""
import numpy as np
import pytest

from conftest import golden, max_rel, same_special
from oracle import fz_oracle as fo

pytestmark = pytest.mark.gpu


def test_gauss_kde_dict_and_grid_match_oracle():
    import frankenz_b200 as fz
    g = golden("kde_edges.npz")
    zgrid = g["zgrid"]
    sig = np.linspace(0.005, 2, 500)
    rdict = fz.pdf.PDFDict(zgrid, sig)
    kd = fo.KernelDict(zgrid, sig)
    for row in range(len(g["logwt"])):
        lw = g["logwt"][row]
        if not np.all(np.isfinite(lw)):
            continue
        wt = np.exp(lw - lw.max())
        ref = fo.kde_dict(kd, g["y_idx"], g["y_std_idx"], y_wt=wt)
        got = fz.pdf.gauss_kde_dict(rdict, y_idx=g["y_idx"], y_std_idx=g["y_std_idx"], y_wt=wt)
        assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
        got2 = fz.pdf.gauss_kde_dict(rdict, y=g["labels"], y_std=g["label_errs"], y_wt=wt)
        assert np.max(np.abs(got2 - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
        refg = fo.kde_grid(g["labels"], g["label_errs_grid"], zgrid, y_wt=wt)
        gotg = fz.pdf.gauss_kde(g["labels"], g["label_errs_grid"], zgrid, y_wt=wt)
        assert np.max(np.abs(gotg - refg)) <= 1e-11 * max(1.0, np.max(np.abs(refg)))
    # no weights: every label counts once; the CDF rule; all-zero weights give an empty PDF
    ref = fo.kde_dict(kd, g["y_idx"], g["y_std_idx"])
    assert np.max(np.abs(fz.pdf.gauss_kde_dict(rdict, y_idx=g["y_idx"], y_std_idx=g["y_std_idx"]) - ref)) <= 1e-11
    wt = np.linspace(0.1, 1.0, len(g["y_idx"]))
    ref = fo.kde_dict(kd, g["y_idx"], g["y_std_idx"], y_wt=wt, wt_thresh=None, cdf_thresh=0.05)
    got = fz.pdf.gauss_kde_dict(rdict, y_idx=g["y_idx"], y_std_idx=g["y_std_idx"], y_wt=wt, wt_thresh=None,
                                cdf_thresh=0.05)
    assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))
    assert not fz.pdf.gauss_kde_dict(rdict, y_idx=g["y_idx"], y_std_idx=g["y_std_idx"], y_wt=np.zeros(14)).any()
    with pytest.raises(ValueError):
        fz.pdf.gauss_kde_dict(rdict)


def test_internal_likelihood_names():
    import frankenz_b200 as fz
    g = golden("loglike_combos.npz")
    m, me, mm = g["models"], g["models_err"], g["models_mask"]
    x, xe, xm = g["data"][0].copy(), g["data_err"][0].copy(), g["data_mask"][0].copy()
    a = fz.pdf._loglike(x.copy(), xe.copy(), xm.copy(), m, me, mm, ignore_model_err=False, dim_prior=True)
    b = fz.pdf.loglike(x.copy(), xe.copy(), xm.copy(), m, me, mm)
    for u, v in zip(a, b):
        assert np.array_equal(u, v, equal_nan=True)
    a = fz.pdf._loglike_s(x.copy(), xe.copy(), xm.copy(), m, me, mm, ignore_model_err=True, return_scale=True)
    b = fz.pdf.loglike(x.copy(), xe.copy(), xm.copy(), m, me, mm, free_scale=True, ignore_model_err=True,
                       return_scale=True)
    assert len(a) == 5
    for u, v in zip(a, b):
        assert np.array_equal(u, v, equal_nan=True)
    bins = np.linspace(-3, 3, 13)
    from scipy.special import erf
    cdf = 0.5 * (1 + erf((bins - 0.2) / (np.sqrt(2) * 0.7)))
    assert np.array_equal(fz.pdf.gaussian_bin(0.2, 0.7, bins), cdf[1:] - cdf[:-1])


def test_labels_off_the_grid_raise_only_when_selected(sdss_mock):
    """pdf.py:603-620: a training label beyond the grid is harmless until the weight threshold selects it."""
    import frankenz_b200 as fz
    from frankenz_b200._lib import FzbError
    phot, err, z = sdss_mock
    m, me, mm = phot[:1500].copy(), err[:1500].copy(), np.ones((1500, 5))
    x, xe, xm = phot[3000:3040].copy(), err[3000:3040].copy(), np.ones((40, 5))
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    sig = np.linspace(0.005, 2, 500)
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(1500, 0.05)
    lab = z[:1500].copy()
    bf = fz.BruteForce(m, me, mm)
    p0, (lm0, le0) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), lab, labe, label_dict=rdict, return_gof=True,
                                    verbose=False, save_fits=False)
    # a hopeless model (fluxes 1e6 times too bright) gets a label far outside the grid: never selected -> same PDFs
    m2 = m.copy()
    m2[7] *= 1e6
    lab2 = lab.copy()
    lab2[7] = 9.5
    bf2 = fz.BruteForce(m2, me, mm)
    p1, (lm1, le1) = bf2.fit_predict(x.copy(), xe.copy(), xm.copy(), lab2, labe, label_dict=rdict, return_gof=True,
                                     verbose=False, save_fits=False)
    kd = fo.KernelDict(zgrid, sig)
    po, lmo, leo = fo.bruteforce_fit_predict(m2, me, mm, x.copy(), xe.copy(), xm.copy(), lab2, labe, label_dict=kd)
    assert np.max(np.sum(np.abs(p1 - po), axis=1)) <= 1e-9 and np.allclose(lm1, lmo, rtol=1e-10)
    # the same label on a model that IS selected raises, like the reference's IndexError / shape error
    best = int(bf.best_idx[0])
    lab3 = lab.copy()
    lab3[best] = 9.5
    with pytest.raises(FzbError):
        bf.fit_predict(x.copy(), xe.copy(), xm.copy(), lab3, labe, label_dict=rdict, verbose=False, save_fits=False)
    # and the handle recovers
    p2 = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), lab, labe, label_dict=rdict, verbose=False, save_fits=False)
    assert np.max(np.sum(np.abs(p2 - p0), axis=1)) <= 1e-12


def test_generator_twins_stream_in_chunks(sdss_mock, monkeypatch):
    import frankenz_b200 as fz
    from frankenz_b200 import bruteforce as bfmod
    phot, err, z = sdss_mock
    m, me, mm = phot[:900].copy(), err[:900].copy(), np.ones((900, 5))
    x, xe, xm = phot[3000:3050].copy(), err[3000:3050].copy(), np.ones((50, 5))
    bf = fz.BruteForce(m, me, mm)
    bf.fit(x.copy(), xe.copy(), xm.copy(), verbose=False)
    full = bf.fit_lnprob.copy()
    monkeypatch.setattr(bfmod, "STREAM_BYTES", 7 * 8 * 900 * 8)        # 8 objects per chunk
    bf2 = fz.BruteForce(m, me, mm)
    rows = [r[2].copy() for r in bf2._fit(x.copy(), xe.copy(), xm.copy(), save_fits=False)]
    assert bf2.fit_lnprob is None and len(rows) == 50
    assert np.array_equal(np.array(rows), full)
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    rdict = fz.pdf.PDFDict(zgrid, np.linspace(0.005, 2, 500))
    labe = np.full(900, 0.05)
    ref = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z[:900], labe, label_dict=rdict, verbose=False, save_fits=False)
    monkeypatch.setattr(bfmod, "STREAM_BYTES", 8 * 701 * 16)
    got = [p.copy() for p, _ in bf2._fit_predict(x.copy(), xe.copy(), xm.copy(), z[:900], labe, label_dict=rdict,
                                                 save_fits=False)]
    assert np.max(np.abs(np.array(got) - ref)) <= 1e-12


def test_knn_fit_predict_without_saved_fits(sdss_mock):
    """NearestNeighbors.fit_predict(save_fits=False) keeps search, union, fits and KDE on the device
    (fzb_knn_fit_predict); the PDFs must equal those of the fit + predict route, and nothing is stored."""
    import frankenz_b200 as fz
    phot, err, z = sdss_mock
    m, me, mm = phot[:4500].copy(), err[:4500].copy(), np.ones((4500, 5))
    x, xe, xm = phot[4500:4800].copy(), err[4500:4800].copy(), np.ones((300, 5))
    x[3, 2] = np.nan
    depth = golden("sdss_cww_mock.npz")["depth_flux1sig"]
    kw = dict(skynoise=depth, zeropoints=10 ** (-0.4 * -23.9))
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    rdict = fz.pdf.PDFDict(zgrid, np.linspace(0.005, 2, 500))
    labe = np.full(4500, 0.05)
    for lk in (dict(), dict(free_scale=True, ignore_model_err=True)):
        nn = fz.NearestNeighbors(m, me, mm, K=4, fmap_kwargs=kw, rstate=np.random.RandomState(1), verbose=False)
        p1, (lm1, le1) = nn.fit_predict(x.copy(), xe.copy(), xm.copy(), z[:4500], labe, label_dict=rdict, k=12, eps=0,
                                        rstate=np.random.RandomState(2), return_gof=True, verbose=False, lprob_kwargs=lk)
        nnb = nn.Nneighbors.copy()
        nn2 = fz.NearestNeighbors(m, me, mm, K=4, fmap_kwargs=kw, rstate=np.random.RandomState(1), verbose=False)
        p2, (lm2, le2) = nn2.fit_predict(x.copy(), xe.copy(), xm.copy(), z[:4500], labe, label_dict=rdict, k=12, eps=0,
                                         rstate=np.random.RandomState(2), return_gof=True, verbose=False,
                                         lprob_kwargs=lk, save_fits=False)
        assert nn2.fit_lnprob is None and nn2.neighbors is None
        assert np.array_equal(nn2.Nneighbors_last, nnb)
        assert np.allclose(p1, p2, rtol=0, atol=1e-13) and np.array_equal(lm1, lm2)
        assert np.allclose(le1, le2, rtol=1e-14, atol=0)
