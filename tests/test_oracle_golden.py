"""The oracle against the reference's own outputs (tests/golden/*.npz). CPU only."""
import numpy as np
import pytest

from conftest import golden, same_special
from oracle import fz_oracle as fo

COMBOS = [(fs, ime, dp) for fs in (False, True) for ime in (False, True) for dp in (False, True)]


def tag(fs, ime, dp):
    return "fs%d_ime%d_dp%d" % (fs, ime, dp)


def exact(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("fname", ["loglike_combos.npz", "loglike_degenerate.npz"])
@pytest.mark.parametrize("fs,ime,dp", COMBOS)
def test_loglike_bit_exact(fname, fs, ime, dp):
    g = golden(fname)
    m, me, mm = g["models"], g["models_err"], g["models_mask"]
    x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
    t = tag(fs, ime, dp)
    for i in range(len(x)):
        r = fo.loglike(x[i], xe[i], xm[i], m, me, mm, free_scale=fs, ignore_model_err=ime, dim_prior=dp,
                       ltol=1e-4, return_scale=True)
        assert exact(r[0], g[t + "_lnl"][i]), (t, i)
        assert exact(r[1], g[t + "_ndim"][i])
        assert exact(r[2], g[t + "_chi2"][i])
        if fs:
            assert exact(r[3], g[t + "_scale"][i])
            assert exact(r[4], g[t + "_scale_err"][i])
    if "cleaned_data" in g.files:   # in-place cleaning (pdf.py:310-311)
        assert exact(x, g["cleaned_data"]) and exact(xe, g["cleaned_err"]) and exact(xm, g["cleaned_mask"])


def test_logprob_protocol():
    g = golden("loglike_combos.npz")
    r = fo.logprob(g["data"][0].copy(), g["data_err"][0].copy(), g["data_mask"][0].copy(), g["models"],
                   g["models_err"], g["models_mask"])
    assert len(r) == 5 and np.all(r[0] == 0) and exact(r[1], r[2])
    lp = np.linspace(-1, 0, len(g["models"]))
    r2 = fo.logprob(g["data"][0].copy(), g["data_err"][0].copy(), g["data_mask"][0].copy(), g["models"],
                    g["models_err"], g["models_mask"], lnprior=lp)
    assert exact(r2[2], r[1] + lp)


def test_fs1_ltol():
    g = golden("fs1_ltol.npz")
    for ltol in (1e-2, 1e-4, 1e-7):
        for i in range(len(g["data"])):
            r = fo.loglike(g["data"][i].copy(), g["data_err"][i].copy(), g["data_mask"][i].copy(), g["models"],
                           g["models_err"], g["models_mask"], free_scale=True, dim_prior=False, ltol=ltol)
            assert exact(r[0], g["lnl_ltol%g" % ltol][i])


def _dict():
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    return zgrid, fo.KernelDict(zgrid, np.linspace(0.005, 2, 500))


def test_bruteforce_fit_predict():
    g = golden("bruteforce_c1small.npz")
    zgrid, kd = _dict()
    m, me, mm = g["models"], g["models_err"], g["models_mask"]
    fit = fo.bruteforce_fit(m, me, mm, g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy())
    assert exact(fit["lnprob"], g["fit_lnprob"]) and exact(fit["chi2"], g["fit_chi2"])
    assert np.array_equal(fit["Ndim"], g["fit_Ndim"])
    lab, labe = g["labels"], g["label_errs"]
    p, lm, le = fo.bruteforce_predict(fit["lnprob"], lab, labe, label_dict=kd)
    assert exact(p, g["pdf_dict"]) and exact(lm, g["lmap"]) and exact(le, g["levid"])
    p, _, _ = fo.bruteforce_predict(fit["lnprob"], lab, labe, label_grid=zgrid)
    assert np.allclose(p, g["pdf_grid"], rtol=1e-13, atol=1e-300)
    p, _, _ = fo.bruteforce_predict(fit["lnprob"], lab, g["label_errs2"], label_dict=kd)
    assert exact(p, g["pdf_dict_mixed"])
    p, _, _ = fo.bruteforce_predict(fit["lnprob"], lab, g["label_errs2"], label_grid=zgrid)
    assert np.allclose(p, g["pdf_grid_mixed"], rtol=1e-13, atol=1e-300)
    p, _, _ = fo.bruteforce_predict(fit["lnprob"], lab, labe, label_dict=kd, wt_thresh=None, cdf_thresh=None)
    assert exact(p, g["pdf_dict_nothresh"])
    p, _, _ = fo.bruteforce_predict(fit["lnprob"], lab, labe, label_dict=kd, wt_thresh=None, cdf_thresh=2e-4)
    assert exact(p, g["pdf_dict_cdf"])
    p, _, _ = fo.bruteforce_predict(fit["lnprob"], lab, labe, label_grid=zgrid, wt_thresh=None, cdf_thresh=2e-4)
    assert np.allclose(p, g["pdf_grid_cdf"], rtol=1e-13, atol=1e-300)


@pytest.mark.parametrize("fs,ime,dp", COMBOS)
def test_bruteforce_fused(fs, ime, dp):
    g = golden("bruteforce_c1small.npz")
    _, kd = _dict()
    p, lm, le = fo.bruteforce_fit_predict(g["models"], g["models_err"], g["models_mask"], g["data"].copy(),
                                          g["data_err"].copy(), g["data_mask"].copy(), g["labels"],
                                          g["label_errs"], label_dict=kd, free_scale=fs, ignore_model_err=ime,
                                          dim_prior=dp)
    t = tag(fs, ime, dp)
    assert exact(p, g[t + "_pdf"]) and exact(lm, g[t + "_lmap"]) and exact(le, g[t + "_levid"])


def test_kde_edges():
    g = golden("kde_edges.npz")
    zgrid, kd = _dict()
    yi, si = kd.fit(g["labels"], g["label_errs"])
    assert np.array_equal(yi, g["y_idx"]) and np.array_equal(si, g["y_std_idx"])
    with np.errstate(all="ignore"):
        p, lm, le = fo.bruteforce_predict(g["logwt"], g["labels"], g["label_errs"], label_dict=kd)
        assert exact(p, g["pdf_dict"]) and exact(lm, g["lmap"]) and exact(le, g["levid"])
        p, _, _ = fo.bruteforce_predict(g["logwt"], g["labels"], g["label_errs_grid"], label_grid=zgrid)
        assert same_special(p, g["pdf_grid"])
        assert np.allclose(p, g["pdf_grid"], rtol=1e-13, atol=1e-300, equal_nan=True)
    d2 = fo.KernelDict(g["d2_grid"], g["d2_sig"], sigma_trunc=4.0)
    assert np.array_equal(d2.sigma_width, g["d2_width"])
    for i in (0, 3, 5):
        assert exact(d2.sigma_dict[i], g["d2_kernel%d" % i]) and exact(d2.sigma_dict_cdf[i], g["d2_cdf%d" % i])


@pytest.mark.parametrize("name,K,k,fmap", [("a", 5, 20, "luptitude"), ("b", 1, 1, "luptitude"),
                                           ("c", 8, 7, "identity"), ("d", 3, 25, "magnitude")])
def test_knn_exact(name, K, k, fmap):
    g = golden("knn_exact.npz")
    zgrid, kd = _dict()
    m, me, mm = g["models"], g["models_err"], g["models_mask"]
    x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
    kw = {}
    if fmap == "luptitude":
        kw = dict(skynoise=g["skynoise"], zeropoints=float(g["zeropoints"]))
    elif fmap == "magnitude":
        kw = dict(zeropoints=float(g["zeropoints"]))
        m, me = np.abs(m) + 5 * me, me * 1e-3
        x, xe = np.abs(x) + 5 * xe, xe * 1e-3
    feats = fo.knn_train_features(m, me, K, feature_map=fmap, fmap_kwargs=kw, rstate=np.random.RandomState(1))
    assert np.array_equal(feats, g[name + "_feats"])
    fit = fo.knn_fit(m, me, mm, feats, x, xe, xm, k=k, p=2, feature_map=fmap, fmap_kwargs=kw,
                     rstate=np.random.RandomState(2))
    assert np.array_equal(fit["Nneighbors"], g[name + "_Nneighbors"])
    assert np.array_equal(fit["neighbors"], g[name + "_neighbors"])
    assert exact(fit["lnprob"], g[name + "_lnprob"]) and exact(fit["chi2"], g[name + "_chi2"])
    p, lm, le = fo.knn_predict(fit, g["labels"], g["label_errs"], label_dict=kd)
    assert exact(p, g[name + "_pdf"]) and exact(lm, g[name + "_lmap"]) and exact(le, g[name + "_levid"])
    p, _, _ = fo.knn_predict(fit, g["labels"], g["label_errs"], label_grid=zgrid)
    assert np.allclose(p, g[name + "_pdf_grid"], rtol=1e-13, atol=1e-300)


# ---- PDF summaries ----------------------------------------------------------------
SUMM_KEYS = ["%s%s" % (n, q) for n in ("mean", "med", "mode", "best") for q in ("", "_std", "_conf", "_risk")] + \
            ["low95", "low68", "high68", "high95", "mc"]


def _summ_flat(res):
    return [res[k][j] for k in range(4) for j in range(4)] + list(res[4]) + [res[5]]


@pytest.mark.parametrize("tag,kw,seed", [("lorentz", dict(pkern="lorentz"), 5), ("gaussian", dict(pkern="gaussian"), 5),
                                         ("tophat", dict(pkern="tophat"), 5), ("noren", dict(renormalize=False), 6),
                                         ("custom", dict(pkern=lambda x: np.exp(-np.abs(x)),
                                                         wconf_func=lambda z: 0.02 + 0.05 * z * z), 7)])
def test_pdfs_summarize_bit_exact(tag, kw, seed):
    g = golden("pdfs_summarize.npz")
    p = (g["pdfs_normed"] if tag == "noren" else g["pdfs"]).copy()
    res = fo.pdfs_summarize(p, g["zgrid"], rstate=np.random.RandomState(seed), **kw)
    for name, got in zip(SUMM_KEYS, _summ_flat(res)):
        assert exact(got, g["%s_%s" % (tag, name)]), (tag, name)
    assert exact(p, g[tag + "_pdfs_after"])          # in-place renormalisation (pdf.py:980)


def test_pdfs_summarize_user_kernel_grid_and_resample():
    g = golden("pdfs_summarize.npz")
    q = g["pdfs2"].copy()
    res = fo.pdfs_summarize(q, g["grid2"], rstate=np.random.RandomState(8), pkern="gaussian", pkern_grid=g["kgrid2"])
    for name, got in zip(SUMM_KEYS, _summ_flat(res)):
        assert exact(got, g["grid2_" + name]), name
    assert exact(fo.pdfs_resample(g["pdfs"].copy(), g["zgrid"], g["resample_grid"]), g["resampled"])
    assert exact(fo.pdfs_resample(g["pdfs"].copy(), g["zgrid"], g["resample_grid"], renormalize=False, left=0.5,
                                  right=0.25), g["resampled_noren"])


def _unragged(g, name):
    off, flat = g[name + "_off"], g[name]
    return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def test_loglike_nz_matches_reference():
    """samplers.py:24-76."""
    g = golden("loglike_nz.npz")
    p = g["pdfs"]
    for nz, ref in zip(g["nz"], g["lnlike"]):
        assert fo.loglike_nz(nz, p) == ref
    ll, ov = fo.loglike_nz(g["nz"][1], p, return_overlap=True)
    assert ll == g["lnlike_ov"] and np.array_equal(ov, g["overlap"])
    ll, ov = fo.loglike_nz(g["nz"][2], p, return_overlap=True, pair=(120, 260), pair_step=3e-4)
    assert ll == g["lnlike_pair"] and np.array_equal(ov, g["overlap_pair"])
    ll, ov = fo.loglike_nz(g["nz_bad"], p, return_overlap=True)
    assert ll == -np.inf and not ov.any()


def test_network_node_fit_matches_reference():
    """networks.py:246-356 and :782-936 on the nodes of a reference-trained SOM."""
    g = golden("som_nodefit.npz")
    m, me, mm, nodes = g["models"], g["models_err"], g["models_mask"], g["nodes"]
    pop = fo.network_populate(m, me, mm, nodes)
    assert np.array_equal(pop["nodes_Nmatch"], g["nodes_Nmatch"])
    for name in ("nodes_idxs", "nodes_bmus"):
        for a, b in zip(pop[name], _unragged(g, name)):
            assert np.array_equal(np.asarray(a, dtype=np.int64), b)
    for name in ("nodes_logwts", "nodes_scales", "nodes_scales_err"):
        for a, b in zip(pop[name], _unragged(g, name)):
            assert np.array_equal(np.asarray(a, dtype=np.float64), b)
    assert np.array_equal(pop["models_lmap"], g["models_lmap"]) and np.array_equal(pop["models_levid"], g["models_levid"])
    kw = dict(free_scale=True, ignore_model_err=True, dim_prior=True)
    for tag, nodes_only in (("nodes", True), ("full", False)):
        nb, res = fo.network_fit(m, me, mm, nodes, pop, g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy(),
                                 nodes_only=nodes_only, lprob_kwargs=kw)
        for a, b in zip(nb, _unragged(g, tag + "_neighbors")):
            assert np.array_equal(np.asarray(a, dtype=np.int64), b)
        for r, b in zip(res, _unragged(g, tag + "_lnprob")):
            assert np.array_equal(r[2], b, equal_nan=True)
        for r, b in zip(res, _unragged(g, tag + "_chi2")):
            assert np.array_equal(r[4], b, equal_nan=True)
