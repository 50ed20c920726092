"""CUDA path (through the C ABI / the reference-shaped Python API) against the reference's golden
vectors and the oracle.  Tolerances are the north_star's: lnL / chi2 / scale 1e-5 relative with
identical inf/nan positions, PDFs 1e-5 L1, lmap / levid 1e-5*max(1,|x|), kNN indices exact."""
import numpy as np
import pytest

from conftest import golden, max_rel, same_special
from oracle import fz_oracle as fo

pytestmark = pytest.mark.gpu

COMBOS = [(fs, ime, dp) for fs in (False, True) for ime in (False, True) for dp in (False, True)]
REL = 1e-5          # north_star tolerance
REL64 = 1e-9        # what the float64 kernels are expected to deliver


def tag(fs, ime, dp):
    return "fs%d_ime%d_dp%d" % (fs, ime, dp)


def l1(p, q):
    return float(np.max(np.sum(np.abs(p - q), axis=1)))


def close_gof(a, b, tol=REL):
    a, b = np.asarray(a), np.asarray(b)
    assert same_special(a, b)
    ok = np.isfinite(b)
    assert np.all(np.abs(a[ok] - b[ok]) <= tol * np.maximum(1.0, np.abs(b[ok])))


@pytest.fixture(scope="module")
def fz():
    import frankenz_b200
    return frankenz_b200


@pytest.mark.parametrize("fname", ["loglike_combos.npz", "loglike_degenerate.npz"])
@pytest.mark.parametrize("fs,ime,dp", COMBOS)
def test_fit_matrix_vs_golden(fz, fname, fs, ime, dp):
    g = golden(fname)
    bf = fz.BruteForce(g["models"], g["models_err"], g["models_mask"])
    x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
    kw = dict(free_scale=fs, ignore_model_err=ime, dim_prior=dp, ltol=1e-4)
    if fs:
        kw["return_scale"] = True
    bf.fit(x, xe, xm, lprob_kwargs=kw, track_scale=fs, verbose=False)
    t = tag(fs, ime, dp)
    for name, arr in (("lnl", bf.fit_lnlike), ("chi2", bf.fit_chi2)):
        ref = g[t + "_" + name]
        assert same_special(arr, ref), (t, name)
        assert max_rel(arr, ref) <= REL64, (t, name, max_rel(arr, ref))
    assert np.array_equal(bf.fit_Ndim, g[t + "_ndim"].astype(np.int64))
    assert np.array_equal(bf.fit_lnprob, bf.fit_lnlike, equal_nan=True) and np.all(bf.fit_lnprior == 0)
    if fs:
        for name, arr in (("scale", bf.fit_scale), ("scale_err", bf.fit_scale_err)):
            ref = g[t + "_" + name]
            assert same_special(arr, ref), (t, name)
            assert max_rel(arr, ref) <= REL64, (t, name)
    else:
        assert np.all(bf.fit_scale == 1) and np.all(bf.fit_scale_err == 0)
    if "cleaned_data" in g.files:   # drop-in mutates the caller's arrays like pdf.py:310-311
        assert np.array_equal(x, g["cleaned_data"]) and np.array_equal(xe, g["cleaned_err"])
        assert np.array_equal(xm, g["cleaned_mask"])


def test_pdf_loglike_single_object(fz):
    g = golden("loglike_combos.npz")
    r = fz.pdf.loglike(g["data"][7].copy(), g["data_err"][7].copy(), g["data_mask"][7].copy(), g["models"],
                       g["models_err"], g["models_mask"], free_scale=True, return_scale=True)
    t = tag(True, False, True)
    assert len(r) == 5
    for got, name in zip(r, ("lnl", "ndim", "chi2", "scale", "scale_err")):
        assert same_special(got, g[t + "_" + name][7]) and max_rel(got, g[t + "_" + name][7]) <= REL64
    lp = np.linspace(-2, 0, len(g["models"]))
    r = fz.pdf.logprob(g["data"][7].copy(), g["data_err"][7].copy(), g["data_mask"][7].copy(), g["models"],
                       g["models_err"], g["models_mask"], lnprior=lp)
    assert len(r) == 5 and np.array_equal(r[0], lp)
    ok = np.isfinite(r[1])
    assert np.allclose(r[2][ok], r[1][ok] + lp[ok], rtol=1e-14)


def test_fs1_ltol_iteration_rule(fz):
    g = golden("fs1_ltol.npz")
    bf = fz.BruteForce(g["models"], g["models_err"], g["models_mask"])
    for ltol in (1e-2, 1e-4, 1e-7):
        bf.fit(g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy(), verbose=False,
               lprob_kwargs=dict(free_scale=True, dim_prior=False, ltol=ltol))
        ref = g["lnl_ltol%g" % ltol]
        # the stopping rule is object-wide: a wrong iteration count shows up at ~ltol, far above 1e-9
        assert max_rel(bf.fit_lnlike, ref) <= REL64, (ltol, max_rel(bf.fit_lnlike, ref))


def _dict(fz):
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    return zgrid, fz.pdf.PDFDict(zgrid, np.linspace(0.005, 2, 500))


def test_bruteforce_fit_then_predict(fz):
    g = golden("bruteforce_c1small.npz")
    zgrid, rdict = _dict(fz)
    bf = fz.BruteForce(g["models"], g["models_err"], g["models_mask"])
    bf.fit(g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy(), verbose=False)
    assert max_rel(bf.fit_lnprob, g["fit_lnprob"]) <= REL64 and max_rel(bf.fit_chi2, g["fit_chi2"]) <= REL64
    assert np.array_equal(bf.fit_Ndim, g["fit_Ndim"])
    lab, labe = g["labels"], g["label_errs"]
    p, (lm, le) = bf.predict(lab, labe, label_dict=rdict, return_gof=True, verbose=False)
    assert l1(p, g["pdf_dict"]) <= 1e-9
    close_gof(lm, g["lmap"], 1e-9)
    close_gof(le, g["levid"], 1e-9)
    assert l1(bf.predict(lab, labe, label_grid=zgrid, verbose=False), g["pdf_grid"]) <= 1e-9
    assert l1(bf.predict(lab, g["label_errs2"], label_dict=rdict, verbose=False), g["pdf_dict_mixed"]) <= 1e-9
    assert l1(bf.predict(lab, g["label_errs2"], label_grid=zgrid, verbose=False), g["pdf_grid_mixed"]) <= 1e-9
    p = bf.predict(lab, labe, label_dict=rdict, verbose=False, kde_kwargs=dict(wt_thresh=None, cdf_thresh=None))
    assert l1(p, g["pdf_dict_nothresh"]) <= 1e-9
    p = bf.predict(lab, labe, label_dict=rdict, verbose=False, kde_kwargs=dict(wt_thresh=None, cdf_thresh=2e-4))
    assert l1(p, g["pdf_dict_cdf"]) <= 1e-9
    p = bf.predict(lab, labe, label_grid=zgrid, verbose=False, kde_kwargs=dict(wt_thresh=None, cdf_thresh=2e-4))
    assert l1(p, g["pdf_grid_cdf"]) <= 1e-9
    # user-supplied log-weights (demo 2 cell 71 passes fit_lnlike)
    p = bf.predict(lab, labe, label_dict=rdict, logwt=bf.fit_lnlike, verbose=False)
    assert l1(p, g["pdf_dict"]) <= 1e-9
    # generator twin
    rows = list(bf._predict(lab, labe, label_dict=rdict))
    assert len(rows) == len(g["data"]) and np.allclose(rows[3][0], g["pdf_dict"][3], atol=1e-12)


@pytest.mark.parametrize("precision", ["fp64", "auto"])
@pytest.mark.parametrize("fs,ime,dp", COMBOS)
def test_bruteforce_fused_fit_predict(fz, fs, ime, dp, precision):
    g = golden("bruteforce_c1small.npz")
    _, rdict = _dict(fz)
    bf = fz.BruteForce(g["models"], g["models_err"], g["models_mask"])
    p, (lm, le) = bf.fit_predict(g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy(), g["labels"],
                                 g["label_errs"], label_dict=rdict, return_gof=True, verbose=False, save_fits=False,
                                 lprob_kwargs=dict(free_scale=fs, ignore_model_err=ime, dim_prior=dp,
                                                   precision=precision))
    t = tag(fs, ime, dp)
    tol = 1e-9 if precision == "fp64" else REL
    assert bf.fit_lnprob is None
    assert l1(p, g[t + "_pdf"]) <= tol, l1(p, g[t + "_pdf"])
    close_gof(lm, g[t + "_lmap"], tol)
    close_gof(le, g[t + "_levid"], tol)


def test_bruteforce_save_fits_and_errors(fz):
    g = golden("bruteforce_c1small.npz")
    zgrid, rdict = _dict(fz)
    bf = fz.BruteForce(g["models"], g["models_err"], g["models_mask"])
    p = bf.fit_predict(g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy(), g["labels"], g["label_errs"],
                       label_dict=rdict, verbose=False, save_fits=True)
    assert l1(p, g["pdf_dict"]) <= 1e-9 and max_rel(bf.fit_lnprob, g["fit_lnprob"]) <= REL64
    with pytest.raises(ValueError):
        bf.predict(g["labels"], g["label_errs"], verbose=False)
    with pytest.raises(ValueError):
        fz.BruteForce(g["models"], g["models_err"], g["models_mask"]).predict(g["labels"], g["label_errs"],
                                                                              label_dict=rdict, verbose=False)
    with pytest.raises(NotImplementedError):
        bf.fit(g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy(), lprob_func=lambda *a: None,
               verbose=False)


def test_kde_edges(fz):
    g = golden("kde_edges.npz")
    zgrid, rdict = _dict(fz)
    n = len(g["labels"])
    bf = fz.BruteForce(np.ones((n, 5)), np.ones((n, 5)), np.ones((n, 5)))
    p, (lm, le) = bf.predict(g["labels"], g["label_errs"], label_dict=rdict, logwt=g["logwt"], return_gof=True,
                             verbose=False)
    assert same_special(p, g["pdf_dict"])
    assert np.nanmax(np.abs(p - g["pdf_dict"])) <= 1e-12
    close_gof(lm, g["lmap"], 1e-12)
    close_gof(le, g["levid"], 1e-12)
    p = bf.predict(g["labels"], g["label_errs_grid"], label_grid=zgrid, logwt=g["logwt"], verbose=False)
    assert same_special(p, g["pdf_grid"]) and np.nanmax(np.abs(p - g["pdf_grid"])) <= 1e-12


@pytest.mark.parametrize("name,K,k,fmap", [("a", 5, 20, "luptitude"), ("b", 1, 1, "luptitude"),
                                           ("c", 8, 7, "identity"), ("d", 3, 25, "magnitude")])
def test_knn_exact(fz, name, K, k, fmap):
    g = golden("knn_exact.npz")
    zgrid, rdict = _dict(fz)
    m, me, mm = g["models"], g["models_err"], g["models_mask"]
    x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
    kw = {}
    if fmap == "luptitude":
        kw = dict(skynoise=g["skynoise"], zeropoints=float(g["zeropoints"]))
    elif fmap == "magnitude":
        kw = dict(zeropoints=float(g["zeropoints"]))
        m, me = np.abs(m) + 5 * me, me * 1e-3
        x, xe = np.abs(x) + 5 * xe, xe * 1e-3
    nn = fz.NearestNeighbors(m, me, mm, K=K, feature_map=fmap, fmap_kwargs=kw, rstate=np.random.RandomState(1),
                             verbose=False)
    assert np.array_equal(nn.features, g[name + "_feats"])
    p, (lm, le) = nn.fit_predict(x, xe, xm, g["labels"], g["label_errs"], label_dict=rdict, k=k, eps=0.0,
                                 rstate=np.random.RandomState(2), return_gof=True, verbose=False)
    assert np.array_equal(nn.Nneighbors, g[name + "_Nneighbors"])
    assert np.array_equal(nn.neighbors, g[name + "_neighbors"])          # bit-exact indices
    assert same_special(nn.fit_lnprob, g[name + "_lnprob"]) and max_rel(nn.fit_lnprob, g[name + "_lnprob"]) <= REL64
    assert same_special(nn.fit_chi2, g[name + "_chi2"]) and max_rel(nn.fit_chi2, g[name + "_chi2"]) <= REL64
    assert l1(p, g[name + "_pdf"]) <= 1e-9
    close_gof(lm, g[name + "_lmap"], 1e-9)
    close_gof(le, g[name + "_levid"], 1e-9)
    p2 = nn.predict(g["labels"], g["label_errs"], label_grid=zgrid, verbose=False)
    assert l1(p2, g[name + "_pdf_grid"]) <= 1e-9


def test_knn_query_matches_oracle(fz):
    g = golden("knn_exact.npz")
    from frankenz_b200._engine import Engine
    eng = Engine(g["models"], g["models_err"], g["models_mask"])
    feats = g["a_feats"]
    eng.knn_build(feats)
    rs = np.random.RandomState(3)
    q = feats[0][rs.choice(len(feats[0]), 16)].astype(np.float64) + rs.normal(size=(16, 5)) * 0.05
    for p in (2, 1, np.inf):
        idx, dist = eng.knn_query(q, 9, p=p)
        for i in range(len(q)):
            oi, od = fo.knn_query_exact(feats, q[i], 9, p)
            assert np.array_equal(idx[i], oi) and np.allclose(dist[i], od, rtol=1e-14)


@pytest.mark.parametrize("precision", ["fp64", "auto"])
def test_model_sharded_single_rank(fz, precision):
    """fzb_shard_pass1/2_dev + the merge arithmetic (world size 1 degenerates to identity collectives)."""
    from frankenz_b200.distributed import fit_predict_model_sharded
    g = golden("bruteforce_c1small.npz")
    _, rdict = _dict(fz)
    p, (lm, le), best = fit_predict_model_sharded(g["models"], g["models_err"], g["models_mask"], g["data"].copy(),
                                                  g["data_err"].copy(), g["data_mask"].copy(), g["labels"],
                                                  g["label_errs"], label_dict=rdict, return_best=True,
                                                  lprob_kwargs=dict(precision=precision))
    tol = 1e-9 if precision == "fp64" else REL
    assert l1(p, g["pdf_dict"]) <= (1e-6 if precision == "fp64" else REL)   # PDF partials cross the collective in fp32
    close_gof(lm, g["lmap"], tol)
    close_gof(le, g["levid"], tol)
    if precision == "fp64":
        assert np.array_equal(best, g["fit_lnprob"].argmax(axis=1))


@pytest.mark.parametrize("precision", ["fp64", "auto"])
def test_model_sharded_two_shards_one_gpu(fz, precision):
    """Two model shards evaluated one after the other on one GPU, merged with the same arithmetic the
    NCCL path uses (the collectives replaced by their definitions)."""
    import ctypes as C
    import torch
    from frankenz_b200 import _lib
    from frankenz_b200._engine import Engine, make_config
    g = golden("bruteforce_c1small.npz")
    _, rdict = _dict(fz)
    m, me, mm = g["models"], g["models_err"], g["models_mask"]
    x = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (g["data"], g["data_err"], g["data_mask"])]
    no, cut = len(g["data"]), 500
    cfg = make_config(dict(precision=precision), None)
    tol = 1e-9 if precision == "fp64" else REL
    engs, parts = [], []
    for lo, hi in ((0, cut), (cut, len(m))):
        e = Engine(m[lo:hi], me[lo:hi], mm[lo:hi])
        e.set_kde(g["labels"][lo:hi], g["label_errs"][lo:hi], label_dict=rdict)
        pm, ps = torch.empty(no, dtype=torch.float64).cuda(), torch.empty(no, dtype=torch.float64).cuda()
        pb = torch.empty(no, dtype=torch.int64).cuda()
        _lib.check(e.lib.fzb_shard_pass1_dev(e.h, x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), no, C.byref(cfg),
                                             pm.data_ptr(), ps.data_ptr(), pb.data_ptr()))
        engs.append(e)
        parts.append((pm, ps, pb))
    gmax = torch.maximum(parts[0][0], parts[1][0])
    s = sum(ps * torch.exp(pm - gmax) for pm, ps, _ in parts)
    levid = gmax + torch.log(s)
    close_gof(gmax.cpu().numpy(), g["lmap"], tol)
    close_gof(levid.cpu().numpy(), g["levid"], tol)
    tot = torch.zeros((no, rdict.Ngrid), dtype=torch.float64).cuda()
    for e in engs:
        part = torch.empty((no, rdict.Ngrid), dtype=torch.float64).cuda()
        _lib.check(e.lib.fzb_shard_pass2_dev(e.h, x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), no, C.byref(cfg),
                                             gmax.data_ptr(), levid.data_ptr(), part.data_ptr()))
        tot += part
    p = (tot / tot.sum(dim=1, keepdim=True)).cpu().numpy()
    assert l1(p, g["pdf_dict"]) <= tol


def test_object_conditioned_prior_table(fz):
    """lnprior[i, j] = table[bin_i, j] (the built-in form of demo 2's Python lprob_bpz)."""
    g = golden("bruteforce_c1small.npz")
    _, rdict = _dict(fz)
    m, me, mm = g["models"], g["models_err"], g["models_mask"]
    x, xe, xm = g["data"], g["data_err"], g["data_mask"]
    rs = np.random.RandomState(8)
    table = rs.normal(size=(4, len(m))) * 2.0
    table[2, ::7] = -np.inf                       # some models excluded for bin 2
    bins = rs.randint(0, 4, size=len(x))
    bf = fz.BruteForce(m, me, mm)
    kw = dict(lnprior=table, lnprior_bin=bins)
    bf.fit(x.copy(), xe.copy(), xm.copy(), lprob_kwargs=kw, verbose=False)
    kd = fo.KernelDict(np.arange(0, 7 + 1e-5, 0.01), np.linspace(0.005, 2, 500))
    for i in (0, 5, 17, 31):
        r = fo.logprob(x[i].copy(), xe[i].copy(), xm[i].copy(), m, me, mm, lnprior=table[bins[i]])
        assert np.array_equal(bf.fit_lnprior[i], table[bins[i]])
        assert same_special(bf.fit_lnprob[i], r[2]) and max_rel(bf.fit_lnprob[i], r[2]) <= REL64
        assert max_rel(bf.fit_lnlike[i], r[1]) <= REL64
    p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), g["labels"], g["label_errs"], label_dict=rdict,
                                 lprob_kwargs=kw, return_gof=True, verbose=False, save_fits=False)
    for i in (0, 5, 17, 31):
        r = fo.logprob(x[i].copy(), xe[i].copy(), xm[i].copy(), m, me, mm, lnprior=table[bins[i]])
        yi, si = kd.fit(g["labels"], g["label_errs"])
        po, lmo, leo = fo.weights_to_pdf(r[2], g["labels"], g["label_errs"], yi, si, label_dict=kd)
        assert np.sum(np.abs(p[i] - po)) <= 1e-9 and abs(lm[i] - lmo) <= 1e-9 * max(1, abs(lmo))
        assert abs(le[i] - leo) <= 1e-9 * max(1, abs(leo))
    # the bins are consumed by the call: a second call without them uses no prior
    bf.fit(x.copy(), xe.copy(), xm.copy(), verbose=False)
    assert np.all(bf.fit_lnprior == 0)
    with pytest.raises(ValueError):
        bf.fit(x.copy(), xe.copy(), xm.copy(), lprob_kwargs=dict(lnprior=table), verbose=False)
