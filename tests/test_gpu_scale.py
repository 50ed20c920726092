"""Full-size behaviour of the fused path (C3 workload) through properties that do
not need a CPU run of the whole problem, plus oracle / float64 spot checks on sub-samples."""
import numpy as np
import pytest

import bench_data
from oracle import fz_oracle as fo

pytestmark = pytest.mark.gpu
LPROB = dict(free_scale=True, ignore_model_err=True, dim_prior=True)


@pytest.fixture(scope="module")
def c3():
    import frankenz_b200 as fz
    models, labels, depth = bench_data.c3_models()
    x, xe, xm, jtrue, mag = bench_data.c3_objects(98304 + 4096, models, depth, seed=11)
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(models), 0.05)
    bf = fz.BruteForce(models, np.zeros_like(models), np.ones_like(models))
    p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), labels, labe, label_dict=rdict, return_gof=True,
                                 verbose=False, save_fits=False, lprob_kwargs=LPROB)
    return dict(fz=fz, models=models, labels=labels, labe=labe, x=x, xe=xe, xm=xm, rdict=rdict, zgrid=zgrid, sig=sig,
                p=p, lm=lm, le=le, best=bf.best_idx.copy(), jtrue=jtrue, stats=bf._eng().stats())


def test_full_model_grid_properties(c3):
    p, lm, le = c3["p"], c3["lm"], c3["le"]
    assert p.shape == (len(c3["x"]), 701) and np.all(np.isfinite(p)) and np.all(p >= 0)
    assert np.max(np.abs(p.sum(axis=1) - 1.0)) < 1e-12            # bruteforce.py:370
    assert np.all(le >= lm) and np.all(le <= lm + np.log(len(c3["models"])) + 1e-9)   # logsumexp bounds
    assert c3["stats"]["pairs_fp32"] > 1.0 * len(c3["x"]) * len(c3["models"])     # the fp32 kernels did the work
    # bright objects recover the redshift of the model they were drawn from
    snr = np.sqrt(np.sum((c3["x"] / c3["xe"]) ** 2, axis=1))
    b = snr > 300
    zb = c3["labels"][c3["best"][b]]
    assert np.mean(np.abs(zb - c3["labels"][c3["jtrue"][b]]) < 0.02) > 0.9


def test_object_chunking_and_order_invariance(c3):
    """Objects are independent: any sub-batch, in any order, gives the same rows."""
    fz = c3["fz"]
    rs = np.random.RandomState(3)
    sel = rs.choice(len(c3["x"]), 3000, replace=False)
    bf = fz.BruteForce(c3["models"], np.zeros_like(c3["models"]), np.ones_like(c3["models"]))
    p, (lm, le) = bf.fit_predict(c3["x"][sel].copy(), c3["xe"][sel].copy(), c3["xm"][sel].copy(), c3["labels"],
                                 c3["labe"], label_dict=c3["rdict"], return_gof=True, verbose=False, save_fits=False,
                                 lprob_kwargs=LPROB)
    assert np.max(np.sum(np.abs(p - c3["p"][sel]), axis=1)) < 2e-6     # fp32 atomics order differs between launches
    assert np.allclose(lm, c3["lm"][sel], rtol=0, atol=1e-9) and np.allclose(le, c3["le"][sel], rtol=0, atol=2e-6)


def test_model_permutation_invariance(c3):
    """The reductions are associative: shuffling the models (and their labels) leaves PDFs unchanged."""
    fz = c3["fz"]
    rs = np.random.RandomState(4)
    perm = rs.permutation(len(c3["models"]))
    sel = slice(0, 2048)
    bf = fz.BruteForce(c3["models"][perm], np.zeros_like(c3["models"]), np.ones_like(c3["models"]))
    p, (lm, le) = bf.fit_predict(c3["x"][sel].copy(), c3["xe"][sel].copy(), c3["xm"][sel].copy(),
                                 c3["labels"][perm], c3["labe"], label_dict=c3["rdict"], return_gof=True,
                                 verbose=False, save_fits=False, lprob_kwargs=LPROB)
    assert np.max(np.sum(np.abs(p - c3["p"][sel]), axis=1)) < 2e-6
    assert np.allclose(lm, c3["lm"][sel], rtol=0, atol=1e-6) and np.allclose(le, c3["le"][sel], rtol=0, atol=2e-6)


def test_against_float64_path_and_oracle(c3):
    fz = c3["fz"]
    sel = np.arange(0, len(c3["x"]), 211)[:480]
    bf = fz.BruteForce(c3["models"], np.zeros_like(c3["models"]), np.ones_like(c3["models"]))
    p64, (lm64, le64) = bf.fit_predict(c3["x"][sel].copy(), c3["xe"][sel].copy(), c3["xm"][sel].copy(), c3["labels"],
                                       c3["labe"], label_dict=c3["rdict"], return_gof=True, verbose=False,
                                       save_fits=False, lprob_kwargs=dict(LPROB, precision="fp64"))
    assert np.max(np.sum(np.abs(c3["p"][sel] - p64), axis=1)) <= 1e-5            # north_star: PDFs 1e-5 L1
    assert np.all(np.abs(c3["lm"][sel] - lm64) <= 1e-5 * np.maximum(1, np.abs(lm64)))
    assert np.all(np.abs(c3["le"][sel] - le64) <= 1e-5 * np.maximum(1, np.abs(le64)))
    # oracle (numpy float64, reference arithmetic) on a handful of objects against all 199,950 models
    kd = fo.KernelDict(c3["zgrid"], c3["sig"])
    o = sel[:6]
    po, lmo, leo = fo.bruteforce_fit_predict(c3["models"], np.zeros_like(c3["models"]), np.ones_like(c3["models"]),
                                             c3["x"][o].copy(), c3["xe"][o].copy(), c3["xm"][o].copy(), c3["labels"],
                                             c3["labe"], label_dict=kd, **LPROB)
    assert np.max(np.sum(np.abs(c3["p"][o] - po), axis=1)) <= 1e-5
    assert np.max(np.sum(np.abs(p64[:6] - po), axis=1)) <= 1e-9
    assert np.all(np.abs(c3["lm"][o] - lmo) <= 1e-5 * np.maximum(1, np.abs(lmo)))
    assert np.all(np.abs(c3["le"][o] - leo) <= 1e-5 * np.maximum(1, np.abs(leo)))


def test_default_likelihood_c1_shape():
    """Config C1 shape (2k x 20k SDSS mock, default likelihood with model errors): fit + predict vs fused, and
    the fused fp32 path (FX1: per-band reciprocal variances) vs the float64 path."""
    import frankenz_b200 as fz
    m, me, mm, z, x, xe, xm, _ = bench_data.c1_dataset(20000, 2000)
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(m), 0.05)
    bf = fz.BruteForce(m, me, mm)
    bf.fit(x.copy(), xe.copy(), xm.copy(), verbose=False)
    p1, (lm1, le1) = bf.predict(z, labe, label_dict=rdict, return_gof=True, verbose=False)
    assert bf.fit_lnprob.shape == (2000, 20000) and np.max(np.abs(p1.sum(axis=1) - 1)) < 1e-12
    p2, (lm2, le2) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict, return_gof=True,
                                    verbose=False, save_fits=False)
    assert np.max(np.sum(np.abs(p1 - p2), axis=1)) <= 1e-5
    assert np.all(np.abs(lm1 - lm2) <= 1e-5 * np.maximum(1, np.abs(lm1)))
    assert np.all(np.abs(le1 - le2) <= 1e-5 * np.maximum(1, np.abs(le1)))
    kd = fo.KernelDict(zgrid, sig)
    po, lmo, leo = fo.bruteforce_fit_predict(m, me, mm, x[:8].copy(), xe[:8].copy(), xm[:8].copy(), z, labe,
                                             label_dict=kd)
    assert np.max(np.sum(np.abs(p1[:8] - po), axis=1)) <= 1e-9 and np.allclose(lm1[:8], lmo, rtol=1e-10)


def test_knn_fast_scan_matches_exact_kernel_and_oracle(monkeypatch):
    """fp32 scan + float64 re-rank (large training sets) against the all-float64 kernel and the oracle,
    including duplicated training rows (exact-distance ties) and queries that coincide with training rows."""
    import frankenz_b200 as fz
    from frankenz_b200._engine import Engine
    m, me, mm, z, x, xe, xm, _ = bench_data.c1_dataset(20000, 700)
    depth = np.load(bench_data.GOLDEN + "/sdss_cww_mock.npz")["depth_flux1sig"]
    kw = dict(skynoise=depth, zeropoints=10 ** (-0.4 * -23.9))
    feats = fo.knn_train_features(m, me, 3, feature_map="luptitude", fmap_kwargs=kw, rstate=np.random.RandomState(5))
    feats[1, 5000:5040] = feats[1, 100:140]          # duplicated rows -> exact ties
    q, _ = fo.luptitude(np.random.RandomState(6).normal(x, xe), xe, **kw)
    q[:10] = feats[0, 200:210].astype(np.float64)    # zero-distance hits in tree 0
    q[10:20] = feats[1, 100:110].astype(np.float64)  # zero-distance ties in tree 1
    eng = Engine(m, me, mm)
    eng.knn_build(feats)
    idx_fast, dist_fast = eng.knn_query(q, 25, p=2)
    monkeypatch.setenv("FZB_KNN_EXACT_ONLY", "1")
    idx_exact, dist_exact = eng.knn_query(q, 25, p=2)
    monkeypatch.delenv("FZB_KNN_EXACT_ONLY")
    assert np.array_equal(idx_fast, idx_exact)        # bit-exact indices, ties included
    assert np.array_equal(dist_fast, dist_exact)
    for i in (0, 3, 12, 15, 100, 699):
        oi, od = fo.knn_query_exact(feats, q[i], 25, 2)
        assert np.array_equal(idx_fast[i], oi) and np.allclose(dist_fast[i], od, rtol=1e-14, atol=0)
    # through the estimator: neighbours / fits identical between the two search kernels
    nn = fz.NearestNeighbors(m, me, mm, K=3, fmap_kwargs=kw, rstate=np.random.RandomState(5), verbose=False)
    nn.fit(x.copy(), xe.copy(), xm.copy(), k=25, eps=0, rstate=np.random.RandomState(6), verbose=False)
    nb, nnb, lp = nn.neighbors.copy(), nn.Nneighbors.copy(), nn.fit_lnprob.copy()
    monkeypatch.setenv("FZB_KNN_EXACT_ONLY", "1")
    nn.fit(x.copy(), xe.copy(), xm.copy(), k=25, eps=0, rstate=np.random.RandomState(6), verbose=False)
    assert np.array_equal(nb, nn.neighbors) and np.array_equal(nnb, nn.Nneighbors)
    assert np.array_equal(lp, nn.fit_lnprob)


@pytest.mark.parametrize("form", ["dot", "diff"])
def test_knn_filter_scan_adversarial_inputs(monkeypatch, form):
    """Threshold-filter scan (nested prefixes + select) against the all-float64 kernel: training rows SORTED by the
    first feature (a prefix of the caller's order is then the worst possible sample), several query tiles with a
    ragged tail, a far-away query, a NaN query, k = 1 / 25 / 100, both forms of the fp32 distance.  Indices and
    distances must be identical; the counters say how the fast path fared."""
    from frankenz_b200._engine import Engine
    monkeypatch.setenv("FZB_KNN_FORM", form)
    rs = np.random.RandomState(11)
    nm, nq, K = 60000, 2500, 2
    base = rs.normal(size=(nm, 5)) * np.array([1.0, 0.7, 0.5, 0.9, 1.3]) + np.array([21.0, 20.5, 20.0, 19.8, 19.5])
    feats = np.stack([base + rs.normal(size=base.shape) * 0.02 for _ in range(K)]).astype(np.float32)
    order = np.argsort(feats[0][:, 0])
    feats = np.ascontiguousarray(feats[:, order])               # sorted rows
    feats[1, 777] = feats[1, 12345]                             # an exact duplicate
    q = base[rs.choice(nm, nq)] + rs.normal(size=(nq, 5)) * 0.05
    q[5] = [35.0, 5.0, 20.0, 19.0, 50.0]                        # far from every row
    q[6] = np.nan
    q[7] = feats[1, 777].astype(np.float64)                     # zero-distance tie
    m = np.ones((nm, 5))
    eng = Engine(m, m, m)
    eng.knn_build(feats)
    for k in (1, 25, 100):
        idx_fast, dist_fast = eng.knn_query(q, k, p=2)
        st = eng.stats()
        monkeypatch.setenv("FZB_KNN_EXACT_ONLY", "1")
        idx_exact, dist_exact = eng.knn_query(q, k, p=2)
        monkeypatch.delenv("FZB_KNN_EXACT_ONLY")
        assert np.array_equal(idx_fast, idx_exact), (form, k, st)
        assert np.array_equal(dist_fast, dist_exact)
        # the fast path must have carried (nearly) all searches: the NaN query and the ties go to the float64 kernel, and
        # so does the far-away query in the tree that borrows the threshold of tree 0 (1.5 x its radius holds every row)
        assert st["knn_redo"] <= 8 and st["knn_overflow"] <= 2, st
        if form == "dot":
            assert st["knn_tc_err"] <= 2e-6, st                  # measured error of the fp32 values vs the bound
    for i in (0, 5, 7, 2499):
        oi, od = fo.knn_query_exact(feats, q[i], 100, 2)
        assert np.array_equal(idx_fast[i], oi) and np.allclose(dist_fast[i], od, rtol=1e-14, atol=0)


def test_knn_trees_that_are_not_alike(monkeypatch):
    """The filter scan borrows the threshold of tree 0 for the other trees (they are noise realisations of one training
    set).  Trees that are NOT alike break that shortcut, never the result: the pairs whose borrowed threshold holds too few
    (or too many) rows are re-done by the float64 kernel, and the handle switches to the staged search of every tree."""
    from frankenz_b200._engine import Engine
    rs = np.random.RandomState(5)
    base = rs.normal(size=(20000, 5)) + 20.0
    feats = np.stack([base, 20.0 + (base - 20.0) * 6.0, 20.0 + (base - 20.0) * 0.2]).astype(np.float32)   # radii x6, x0.2
    q = base[rs.choice(len(base), 1500)] + rs.normal(size=(1500, 5)) * 0.05
    ones = np.ones((len(base), 5))
    eng = Engine(ones, ones, ones)
    eng.knn_build(feats)
    monkeypatch.setenv("FZB_KNN_EXACT_ONLY", "1")
    idx_exact, dist_exact = eng.knn_query(q, 10, p=2)
    monkeypatch.delenv("FZB_KNN_EXACT_ONLY")
    idx1, dist1 = eng.knn_query(q, 10, p=2)
    st1 = eng.stats()
    idx2, dist2 = eng.knn_query(q, 10, p=2)
    st2 = eng.stats()
    assert np.array_equal(idx1, idx_exact) and np.array_equal(dist1, dist_exact)
    assert np.array_equal(idx2, idx_exact) and np.array_equal(dist2, dist_exact)
    assert st1["knn_redo"] > 0.3 * len(q) and st2["knn_redo"] <= 8, (st1["knn_redo"], st2["knn_redo"])


def test_default_likelihood_fused_single_pass(monkeypatch):
    """The reference's default likelihood (fixed scale, model errors, dim_prior) on a batch large enough for the fused single
    pass of the packed sweep (>= 49,152 objects, four per thread): against the two-pass sweep on every object and against the
    float64 kernels on a sub-sample; the counters say that the single pass carried most objects."""
    import frankenz_b200 as fz
    tr, tre, trm, ztr, x, xe, xm = bench_data.c5_dataset(98304, 120000, seed=77)     # dense enough for the coarse pre-pass;
    # the host pipeline splits 120,000 objects into chunks of 61,440 (four objects per thread: fused) + 32,768 + 25,792
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(tr), 0.05)
    bf = fz.BruteForce(tr, tre, trm)
    res = {}
    for fused in (True, False):
        if not fused:
            monkeypatch.setenv("FZB_NO_FUSE", "1")
        p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), ztr, labe, label_dict=rdict, return_gof=True,
                                     verbose=False, save_fits=False)
        monkeypatch.delenv("FZB_NO_FUSE", raising=False)
        res[fused] = (p, lm, le, bf.best_idx.copy(), bf._eng().stats())
    (p1, lm1, le1, b1, st1), (p2, lm2, le2, b2, st2) = res[True], res[False]
    assert st1["sweep_kind"] == 1 and st2["objects_fused"] == 0 and st1["objects_fused"] > 0.1 * len(x), (st1, st2)
    assert st1["pairs_pass2"] < 0.95 * st2["pairs_pass2"], (st1["pairs_pass2"], st2["pairs_pass2"])
    assert np.array_equal(b1, b2) and np.array_equal(lm1, lm2)
    ok = np.isfinite(lm1)
    assert np.max(np.abs(le1[ok] - le2[ok])) <= 2e-6 and np.max(np.sum(np.abs(p1[ok] - p2[ok]), axis=1)) <= 2e-6
    sub = np.arange(0, len(x), 60)
    p64, (lm64, le64) = bf.fit_predict(x[sub].copy(), xe[sub].copy(), xm[sub].copy(), ztr, labe, label_dict=rdict,
                                       return_gof=True, verbose=False, save_fits=False, lprob_kwargs=dict(precision="fp64"))
    good = np.isfinite(lm64)
    assert np.max(np.sum(np.abs(p1[sub][good] - p64[good]), axis=1)) <= 1e-5
    assert np.all(np.abs(lm1[sub][good] - lm64[good]) <= 1e-5 * np.maximum(1, np.abs(lm64[good])))
    assert np.all(np.abs(le1[sub][good] - le64[good]) <= 1e-5 * np.maximum(1, np.abs(le64[good])))


def test_float64_sweep_route_matches_generic(c3):
    """Fixed-scale fits of the C3 objects: most best-fit chi2 are far above the fp32 bound, so the objects take the
    register-tiled float64 sweep (k_sweep64).  Its PDFs must agree with the reference-order float64 kernel."""
    fz = c3["fz"]
    sel = np.arange(0, 60000, 7)[:6000]
    for kw in (dict(free_scale=False, ignore_model_err=True), dict()):
        bf = fz.BruteForce(c3["models"], np.zeros_like(c3["models"]), np.ones_like(c3["models"]))
        p, (lm, le) = bf.fit_predict(c3["x"][sel].copy(), c3["xe"][sel].copy(), c3["xm"][sel].copy(), c3["labels"],
                                     c3["labe"], label_dict=c3["rdict"], return_gof=True, verbose=False,
                                     save_fits=False, lprob_kwargs=kw)
        st = bf._eng().stats()
        assert st["objects_fp64"] > 500 and st["pairs_fp64"] > 0, st      # the float64 sweep really ran
        sub = np.arange(0, len(sel), 13)
        p64, (lm64, le64) = bf.fit_predict(c3["x"][sel][sub].copy(), c3["xe"][sel][sub].copy(),
                                           c3["xm"][sel][sub].copy(), c3["labels"], c3["labe"],
                                           label_dict=c3["rdict"], return_gof=True, verbose=False, save_fits=False,
                                           lprob_kwargs=dict(kw, precision="fp64"))
        assert np.max(np.sum(np.abs(p[sub] - p64), axis=1)) <= 1e-5
        assert np.all(np.abs(lm[sub] - lm64) <= 1e-5 * np.maximum(1, np.abs(lm64)))
        assert np.all(np.abs(le[sub] - le64) <= 1e-5 * np.maximum(1, np.abs(le64)))


def test_model_sharded_fast_path_matches_unsharded(c3):
    """Model-sharded passes on the fp32 sweep kernels: three model shards of the C3 grid on one GPU, merged with the
    arithmetic of frankenz_b200.distributed (collectives replaced by their definitions), against the unsharded run."""
    import ctypes as C
    import torch
    from frankenz_b200 import _lib
    from frankenz_b200._engine import Engine, make_config
    from frankenz_b200.distributed import shard_bounds
    n = 4096
    x = [torch.from_numpy(np.ascontiguousarray(a[:n])).cuda() for a in (c3["x"], c3["xe"], c3["xm"])]
    cfg = make_config(LPROB, None)
    nm = len(c3["models"])
    engs, parts = [], []
    for r in range(3):
        lo, hi = shard_bounds(nm, 3, r)
        e = Engine(c3["models"][lo:hi], np.zeros((hi - lo, 5)), np.ones((hi - lo, 5)))
        e.set_kde(c3["labels"][lo:hi], c3["labe"][lo:hi], label_dict=c3["rdict"])
        pm = torch.empty(n, dtype=torch.float64).cuda()
        ps = torch.empty(n, dtype=torch.float64).cuda()
        pb = torch.empty(n, dtype=torch.int64).cuda()
        _lib.check(e.lib.fzb_shard_pass1_dev(e.h, x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), n, C.byref(cfg),
                                             pm.data_ptr(), ps.data_ptr(), pb.data_ptr()))
        assert e.stats()["pairs_fp32"] > 0          # the sweep kernels, not the generic path
        engs.append(e)
        parts.append((pm, ps, pb + lo))
    gmax = torch.stack([p[0] for p in parts]).max(dim=0).values
    s = sum(ps * torch.exp(pm - gmax) for pm, ps, _ in parts)
    levid = gmax + torch.log(s)
    assert np.allclose(gmax.cpu().numpy(), c3["lm"][:n], rtol=0, atol=1e-6)
    assert np.allclose(levid.cpu().numpy(), c3["le"][:n], rtol=0, atol=3e-6)
    tot = torch.zeros((n, 701), dtype=torch.float64).cuda()
    torch.cuda.synchronize()          # the library runs on its own stream: torch's results must be complete before it reads them
    for e in engs:
        part = torch.empty((n, 701), dtype=torch.float64).cuda()
        _lib.check(e.lib.fzb_shard_pass2_dev(e.h, x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), n, C.byref(cfg),
                                             gmax.data_ptr(), levid.data_ptr(), part.data_ptr()))
        assert e.stats()["pairs_fp32"] > 0
        tot += part
    p = (tot / tot.sum(dim=1, keepdim=True)).cpu().numpy()
    assert np.max(np.sum(np.abs(p - c3["p"][:n]), axis=1)) <= 3e-6


@pytest.mark.parametrize("kw", [dict(), dict(free_scale=True, ignore_model_err=True),
                                dict(free_scale=False, ignore_model_err=True, dim_prior=False),
                                dict(free_scale=True, ignore_model_err=True, dim_prior=False)])
def test_model_masks_on_the_sweep_path(kw):
    """Config C2 flavour: training rows with 5 % band dropouts (model masks) and objects with dropouts and NaNs.
    The packed sweep handles binary model masks (pair dimensionality = popc(object bits & model bits)); results
    must agree with the float64 reference-order kernel and with the oracle."""
    import frankenz_b200 as fz
    m, me, mm, z, x, xe, xm, _ = bench_data.c1_dataset(20000, 3000)
    rs = np.random.RandomState(12)
    mm = (rs.uniform(size=m.shape) > 0.05).astype(float)
    mm[:, 2] = 1.0                                   # keep one band everywhere so that no pair is empty
    xm = (rs.uniform(size=x.shape) > 0.10).astype(float)
    xm[:, 2] = 1.0
    x = x.copy()
    x[rs.uniform(size=x.shape) < 0.005] = np.nan
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = 0.01 * (1 + z)                            # several kernel widths
    bf = fz.BruteForce(m, me, mm)
    p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict, return_gof=True,
                                 verbose=False, save_fits=False, lprob_kwargs=kw)
    st = bf._eng().stats()
    assert st["pairs_fp32"] >= len(x) * len(m), st                        # the sweep kernels ran
    sub = np.arange(0, len(x), 5)
    p64, (lm64, le64) = bf.fit_predict(x[sub].copy(), xe[sub].copy(), xm[sub].copy(), z, labe, label_dict=rdict,
                                       return_gof=True, verbose=False, save_fits=False,
                                       lprob_kwargs=dict(kw, precision="fp64"))
    ok = np.isfinite(p64).all(axis=1)
    assert ok.mean() > 0.5       # free scale + dim_prior: a one-band pair has chi2 = 0, a = 0 -> NaN row in the reference too
    assert np.array_equal(np.isfinite(p[sub]).all(axis=1), ok)
    assert np.max(np.sum(np.abs(p[sub][ok] - p64[ok]), axis=1)) <= 1e-5
    assert np.all(np.abs(lm[sub][ok] - lm64[ok]) <= 1e-5 * np.maximum(1, np.abs(lm64[ok])))
    assert np.all(np.abs(le[sub][ok] - le64[ok]) <= 1e-5 * np.maximum(1, np.abs(le64[ok])))
    kd = fo.KernelDict(zgrid, sig)
    o = np.arange(6)
    po, lmo, leo = fo.bruteforce_fit_predict(m, me, mm, x[o].copy(), xe[o].copy(), xm[o].copy(), z, labe,
                                             label_dict=kd, **kw)
    fin = np.isfinite(po).all(axis=1)
    assert np.max(np.sum(np.abs(p[o][fin] - po[fin]), axis=1)) <= 1e-5
    assert np.all(np.abs(lm[o][fin] - lmo[fin]) <= 1e-5 * np.maximum(1, np.abs(lmo[fin])))


def test_model_sharded_driver_single_rank(c3):
    """frankenz_b200.distributed.ModelShardedBruteForce with one rank (no process group): device tensors in, device
    tensors out, the shard kept between calls; equals the unsharded run."""
    import torch
    from frankenz_b200.distributed import ModelShardedBruteForce
    n = 3000
    sb = ModelShardedBruteForce(c3["models"], np.zeros_like(c3["models"]), np.ones_like(c3["models"]))
    tx = [torch.from_numpy(np.ascontiguousarray(a[:n])).cuda() for a in (c3["x"], c3["xe"], c3["xm"])]
    for rep in range(2):
        p, (lm, le), best = sb.fit_predict(tx[0], tx[1], tx[2], c3["labels"], c3["labe"], label_dict=c3["rdict"],
                                           lprob_kwargs=LPROB, return_best=True, as_torch=True)
        assert np.max(np.sum(np.abs(p.cpu().numpy() - c3["p"][:n]), axis=1)) <= 3e-6
        assert np.allclose(lm.cpu().numpy(), c3["lm"][:n], rtol=0, atol=1e-6)
        assert np.allclose(le.cpu().numpy(), c3["le"][:n], rtol=0, atol=3e-6)
        assert np.array_equal(best.cpu().numpy(), c3["best"][:n])
    sb.close()


def test_pinned_output_pool(c3, monkeypatch):
    """Page-locked output arrays (the multi-rank default): same results, buffers recycled once the arrays are gone."""
    from frankenz_b200 import _engine
    pool = _engine.PinnedPool(max_bytes=1 << 30)
    pool.MIN_BYTES = 1 << 20
    monkeypatch.setattr(_engine, "_pinned_pool", pool)
    fz = c3["fz"]
    n = 3000
    bf = fz.BruteForce(c3["models"], np.zeros_like(c3["models"]), np.ones_like(c3["models"]))
    for rep in range(2):
        p = bf.fit_predict(c3["x"][:n].copy(), c3["xe"][:n].copy(), c3["xm"][:n].copy(), c3["labels"], c3["labe"],
                           label_dict=c3["rdict"], verbose=False, save_fits=False, lprob_kwargs=LPROB)
        assert pool.total == n * 701 * 8 and not pool.free            # one pinned buffer, in use
        assert np.max(np.sum(np.abs(p - c3["p"][:n]), axis=1)) <= 2e-6
        q = p[5:10].copy()
        del p
        assert len(pool.free) == 1                                     # back in the pool, reused by the next call
    assert np.all(np.isfinite(q))
