"""Tensor-core sweep (csrc/fzb_sweep_tc.cuh) against the CPU oracle (oracle/fz_oracle.py, the restatement of
frankenz/bruteforce.py:505-631 + pdf.py:103-235) on the branches the full-size C3 tests do not reach: partial / odd
model tiles, a per-model prior, the log-domain form (dim_prior=False; four bands), objects with masked or non-finite
bands and very bright objects (both leave the fp32 path), and batch / split invariance.  Bounds are the north_star's:
PDFs 1e-5 L1, lmap / levid 1e-5 max(1, |x|)."""
import numpy as np
import pytest

import bench_data
from oracle import fz_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fz():
    import frankenz_b200
    return frankenz_b200


def _case(nm, no, nf=5, seed=5):
    models, labels, depth = bench_data.c3_models()
    rs = np.random.RandomState(seed)
    pick = np.sort(rs.choice(len(models), nm, replace=False))
    m = np.ascontiguousarray(models[pick][:, :nf])
    x, xe, xm, _, _ = bench_data.c3_objects(no, models[pick], depth, seed=seed + 1)
    return m, labels[pick], np.ascontiguousarray(x[:, :nf]), np.ascontiguousarray(xe[:, :nf]), np.ones((no, nf))


def _run(fz, m, lab, x, xe, xm, lprob, lnprior=None):
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(m), 0.05)
    bf = fz.BruteForce(m, np.zeros_like(m), np.ones_like(m))
    kw = dict(lprob)
    if lnprior is not None:
        kw["lnprior"] = lnprior
    p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), lab, labe, label_dict=rdict, return_gof=True,
                                 verbose=False, save_fits=False, lprob_kwargs=kw)
    st = bf._eng().stats()
    kd = fo.KernelDict(zgrid, sig)
    with np.errstate(all="ignore"):
        po, lmo, leo = fo.bruteforce_fit_predict(m, np.zeros_like(m), np.ones_like(m), x.copy(), xe.copy(), xm.copy(),
                                                 lab, labe, label_dict=kd, lnprior=lnprior, **lprob)
    return p, lm, le, po, lmo, leo, st


def _check(p, lm, le, po, lmo, leo):
    ok = np.isfinite(lmo)
    assert np.array_equal(np.isfinite(lm), ok)
    assert np.max(np.sum(np.abs(p[ok] - po[ok]), axis=1)) <= 1e-5
    assert np.all(np.abs(lm[ok] - lmo[ok]) <= 1e-5 * np.maximum(1, np.abs(lmo[ok])))
    assert np.all(np.abs(le[ok] - leo[ok]) <= 1e-5 * np.maximum(1, np.abs(leo[ok])))


FS = dict(free_scale=True, ignore_model_err=True, dim_prior=True)


@pytest.mark.parametrize("nm", [5003, 4098, 777, 37, 9])
def test_partial_and_odd_model_tiles(fz, nm):
    """Model counts that end in an odd pair / a partial 8-model sub-batch / a partial 32-model chunk."""
    m, lab, x, xe, xm = _case(nm, 300 if nm > 100 else 7)
    p, lm, le, po, lmo, leo, st = _run(fz, m, lab, x, xe, xm, FS)
    assert st["sweep_kind"] == 3 and st["pairs_fp32"] > 0           # linear-domain tensor-core sweep
    _check(p, lm, le, po, lmo, leo)


def test_per_model_prior(fz):
    m, lab, x, xe, xm = _case(6001, 256)
    lnprior = np.random.RandomState(2).uniform(-6, 0, len(m))
    p, lm, le, po, lmo, leo, st = _run(fz, m, lab, x, xe, xm, FS, lnprior=lnprior)
    assert st["sweep_kind"] == 3
    _check(p, lm, le, po, lmo, leo)


def test_log_domain_forms(fz):
    """dim_prior=False (plain exp(-chi2/2) weights) and four bands ((dof/2 - 1) = 1/2) keep the lg2 form."""
    m, lab, x, xe, xm = _case(5003, 256)
    p, lm, le, po, lmo, leo, st = _run(fz, m, lab, x, xe, xm, dict(FS, dim_prior=False))
    assert st["sweep_kind"] == 2
    _check(p, lm, le, po, lmo, leo)
    m, lab, x, xe, xm = _case(5003, 256, nf=4)
    p, lm, le, po, lmo, leo, st = _run(fz, m, lab, x, xe, xm, FS)
    assert st["sweep_kind"] == 2
    _check(p, lm, le, po, lmo, leo)


def test_masked_nonfinite_and_bright_objects_leave_the_fp32_path(fz):
    m, lab, x, xe, xm = _case(5003, 512)
    xm[3, 1] = 0.0                      # masked band: (dof/2 - 1) = 1/2, not served by the linear-domain kernel
    x[7, 4] = np.nan                    # cleaned to a masked band (pdf.py:310-311)
    xe[9, 0] = -1.0
    x[11] = 3e4 * xe[11] * m[100] / m[100, 2]       # S/N ~ 1e5: beyond the tf32-split bound
    p, lm, le, po, lmo, leo, st = _run(fz, m, lab, x, xe, xm, FS)
    assert st["sweep_kind"] == 3 and st["objects_fp64"] >= 4
    _check(p, lm, le, po, lmo, leo)


def test_many_masked_objects_switch_to_the_log_domain_kernel(fz):
    m, lab, x, xe, xm = _case(4098, 256)
    xm[::3, 0] = 0.0
    p, lm, le, po, lmo, leo, st = _run(fz, m, lab, x, xe, xm, FS)
    assert st["sweep_kind"] == 2
    _check(p, lm, le, po, lmo, leo)


def test_batch_and_split_invariance(fz):
    """The arg-max (hence lmap) is decided on frame-independent log-domain values: the same object gives the same
    lmap bit for bit whatever the batch (which sets the model split count) it is processed in."""
    m, lab, x, xe, xm = _case(20011, 3000)
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(m), 0.05)
    bf = fz.BruteForce(m, np.zeros_like(m), np.ones_like(m))
    out = []
    for sel in (slice(0, 3000), slice(0, 100), slice(50, 1500)):
        p, (lm, le) = bf.fit_predict(x[sel].copy(), xe[sel].copy(), xm[sel].copy(), lab, labe, label_dict=rdict,
                                     return_gof=True, verbose=False, save_fits=False, lprob_kwargs=FS)
        out.append((sel, p, lm, le, bf.best_idx.copy()))
    _, p0, lm0, le0, b0 = out[0]
    for sel, p, lm, le, b in out[1:]:
        assert np.array_equal(b, b0[sel]) and np.array_equal(lm, lm0[sel])
        assert np.max(np.abs(le - le0[sel])) <= 2e-6 and np.max(np.sum(np.abs(p - p0[sel]), axis=1)) <= 2e-6


def test_fused_single_pass_matches_two_pass(fz, monkeypatch):
    """The single-pass variant (coarse pre-pass, histogram filled under the running cut, float64 re-decision of the
    recorded band) against the two-pass sweep on the same objects, with and without a per-model prior: same best model
    and lmap, evidence and PDFs within the fp32 noise of the two summation orders, and both within the bounds of the
    oracle.  The counters say that the fused pass really carried most objects."""
    m, lab, depth = bench_data.c3_models(float64_grid=True)      # the full 199,950-model grid: the coarse pre-pass needs a
    x, xe, xm, _, _ = bench_data.c3_objects(2048, m, depth, seed=77)   # dense model set to bound the maximum closely
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(m), 0.05)
    for lnprior in (None, np.random.RandomState(4).uniform(-3, 0, len(m))):
        kw = dict(FS) if lnprior is None else dict(FS, lnprior=lnprior)
        bf = fz.BruteForce(m, np.zeros_like(m), np.ones_like(m))
        res = {}
        for fused in (True, False):
            if not fused:
                monkeypatch.setenv("FZB_NO_FUSE", "1")
            p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), lab, labe, label_dict=rdict, return_gof=True,
                                         verbose=False, save_fits=False, lprob_kwargs=kw)
            monkeypatch.delenv("FZB_NO_FUSE", raising=False)
            res[fused] = (p, lm, le, bf.best_idx.copy(), bf._eng().stats())
        (p1, lm1, le1, b1, st1), (p2, lm2, le2, b2, st2) = res[True], res[False]
        assert st1["sweep_kind"] == 3 and st2["sweep_kind"] == 3
        assert st2["objects_fused"] == 0 and st1["objects_fused"] >= (0.4 if lnprior is None else 0.1) * len(x), (st1, st2)
        assert st1["pairs_pass2"] < st2["pairs_pass2"], (st1["pairs_pass2"], st2["pairs_pass2"])
        # the fused pass sweeps the faint objects without the float64 remainder of the model fluxes (FZB_TC_MLO_SNR): models
        # within the fp32 rounding of the maximum can swap places, lmap stays the exact value of the chosen one
        assert np.mean(b1 == b2) >= 0.9 and np.all(np.abs(lm1 - lm2) <= 1e-6 * np.maximum(1, np.abs(lm2)))
        assert np.max(np.abs(le1 - le2)) <= 3e-6
        assert np.max(np.sum(np.abs(p1 - p2), axis=1)) <= 2e-6
        sub = np.arange(0, len(x), 128)
        kd = fo.KernelDict(zgrid, sig)
        with np.errstate(all="ignore"):
            po, lmo, leo = fo.bruteforce_fit_predict(m, np.zeros_like(m), np.ones_like(m), x[sub].copy(), xe[sub].copy(),
                                                     xm[sub].copy(), lab, labe, label_dict=kd, lnprior=lnprior, **FS)
        _check(p1[sub], lm1[sub], le1[sub], po, lmo, leo)


@pytest.mark.parametrize("wt_thresh", [1e-2, 1e-5])
def test_fused_single_pass_other_weight_thresholds(fz, wt_thresh):
    """The running-cut logic of the single pass with cuts other than the default 1e-3 (kde_kwargs wt_thresh, pdf.py:589-591):
    the fused path against the float64 kernels on the full model grid."""
    m, lab, depth = bench_data.c3_models(float64_grid=True)
    x, xe, xm, _, _ = bench_data.c3_objects(768, m, depth, seed=91)
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(m), 0.05)
    bf = fz.BruteForce(m, np.zeros_like(m), np.ones_like(m))
    out = {}
    for prec in ("auto", "fp64"):
        p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), lab, labe, label_dict=rdict, return_gof=True, verbose=False,
                                     save_fits=False, lprob_kwargs=dict(FS, precision=prec), kde_kwargs=dict(wt_thresh=wt_thresh))
        out[prec] = (p, lm, le, bf._eng().stats())
    (p, lm, le, st), (p64, lm64, le64, _) = out["auto"], out["fp64"]
    assert st["sweep_kind"] == 3 and st["objects_fused"] > 0.3 * len(x), st
    assert np.max(np.sum(np.abs(p - p64), axis=1)) <= 1e-5
    assert np.all(np.abs(lm - lm64) <= 1e-5 * np.maximum(1, np.abs(lm64)))
    assert np.all(np.abs(le - le64) <= 1e-5 * np.maximum(1, np.abs(le64)))


@pytest.mark.parametrize("which", ["faint", "bright", "one"])
def test_fused_pass_with_one_sided_batches(fz, which):
    """Batches in which one of the two object lists of the sweep (faint: fused single pass; bright: seeded pass 1 + pass 2) is
    empty, and a batch of a single object: against the float64 kernels."""
    m, lab, depth = bench_data.c3_models(float64_grid=True)
    x, xe, xm, _, _ = bench_data.c3_objects(6000, m, depth, seed=123)
    snr = np.sqrt(np.sum((x / xe) ** 2, axis=1))
    order = np.argsort(snr)
    sel = {"faint": order[:700], "bright": order[-700:], "one": order[2500:2501]}[which]
    assert (which != "faint" or snr[sel].max() < 32) and (which != "bright" or snr[sel].min() > 32)
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(m), 0.05)
    bf = fz.BruteForce(m, np.zeros_like(m), np.ones_like(m))
    out = {}
    for prec in ("auto", "fp64"):
        p, (lm, le) = bf.fit_predict(x[sel].copy(), xe[sel].copy(), xm[sel].copy(), lab, labe, label_dict=rdict, return_gof=True,
                                     verbose=False, save_fits=False, lprob_kwargs=dict(FS, precision=prec))
        out[prec] = (p, lm, le, bf._eng().stats())
    (p, lm, le, st), (p64, lm64, le64, _) = out["auto"], out["fp64"]
    assert st["sweep_kind"] == 3
    assert (st["objects_fused"] > 0.5 * len(sel)) == (which != "bright"), st
    assert np.max(np.sum(np.abs(p - p64), axis=1)) <= 1e-5
    assert np.all(np.abs(lm - lm64) <= 1e-5 * np.maximum(1, np.abs(lm64)))
    assert np.all(np.abs(le - le64) <= 1e-5 * np.maximum(1, np.abs(le64)))


def test_object_conditioned_prior_table_on_the_fused_path(fz, monkeypatch):
    """SURVEY 8f rank 1 at scale: fit_predict(save_fits=False) with lnprior[i, j] = table[bin_i, j] runs the fused sweep
    once per table row; it must agree with the float64 kernel that reads the table per pair and with the oracle."""
    import frankenz_b200.bruteforce as bfm
    m, lab, x, xe, xm = _case(6001, 400)
    rs = np.random.RandomState(4)
    table = rs.uniform(-5, 0, size=(6, len(m)))
    bins = rs.randint(0, 6, size=len(x))
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(len(m), 0.05)
    kw = dict(FS, lnprior=table, lnprior_bin=bins)
    bf = fz.BruteForce(m, np.zeros_like(m), np.ones_like(m))
    res = {}
    for name, thresh in (("table_kernel", 1e30), ("by_row", 0.0)):
        monkeypatch.setattr(bfm, "TABLE_PRIOR_GROUP_MIN_PAIRS", thresh)
        p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), lab, labe, label_dict=rdict, return_gof=True,
                                     verbose=False, save_fits=False, lprob_kwargs=kw)
        res[name] = (p, lm, le, bf._eng().stats()["sweep_kind"])
    assert res["by_row"][3] == 3 and res["table_kernel"][3] == 0
    assert np.max(np.sum(np.abs(res["by_row"][0] - res["table_kernel"][0]), axis=1)) <= 1e-5
    assert np.allclose(res["by_row"][1], res["table_kernel"][1], rtol=0, atol=1e-5)
    assert np.allclose(res["by_row"][2], res["table_kernel"][2], rtol=0, atol=1e-5)
    kd = fo.KernelDict(zgrid, sig)
    for b in range(6):
        idx = np.nonzero(bins == b)[0][:12]
        with np.errstate(all="ignore"):
            po, lmo, leo = fo.bruteforce_fit_predict(m, np.zeros_like(m), np.ones_like(m), x[idx].copy(), xe[idx].copy(),
                                                     xm[idx].copy(), lab, labe, label_dict=kd, lnprior=table[b], **FS)
        _check(res["by_row"][0][idx], res["by_row"][1][idx], res["by_row"][2][idx], po, lmo, leo)


def test_six_bands(fz):
    """LSST-like six-band free-scale fit: three K = 8 instructions per product, 64-byte pair records; (dof/2 - 1) = 3/2
    keeps the lg2 form."""
    models, labels, depth = bench_data.c3_models()
    rs = np.random.RandomState(8)
    pick = np.sort(rs.choice(len(models), 5003, replace=False))
    m5 = models[pick]
    m = np.concatenate([m5, 0.8 * m5[:, 4:5] + 0.3 * m5[:, 3:4]], axis=1).astype(np.float32).astype(np.float64)
    d6 = np.append(depth, 0.2)
    n = 300
    j = rs.randint(0, len(m), size=n)
    x = m[j] * (10.0 ** rs.uniform(-1, 2, size=(n, 1))) / m[j, 2:3] + rs.normal(size=(n, 6)) * d6
    xe = np.broadcast_to(d6, x.shape).copy()
    xm = np.ones_like(x)
    xm[5, 0] = 0.0
    for lprob, lnprior in ((FS, None), (dict(FS, dim_prior=False), rs.uniform(-3, 0, len(m)))):
        p, lm, le, po, lmo, leo, st = _run(fz, m, labels[pick], x, xe, xm, lprob, lnprior=lnprior)
        assert st["sweep_kind"] == 2 and st["pairs_fp32"] > 0
        _check(p, lm, le, po, lmo, leo)
