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
    assert c3["stats"]["pairs_fp32"] > 0.9 * 2 * len(c3["x"]) * len(c3["models"])     # the fp32 kernels did the work
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
