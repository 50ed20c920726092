"""Synthetic MockSurvey-style workloads of SURVEY.md section 8d (no reference import: the inputs
that needed the reference were generated once by tests/golden/make_mock_inputs.py).

C3  template fitting: HSC grizy `brown` template x redshift grid (1550 x 129 = 199,950 models,
    model errors 0, masks 1); objects are grid models scaled to a reference-band (i) magnitude
    drawn from P(m) ~ m^15 exp(-(m/(maglim-1))^2) on [18, 26.4] (frankenz/priors.py:27-75 with
    maglim = 25.9) plus Gaussian noise of the survey depth (frankenz/simulate.py:481).
C1  SDSS ugriz cww+/BPZ mock: training rows are models with their own errors.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ZP = 23.9   # AB zero-point of the uJy-like flux units (simulate.py:481)


def c3_models(float64_grid=False):
    """HSC grizy brown template x redshift grid (199,950 x 5).  float64_grid=False: the fluxes rounded to float32
    (every value exactly fp32-representable; the input class of round 1).  float64_grid=True: the float64 output of the
    reference's `make_model_grid` (simulate.py:954-1021), bit for bit: float32 part + the committed remainder."""
    d = np.load(os.path.join(GOLDEN, "hsc_brown_grid.npz"))
    models = np.ascontiguousarray(d["models"], dtype=np.float64)
    if float64_grid:
        models = models + np.load(os.path.join(GOLDEN, "hsc_brown_grid_lo.npz"))["models_lo"]
    zgrid, nt = d["zgrid"], int(d["ntemplate"])
    labels = np.repeat(zgrid, nt)
    return models, labels, d["depth_flux1sig"].astype(np.float64)


def lsst_models():
    """LSST ugrizY brown template x redshift grid (199,950 x 6), SURVEY.md section 8d, C5."""
    d = np.load(os.path.join(GOLDEN, "lsst_brown_grid.npz"))
    models = np.ascontiguousarray(d["models"], dtype=np.float64)
    labels = np.repeat(d["zgrid"], int(d["ntemplate"]))
    return models, labels, d["depth_flux1sig"].astype(np.float64)


def c5_dataset(n_train, n_obj, seed=20260105, ref=3):
    """C5-shaped problem: `n_train` training rows and `n_obj` objects drawn like the C3 objects
    from the 6-band LSST grid (float64 noisy fluxes, errors = survey depth), labels = redshift of the parent model.
    Every rank of a model-sharded run calls this with the same seed and keeps its slice."""
    grid, zlab, depth = lsst_models()
    tr, tre, trm, j, _ = c3_objects(n_train, grid, depth, seed=seed, ref=ref)
    x, xe, xm, _, _ = c3_objects(n_obj, grid, depth, seed=seed + 1, ref=ref)
    return tr, tre, trm, zlab[j], x, xe, xm


def c4_dataset(n_train, n_query, seed=5):
    """C4-shaped kNN problem: training rows and queries drawn like the C3 objects from the HSC
    grid, with the luptitude feature-map arguments of demo 2 cell 56."""
    models, labels, depth = c3_models()
    x, xe, xm, j, _ = c3_objects(n_train + n_query, models, depth, seed=seed)
    fmap = dict(skynoise=depth, zeropoints=10 ** (-0.4 * -23.9))
    tr = (x[:n_train], xe[:n_train], xm[:n_train], labels[j[:n_train]])
    q = (x[n_train:], xe[n_train:], xm[n_train:])
    return tr, q, fmap


def draw_mags(n, rs, lo=18.0, hi=26.4, maglim=25.9):
    grid = np.linspace(lo, hi, 4000)
    pm = grid ** 15.0 * np.exp(-(grid / (maglim - 1.0)) ** 2.0)
    cdf = np.cumsum(pm)
    cdf /= cdf[-1]
    return np.interp(rs.uniform(size=n), cdf, grid)


def c3_objects(n, models, depth, seed=20260103, ref=2):
    rs = np.random.RandomState(seed)
    j = rs.randint(0, len(models), size=n)
    mag = draw_mags(n, rs)
    fref = 10.0 ** (-0.4 * (mag - ZP))
    flux = models[j] * (fref / models[j, ref])[:, None]
    data = flux + rs.normal(size=flux.shape) * depth[None, :]
    err = np.broadcast_to(depth, data.shape).copy()
    mask = np.ones_like(data)
    return np.ascontiguousarray(data), err, mask, j, mag


def c3_kde():
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    sig = np.linspace(0.005, 2, 500)
    return zgrid, sig


def c1_dataset(ntrain=20000, ntest=2000):
    d = np.load(os.path.join(GOLDEN, "sdss_cww_mock.npz"))
    phot, err, z = d["phot_obs"], d["phot_err"], d["redshifts"]
    n = len(phot)
    # the committed mock holds ~4.9k selected objects; tile it with fresh noise to reach the C1 sizes
    rs = np.random.RandomState(7)
    need = ntrain + ntest
    rep = (need + n - 1) // n
    true = np.tile(d["phot_true"], (rep, 1))[:need]
    e = np.tile(err, (rep, 1))[:need]
    zz = np.tile(z, rep)[:need]
    obs = true + rs.normal(size=true.shape) * e
    return (obs[:ntrain], e[:ntrain], np.ones((ntrain, 5)), zz[:ntrain],
            obs[ntrain:], e[ntrain:], np.ones((ntest, 5)), zz[ntrain:])
