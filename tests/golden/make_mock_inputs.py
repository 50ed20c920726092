#!/usr/bin/env python
"""Generate MockSurvey-style photometry with the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, read-only). The
outputs are committed under tests/golden/ so that nothing on the GPU box has to
import the reference:

  sdss_cww_mock.npz   SDSS ugriz / cww+ / BPZ-prior mock:
                      phot_obs, phot_err, phot_true, redshifts, mags  (S/N_r > 5 cut)
  hsc_brown_grid.npz  HSC grizy / brown template x redshift model grid
                      (SURVEY.md section 8d, C3): models[Nz*Nt, 5] float32, zgrid,
                      depth_flux1sig.
  hsc_brown_grid_lo.npz  the part of make_model_grid's float64 output that float32 drops:
                      models_lo = models64 - float32(models64) (exact in float64), so that
                      float64(models) + models_lo IS the reference's float64 grid, bit for bit.
  lsst_brown_grid.npz LSST ugrizY / brown grid, 6 bands: models float32.

Reference calls exercised: frankenz/simulate.py:398 (MockSurvey), :444
(load_survey; Npoints passed as int because the 5e4 default crashes np.linspace),
:600 (set_refmag), :880 (make_mock), :954 (make_model_grid).
"""
import os
import sys
import warnings

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def sdss_mock(ndraw=14000, seed=7):
    from frankenz import simulate
    np.random.seed(seed)
    s = simulate.MockSurvey(templates="cww+", prior="bpz")
    s.load_survey("sdss", Npoints=50000)
    s.set_refmag("r")
    s.make_mock(ndraw, mbounds=[14, 25], zbounds=[0, 6], verbose=False)
    d = s.data
    phot_obs, phot_err = d["phot_obs"], d["phot_err"]
    ref = s.ref_filter if hasattr(s, "ref_filter") else 2
    sel = (phot_obs[:, 2] / phot_err[:, 2]) > 5.0
    out = dict(phot_obs=phot_obs[sel], phot_err=phot_err[sel],
               phot_true=d["phot_true"][sel], redshifts=d["redshifts"][sel],
               mags=d["refmags"][sel] if "refmags" in d else np.zeros(sel.sum()),
               depth_flux1sig=np.array([f["depth_flux1sig"] for f in s.filters]))
    np.savez_compressed(os.path.join(HERE, "sdss_cww_mock.npz"), **out)
    print("sdss mock:", phot_obs.shape, "->", int(sel.sum()), "selected")


def hsc_grid(nz=1550):
    from frankenz import simulate
    s = simulate.MockSurvey(templates="brown")
    s.load_survey("hsc", Npoints=50000)
    zgrid = np.linspace(0, 6, nz)
    s.make_model_grid(zgrid, verbose=False)
    m = s.models["data"]  # (Nz, Nt, Nf)
    depth = np.array([f["depth_flux1sig"] for f in s.filters])
    m64 = np.ascontiguousarray(m.reshape(-1, m.shape[-1]), dtype=np.float64)
    m32 = m64.astype(np.float32)
    old = os.path.join(HERE, "hsc_brown_grid.npz")
    if os.path.exists(old):      # the committed float32 grid stays as it is; only check that it is reproduced
        assert np.array_equal(np.load(old)["models"], m32), "float32 grid changed"
    else:
        np.savez_compressed(old, models=m32, zgrid=zgrid, ntemplate=np.int64(m.shape[1]), depth_flux1sig=depth)
    lo = m64 - m32.astype(np.float64)
    assert np.array_equal(m32.astype(np.float64) + lo, m64)
    np.savez_compressed(os.path.join(HERE, "hsc_brown_grid_lo.npz"), models_lo=lo)
    print("hsc grid:", m.shape, "finite:", bool(np.isfinite(m).all()),
          "min:", float(m.min()), "fp32-exact fraction:", float(np.mean(lo == 0)))


def lsst_grid(nz=1550):
    from frankenz import simulate
    s = simulate.MockSurvey(templates="brown")
    s.load_survey("lsst", Npoints=50000)
    zgrid = np.linspace(0, 6, nz)
    s.make_model_grid(zgrid, verbose=False)
    m = s.models["data"]
    depth = np.array([f["depth_flux1sig"] for f in s.filters])
    np.savez_compressed(os.path.join(HERE, "lsst_brown_grid.npz"),
                        models=m.reshape(-1, m.shape[-1]).astype(np.float32), zgrid=zgrid,
                        ntemplate=np.int64(m.shape[1]), depth_flux1sig=depth,
                        filters=np.array([f["name"] for f in s.filters]))
    print("lsst grid:", m.shape, "finite:", bool(np.isfinite(m).all()), "min:", float(m.min()), "depth:", depth)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "sdss"):
        sdss_mock()
    if which in ("all", "hsc"):
        hsc_grid()
    if which in ("all", "lsst"):
        lsst_grid()
