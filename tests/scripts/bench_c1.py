"""Timing of configs C1 / C2 through the public API, with the oracle port beside it.
C1: SDSS ugriz mock, 2k objects x 20k training models, default likelihood: fit + predict, and fit_predict.
C2: self-fit N x N with 5% band dropouts and 0.5% NaN fluxes (model masks -> float64 generic kernel)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_data  # noqa: E402
import frankenz_b200 as fz  # noqa: E402
from oracle import fz_oracle as fo  # noqa: E402

m, me, mm, z, x, xe, xm, _ = bench_data.c1_dataset(20000, 2000)
zgrid, sig = bench_data.c3_kde()
rdict = fz.pdf.PDFDict(zgrid, sig)
labe = np.full(len(m), 0.05)
bf = fz.BruteForce(m, me, mm)
for rep in range(2):
    t = time.time()
    bf.fit(x.copy(), xe.copy(), xm.copy(), verbose=False)
    t1 = time.time()
    p = bf.predict(z, labe, label_dict=rdict, verbose=False)
    t2 = time.time()
    p2 = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict, verbose=False, save_fits=False)
    t3 = time.time()
    print("C1 rep %d: fit %.3f s (%.3e pairs/s, 7 x (2000 x 20000) arrays = 2.2 GB to host), predict %.3f s (%.0f obj/s), "
          "fit_predict(save_fits=False) %.4f s (%.3e pairs/s)" % (rep, t1 - t, 4e7 / (t1 - t), t2 - t1, 2000 / (t2 - t1),
                                                                  t3 - t2, 4e7 / (t3 - t2)))
kd = fo.KernelDict(zgrid, sig)
n = 48
t = time.time()
fit = fo.bruteforce_fit(m, me, mm, x[:n].copy(), xe[:n].copy(), xm[:n].copy())
t1 = time.time()
po, _, _ = fo.bruteforce_predict(fit["lnprob"], z, labe, label_dict=kd)
t2 = time.time()
print("C1 oracle (1 core, %d objects): fit %.3e pairs/s, predict %.1f obj/s" % (n, n * 2e4 / (t1 - t), n / (t2 - t1)))
print("   parity: PDF L1 max %.2e" % np.max(np.sum(np.abs(p[:n] - po), axis=1)))

# C2
N = int(os.environ.get("C2_N", 50000))
rs = np.random.RandomState(3)
mC, meC, _, zC, _, _, _, _ = bench_data.c1_dataset(N, 1)
mask = (rs.uniform(size=mC.shape) > 0.05).astype(float)
phot = mC.copy()
phot[rs.uniform(size=mC.shape) < 0.005] = np.nan
bf2 = fz.BruteForce(np.where(np.isfinite(phot), phot, 0.0), meC, mask)
for rep in range(2):
    t = time.time()
    p, (lm, le) = bf2.fit_predict(phot.copy(), meC.copy(), mask.copy(), zC, 0.01 * (1 + zC), label_dict=rdict,
                                  return_gof=True, verbose=False, save_fits=False)
    dt = time.time() - t
    print("C2 rep %d: %d x %d self-fit with masks/NaNs: %.3f s -> %.3e pairs/s  %s" % (rep, N, N, dt, N * N / dt,
                                                                                        bf2._eng().stats()))
print("   PDFs finite: %d of %d (self-match has chi2 = 0 -> lnL = -inf under dim_prior, SURVEY 8d C2)" %
      (int(np.isfinite(p).all(axis=1).sum()), N))
