"""GPU-side calibration of the fp32 path's routing thresholds: runs the C3 workload on a sample of
objects through the fp32 path (forced) and the float64 path and reports the PDF L1 error against
best-fit chi2 / S/N, plus what the default routing does.  Usage: python tools/calibrate_fast.py [N]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data  # noqa: E402
import frankenz_b200 as fz  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
models, labels, depth = bench_data.c3_models()
nm_use = int(os.environ.get("NM", len(models)))
models, labels = models[:nm_use], labels[:nm_use]
x, xe, xm, jtrue, mag = bench_data.c3_objects(n, models, depth)
zgrid, sig = bench_data.c3_kde()
rdict = fz.pdf.PDFDict(zgrid, sig)
bf = fz.BruteForce(models, np.zeros_like(models), np.ones_like(models))
labe = np.full(len(models), 0.05)
res = {}
for mode in ("fp64", "fp32", "auto"):
    kw = dict(free_scale=True, ignore_model_err=True, dim_prior=True, precision=mode)
    t = time.time()
    p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), labels, labe, label_dict=rdict, return_gof=True,
                                 verbose=False, save_fits=False, lprob_kwargs=kw)
    dt = time.time() - t
    res[mode] = (p, lm, le, bf.best_idx.copy(), bf.best_chi2.copy())
    print(mode, "wall %.3fs" % dt, bf._eng().stats())
p64, lm64, le64, bi64, bc64 = res["fp64"]
snr = np.sqrt(np.sum((x / xe) ** 2, axis=1))
for mode in ("fp32", "auto"):
    p, lm, le, bi, bc = res[mode]
    l1 = np.sum(np.abs(p - p64), axis=1)
    dl = np.abs(lm - lm64) / np.maximum(1, np.abs(lm64))
    de = np.abs(le - le64) / np.maximum(1, np.abs(le64))
    print("== %s: max L1 %.3g  (99.9%% %.3g, median %.3g)  max dlmap %.3g  max dlevid %.3g  argmax mismatch %d"
          % (mode, np.nanmax(l1), np.nanpercentile(l1, 99.9), np.nanmedian(l1), np.nanmax(dl), np.nanmax(de),
             int(np.sum(bi != bi64))))
    if mode == "fp32":
        for lo, hi in ((0, 4), (4, 8), (8, 16), (16, 24), (24, 48), (48, 1e9)):
            s = (bc64 >= lo) & (bc64 < hi)
            if s.any():
                print("   chi2_best in [%g,%g): n=%d  max L1 %.3g  median %.3g" % (lo, hi, s.sum(), l1[s].max(),
                                                                                   np.median(l1[s])))
        for lo, hi in ((0, 10), (10, 30), (30, 100), (100, 1000), (1000, 5000), (5000, 1e9)):
            s = (snr >= lo) & (snr < hi)
            if s.any():
                print("   snr_tot in [%g,%g): n=%d  max L1 %.3g  median %.3g" % (lo, hi, s.sum(), l1[s].max(),
                                                                                 np.median(l1[s])))

# diagnose arg-max mismatches: exact lnl of both candidates
p, lm, le, bi, bc = res["auto"]
mis = np.where(bi != bi64)[0]
print("mismatches:", len(mis))
for o in mis[:12]:
    sub = np.array([bi[o], bi64[o]])
    b2 = fz.BruteForce(models[sub], np.zeros((2, 5)), np.ones((2, 5)))
    b2.fit(x[o:o + 1].copy(), xe[o:o + 1].copy(), xm[o:o + 1].copy(), verbose=False,
           lprob_kwargs=dict(free_scale=True, ignore_model_err=True, dim_prior=True))
    print("  obj %d snr %.1f  fast j=%d  fp64 j=%d  lnl(fast j)=%.12g lnl(fp64 j)=%.12g  lmap64=%.12g z=%.4f/%.4f"
          % (o, snr[o], bi[o], bi64[o], b2.fit_lnlike[0, 0], b2.fit_lnlike[0, 1], lm64[o], labels[bi[o]],
             labels[bi64[o]]))
