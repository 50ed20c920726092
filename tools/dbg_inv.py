import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data, frankenz_b200 as fz
models, labels, depth = bench_data.c3_models()
x, xe, xm, jtrue, mag = bench_data.c3_objects(16384, models, depth)
zgrid, sig = bench_data.c3_kde(); rdict = fz.pdf.PDFDict(zgrid, sig)
labe = np.full(len(models), 0.05)
kw = dict(free_scale=True, ignore_model_err=True, dim_prior=True)
bf = fz.BruteForce(models, np.zeros_like(models), np.ones_like(models))
p, (lm, le) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), labels, labe, label_dict=rdict, return_gof=True, verbose=False, save_fits=False, lprob_kwargs=kw)
bi = bf.best_idx.copy()
sel = np.random.RandomState(3).choice(len(x), 3000, replace=False)
p2, (lm2, le2) = bf.fit_predict(x[sel].copy(), xe[sel].copy(), xm[sel].copy(), labels, labe, label_dict=rdict, return_gof=True, verbose=False, save_fits=False, lprob_kwargs=kw)
bi2 = bf.best_idx.copy()
print("L1", np.max(np.sum(np.abs(p2 - p[sel]), axis=1)), "dlm", np.max(np.abs(lm2 - lm[sel])), "dle", np.max(np.abs(le2 - le[sel])), "best differs", np.sum(bi2 != bi[sel]))
w = np.argsort(-np.abs(lm2 - lm[sel]))[:5]
for k in w: print(k, lm2[k], lm[sel][k], bi2[k], bi[sel][k], le2[k], le[sel][k])
