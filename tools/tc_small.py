"""Small tensor-core-sweep run (for compute-sanitizer): 96 objects x 4081 models of the float64 grid; both passes with
live bits / pruning, the cut records and their float64 fix-up, merge, finish."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data, frankenz_b200 as fz
models, labels, depth = bench_data.c3_models(float64_grid=True)     # the float64 grid: MLO tiles
m, lab = models[::49].copy(), labels[::49].copy()
x, xe, xm, _, _ = bench_data.c3_objects(96, m, depth, seed=3)
zgrid, sig = bench_data.c3_kde()
bf = fz.BruteForce(m, np.zeros_like(m), np.ones_like(m))
p, (lm, le) = bf.fit_predict(x, xe, xm, lab, np.full(len(m), 0.05), label_dict=fz.pdf.PDFDict(zgrid, sig), return_gof=True,
                             verbose=False, save_fits=False, lprob_kwargs=dict(free_scale=True, ignore_model_err=True))
print("sweep_kind", bf._eng().stats()["sweep_kind"], "pdf sums", p.sum(axis=1)[:3], "lmap", lm[:3])
