"""Where the end-to-end (numpy in / numpy out) time of BruteForce.fit_predict goes on the C3 workload."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data, frankenz_b200 as fz
from frankenz_b200._engine import make_config
from frankenz_b200.bruteforce import clean_inplace
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
models, labels, depth = bench_data.c3_models()
x, xe, xm, _, _ = bench_data.c3_objects(n, models, depth)
zgrid, sig = bench_data.c3_kde(); rdict = fz.pdf.PDFDict(zgrid, sig)
labe = np.full(len(models), 0.05)
kw = dict(free_scale=True, ignore_model_err=True, dim_prior=True)
bf = fz.BruteForce(models, np.zeros_like(models), np.ones_like(models))
for rep in range(3):
    t0 = time.perf_counter()
    eng, cfg = bf._setup(None, None, kw, False, None)
    t1 = time.perf_counter()
    eng.set_kde(labels, labe, label_dict=rdict)
    t2 = time.perf_counter()
    clean_inplace(x, xe, xm)
    t3 = time.perf_counter()
    out = eng.fit_predict(x, xe, xm, cfg)
    t4 = time.perf_counter()
    st = eng.stats()
    print("rep %d: setup %.1f ms, set_kde %.1f ms, clean_inplace %.1f ms, fzb_fit_predict wall %.1f ms (device loop %.1f ms: "
          "scan %.1f accum %.1f finish %.1f), total %.1f ms" % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2),
          1e3 * (t4 - t3), st["ms_total"], st["ms_scan"], st["ms_accum"], st["ms_finish"], 1e3 * (t4 - t0)))
    del out
