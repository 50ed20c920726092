"""Small runs for compute-sanitizer.  Tensor-core sweep: 600 objects x 4081 models of the float64 grid - coarse pre-pass, fused
single pass (faint objects, fp32-rounded tiles, records, compact list, float64 re-decisions), seeded pass 1 + pruned pass 2 with
tile masks (bright objects), merge, finish.  kNN: 6,000 rows x 3 trees, 700 queries - staged filter / select for tree 0, shared
threshold for the others, re-rank, float64 re-dos."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data, frankenz_b200 as fz
models, labels, depth = bench_data.c3_models(float64_grid=True)     # the float64 grid: MLO tiles
m, lab = models[::49].copy(), labels[::49].copy()
x, xe, xm, _, _ = bench_data.c3_objects(600, m, depth, seed=3)
zgrid, sig = bench_data.c3_kde()
bf = fz.BruteForce(m, np.zeros_like(m), np.ones_like(m))
p, (lm, le) = bf.fit_predict(x, xe, xm, lab, np.full(len(m), 0.05), label_dict=fz.pdf.PDFDict(zgrid, sig), return_gof=True,
                             verbose=False, save_fits=False, lprob_kwargs=dict(free_scale=True, ignore_model_err=True))
print("sweep_kind", bf._eng().stats()["sweep_kind"], "pdf sums", p.sum(axis=1)[:3], "lmap", lm[:3])
st = bf._eng().stats()
print("objects fused", st["objects_fused"], "pass-2 pairs", st["pairs_pass2"], "weights re-decided", st["cut_recorded"])
from frankenz_b200._engine import Engine
rs = np.random.RandomState(2)
base = rs.normal(size=(6000, 5)) + 20.0
feats = np.stack([base + rs.normal(size=base.shape) * 0.02 for _ in range(3)]).astype(np.float32)
q = base[rs.choice(len(base), 700)] + rs.normal(size=(700, 5)) * 0.05
q[3] = np.nan
ones = np.ones((len(base), 5))
eng = Engine(ones, ones, ones)
eng.knn_build(feats)
idx, dist = eng.knn_query(q, 25, p=2)
st = eng.stats()
print("knn: redo", st["knn_redo"], "overflow", st["knn_overflow"], "first neighbours", idx[0, 0, :3], dist[0, 0, :3])
