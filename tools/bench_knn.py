"""kNN timing at config-C4-like sizes: 1M-row training set, K Monte-Carlo realisations,
k neighbours, luptitude features.  Usage: python tools/bench_knn.py [Ntrain] [Nquery] [K] [k]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data  # noqa: E402
import frankenz_b200 as fz  # noqa: E402

ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
k = int(sys.argv[4]) if len(sys.argv) > 4 else 25
models, labels, depth = bench_data.c3_models()
x, xe, xm, j, mag = bench_data.c3_objects(ntr + nq, models, depth, seed=5)
keep = (x[:, 2] / xe[:, 2]) > 5
x, xe, xm, j = x[keep], xe[keep], xm[keep], j[keep]
ntr = min(ntr, len(x) - nq)
tr, tre, trm, ztr = x[:ntr], xe[:ntr], xm[:ntr], labels[j[:ntr]]
qx, qe, qm = x[ntr:ntr + nq], xe[ntr:ntr + nq], xm[ntr:ntr + nq]
kw = dict(skynoise=depth, zeropoints=10 ** (-0.4 * -23.9))
t = time.time()
nn = fz.NearestNeighbors(tr, tre, trm, K=K, fmap_kwargs=kw, rstate=np.random.RandomState(1), verbose=False)
t_build = time.time() - t
print("build: %d rows x %d trees in %.2f s (host MC draws + feature map + H2D)" % (ntr, K, t_build))
zgrid, sig = bench_data.c3_kde()
rdict = fz.pdf.PDFDict(zgrid, sig)
for rep in range(2):
    t = time.time()
    p = nn.fit_predict(qx.copy(), qe.copy(), qm.copy(), ztr, np.full(ntr, 0.05), label_dict=rdict, k=k, eps=0,
                       rstate=np.random.RandomState(2), verbose=False)
    dt = time.time() - t
    st = nn._engine.stats()
    print("fit_predict: %d queries in %.3f s -> %.3e queries/s, %.3e distance evaluations/s (device ms of last call %.1f)"
          % (len(qx), dt, len(qx) / dt, len(qx) * K * ntr / dt, st["ms_total"]))
print("Nneighbors min/median/max:", nn.Nneighbors.min(), np.median(nn.Nneighbors), nn.Nneighbors.max())
# timing of the search alone
q = nn._query_features(qx, qe, np.random.RandomState(2))
for rep in range(2):
    t = time.time()
    idx, dist = nn._engine.knn_query(q, k, p=2)
    dt = time.time() - t
    st = nn._engine.stats()
    print("knn_query alone: %.3f s wall, call %.1f ms, search %.1f ms -> %.3e distance evaluations/s; tensor-core scan %d, "
          "redo %d (overflow %d) of %d, largest candidate error %.3g" %
          (dt, st["ms_total"], st["ms_scan"], len(q) * K * ntr / (st["ms_scan"] * 1e-3), st["knn_tc"], st["knn_redo"],
           st["knn_overflow"], len(q) * K, st["knn_tc_err"]))
