"""Host-side breakdown of NearestNeighbors.fit_predict(save_fits=False) at C4-like sizes: which part of the wall time
is the device search and which is host work.  Usage: python tools/knn_e2e_breakdown.py [Ntrain] [Nquery] [K] [k]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data  # noqa: E402
import frankenz_b200 as fz  # noqa: E402
from frankenz_b200._engine import clean_inplace, make_config  # noqa: E402

ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
k = int(sys.argv[4]) if len(sys.argv) > 4 else 25
(tr, tre, trm, ztr), (qx, qe, qm), fmap = bench_data.c4_dataset(ntr, nq)
nn = fz.NearestNeighbors(tr, tre, trm, K=K, fmap_kwargs=fmap, rstate=np.random.RandomState(1), verbose=False)
zgrid, sig = bench_data.c3_kde()
rdict = fz.pdf.PDFDict(zgrid, sig)
labe = np.full(ntr, 0.05)
eng = nn._engine
for rep in range(3):
    T = [time.perf_counter()]
    x, xe, xm = qx.copy(), qe.copy(), qm.copy()
    T.append(time.perf_counter())
    eng.set_lnprior(None, None)
    eng.set_kde(ztr, labe, label_dict=rdict)
    T.append(time.perf_counter())
    cfg = make_config({}, None, track_scale=False)
    q = nn._query_features(x, xe, np.random.RandomState(2))
    T.append(time.perf_counter())
    clean_inplace(x, xe, xm)
    T.append(time.perf_counter())
    pdfs, lmap, levid, nnb = eng.knn_fit_predict(q, x, xe, xm, k, 2, cfg)
    T.append(time.perf_counter())
    st = eng.stats()
    names = ["copies", "set_kde", "query features (MC draw + feature map)", "clean", "knn_fit_predict call"]
    print("rep %d: total %.3f s; " % (rep, T[-1] - T[0]) + "; ".join("%s %.3f" % (n, T[i + 1] - T[i]) for i, n in enumerate(names)) +
          "; inside the call: search %.1f ms, device total %.1f ms" % (st["ms_scan"], st["ms_total"]))
