"""The reference's default likelihood (fixed scale, model errors, dim_prior) at the shape of bench.py's `default_likelihood` leg:
262,144 objects x 262,144 training rows, 6 bands, device-resident inputs; prints the phases of the step.
Usage: python tools/bench_fx1.py [Ntrain] [Nobj]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data  # noqa: E402
import frankenz_b200 as fz  # noqa: E402
from frankenz_b200._engine import make_config  # noqa: E402

n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n_obj = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
tr, tre, trm, ztr, x, xe, xm = bench_data.c5_dataset(n_train, n_obj, seed=20260107)
zgrid, sig = bench_data.c3_kde()
bf = fz.BruteForce(tr, tre, trm)
eng = bf._eng()
eng.set_kde(ztr, np.full(n_train, 0.05), label_dict=fz.pdf.PDFDict(zgrid, sig))
cfg = make_config(dict(), None)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (x, xe, xm)]
out = [torch.empty((n_obj, eng.Ng), dtype=torch.float64, device=dev)] + \
      [torch.empty(n_obj, dtype=torch.float64, device=dev) for _ in range(2)] + \
      [torch.empty(n_obj, dtype=torch.int64, device=dev)] + \
      [torch.empty(n_obj, dtype=torch.float64, device=dev) for _ in range(2)]
for rep in range(3):
    torch.cuda.synchronize()
    eng.fit_predict_dev(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n_obj, cfg, *[t.data_ptr() for t in out])
    st = eng.stats()
    print("rep %d: total %.1f ms (first sweep %.1f, pass 2 + fixes %.1f, finish %.1f), %.3e pairs/s; objects to float64 %d, "
          "pass-2 pairs evaluated %.3f of all, launches %d" % (rep, st["ms_total"], st["ms_scan"], st["ms_accum"], st["ms_finish"],
          float(n_obj) * n_train / (st["ms_total"] * 1e-3), st["objects_fp64"], st["pairs_pass2"] / (float(n_obj) * n_train),
          st["kernel_launches"]))
