"""Throughput of the modes that do NOT take the sweep kernels (VERDICT r1 missing #7): iterated free scale with model
errors (pdf.py:197-223), dim_prior=False with model errors (pdf.py:96-98), the CDF threshold rule (pdf.py:592-597), the
exact-Gaussian grid KDE (pdf.py:444-526), three filters (outside 4-6) and non-binary model masks.  All run on the
one-CTA-per-object float64 kernel `k_generic` (reference operation order); the oracle port gives the CPU figure beside it
on a sample.  SDSS ugriz mock of C1: 2,000 objects x 20,000 training models."""
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
kd = fo.KernelDict(zgrid, sig)
labe = np.full(len(m), 0.05)
rs = np.random.RandomState(1)
mm_frac = np.where(rs.uniform(size=mm.shape) < 0.1, 0.5, 1.0)
cases = [
    ("iterated free scale (free_scale=True, model errors)", dict(lprob_kwargs=dict(free_scale=True)), {}, (m, me, mm), 5),
    ("dim_prior=False with model errors", dict(lprob_kwargs=dict(dim_prior=False)), {}, (m, me, mm), 5),
    ("default likelihood, cdf_thresh=2e-3 instead of wt_thresh", dict(kde_kwargs=dict(wt_thresh=None, cdf_thresh=2e-3)), {}, (m, me, mm), 5),
    ("default likelihood, exact-Gaussian grid KDE (label_grid)", dict(), dict(grid=True), (m, me, mm), 5),
    ("default likelihood, three filters", dict(), {}, (m[:, :3].copy(), me[:, :3].copy(), mm[:, :3].copy()), 3),
    ("default likelihood, fractional model masks", dict(), {}, (m, me, mm_frac), 5),
]
for name, kw, opt, (mA, meA, mmA), nf in cases:
    bf = fz.BruteForce(mA, meA, mmA)
    xs, xes, xms = x[:, :nf].copy(), xe[:, :nf].copy(), xm[:, :nf].copy()
    kde = dict(label_grid=zgrid) if opt.get("grid") else dict(label_dict=rdict)
    ts = []
    for rep in range(3):
        t = time.time()
        p = bf.fit_predict(xs.copy(), xes.copy(), xms.copy(), z, labe, verbose=False, save_fits=False, **kde, **kw)
        ts.append(time.time() - t)
    st = bf._eng().stats()
    assert st["sweep_kind"] == 0, "expected the generic float64 kernel"
    dt = min(ts[1:])
    n = 8
    t = time.time()
    with np.errstate(all="ignore"):
        okw = dict(kw.get("lprob_kwargs", {}))
        kk = kw.get("kde_kwargs", {})
        po, _, _ = fo.bruteforce_fit_predict(mA, meA, mmA, xs[:n].copy(), xes[:n].copy(), xms[:n].copy(), z, labe,
                                             label_dict=None if opt.get("grid") else kd,
                                             label_grid=zgrid if opt.get("grid") else None, **okw, **({"kde_kwargs": kk} if kk else {}))
    tcpu = time.time() - t
    l1 = np.nanmax(np.sum(np.abs(p[:n] - po), axis=1))
    print("%-62s GPU %.3f s = %.3e pairs/s (%.0f objects/s); oracle on 1 core %.3e pairs/s; PDF L1 vs oracle %.1e"
          % (name, dt, len(xs) * len(mA) / dt, len(xs) / dt, n * len(mA) / tcpu, l1))
