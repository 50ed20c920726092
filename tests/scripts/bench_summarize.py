"""Timing of pdf.pdfs_summarize on PDFs shaped like the C3 output (Nobj x 701, float64).
Usage: python tests/scripts/bench_summarize.py [Nobj] [cpu_sample]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import frankenz_b200 as fz
from frankenz_b200._engine import SummaryEngine
from oracle import fz_oracle as fo

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
ncpu = int(sys.argv[2]) if len(sys.argv) > 2 else 512
zgrid = np.arange(0, 7 + 1e-5, 0.01)
rs = np.random.RandomState(3)
base = np.zeros((4096, len(zgrid)))
for i in range(len(base)):
    for _ in range(rs.randint(1, 4)):
        mu, sg = rs.uniform(0.05, 6.5), 10 ** rs.uniform(-2, -0.3)
        base[i] += rs.uniform(0.2, 1) * np.exp(-0.5 * ((zgrid - mu) / sg) ** 2)
pdfs = np.tile(base, ((n + len(base) - 1) // len(base), 1))[:n].copy()
for rep in range(3):
    p = pdfs.copy()
    t = time.perf_counter()
    res = fz.pdf.pdfs_summarize(p, zgrid, rstate=np.random.RandomState(1))
    dt = time.perf_counter() - t
    st = SummaryEngine.get().stats()
    flops = 2.0 * len(zgrid) ** 2 * n
    print("rep %d: %d PDFs, end to end %.3f s (%.3g PDFs/s); device loop incl. H2D of the PDFs %.1f ms; risk product alone "
          "would be %.2f TFLOP/s float64 over that time" % (rep, n, dt, n / dt, st["ms_total"], flops / (st["ms_total"] * 1e-3) / 1e12))
q = pdfs[:ncpu].copy()
t = time.perf_counter()
ref = fo.pdfs_summarize(q, zgrid, rstate=np.random.RandomState(1))
dc = time.perf_counter() - t
print("CPU oracle (1 process): %d PDFs in %.2f s = %.3g PDFs/s" % (ncpu, dc, ncpu / dc))
print("max |mean - oracle| %.2e, median identical: %s" % (np.max(np.abs(res[0][0][:ncpu] - ref[0][0])), np.array_equal(res[1][0][:ncpu], ref[1][0])))
