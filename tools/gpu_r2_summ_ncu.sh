#!/bin/bash
# one ncu --set full capture of k_summarize (47,872 PDFs x 701 grid points = one launch of the standalone API)
mkdir -p gpurun_out
cat > /tmp/summ_one.py <<'PY'
import numpy as np
import frankenz_b200 as fz
rs = np.random.RandomState(1)
n, ng = 47872, 701
zg = np.linspace(0, 7, ng)
mu = rs.uniform(0.1, 6, n); sg = rs.uniform(0.02, 0.5, n)
p = np.exp(-0.5 * ((zg[None, :] - mu[:, None]) / sg[:, None]) ** 2)
p /= p.sum(axis=1)[:, None]
for rep in range(2):
    fz.pdf.pdfs_summarize(p.copy(), zg, rstate=np.random.RandomState(3))
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_summarize -s 0 -c 1 -f -o gpurun_out/prof_summ env PYTHONPATH=. python /tmp/summ_one.py > gpurun_out/prof_summ.log 2>&1
tail -3 gpurun_out/prof_summ.log
