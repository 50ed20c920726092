#!/bin/bash
# summaries kernel: parity tests, then the device time of k_summarize on 262,144 PDFs (701 grid points)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_summarize.py -x -q 2>&1 | tail -5
cat > /tmp/summ_time.py <<'PY'
import numpy as np, time
import frankenz_b200 as fz
rs = np.random.RandomState(1)
n, ng = 262144, 701
zg = np.linspace(0, 7, ng)
mu = rs.uniform(0.1, 6, n); sg = rs.uniform(0.02, 0.5, n)
p = np.exp(-0.5 * ((zg[None, :] - mu[:, None]) / sg[:, None]) ** 2)
p /= p.sum(axis=1)[:, None]
for rep in range(3):
    t = time.time()
    res = fz.pdf.pdfs_summarize(p.copy(), zg, rstate=np.random.RandomState(3))
    dt = time.time() - t
    from frankenz_b200 import _engine
    st = _engine.last_stats() if hasattr(_engine, "last_stats") else None
    print("pdfs_summarize %d x %d: wall %.3f s" % (n, ng, dt), st)
PY
PYTHONPATH=. timeout 300 python /tmp/summ_time.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_summarize -c 6 --csv --log-file gpurun_out/summ_launches.csv env PYTHONPATH=. python /tmp/summ_time.py > /dev/null 2>&1
grep k_summarize gpurun_out/summ_launches.csv | awk -F'","' '{print $5, $(NF)}' | head
