#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_knn -c 120 --csv --log-file gpurun_out/r2n_knn_launches.csv python tools/bench_knn.py 1000000 32768 20 25 > gpurun_out/r2n_knn.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2n_knn_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:]:
    print(r[ki][:60], r[gi], r[vi])
PY
