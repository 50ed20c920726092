#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --objects 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs --grid float64 > gpurun_out/r2t_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2t_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
for r in rows[1:]:
    print(r[ki][:70], r[gi], '%.3f ms' % (float(r[vi].replace(',',''))/1e6))
PY
