#!/bin/bash
# default-likelihood leg: phases and launch list
mkdir -p gpurun_out
timeout 300 python tools/bench_fx1.py > gpurun_out/r2_fx1.log 2>&1; tail -3 gpurun_out/r2_fx1.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/r2_fx1_launches.csv python tools/bench_fx1.py > gpurun_out/r2_fx1_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_fx1_launches.csv'))]
hdr=[r for r in rows if 'Kernel Name' in r][0]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
data=[r for r in rows if len(r)>10 and r[0].isdigit()]
preps=[i for i,r in enumerate(data) if 'k_prep_objects' in r[ki]]
for r in data[preps[1]:preps[2]]:
    print(r[ki].replace('<unnamed>::','').replace('void ','')[:70], r[gi], '%.3f ms' % (float(r[vi].replace(',',''))/1e6))
PY
timeout 300 python tests/scripts/bench_c1.py 2>&1 | head -3 | cut -c1-250
