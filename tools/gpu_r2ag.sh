#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --objects 1000000 --steps 3 --warmup 3 --no-cpu --no-legs --no-e2e"
timeout 400 $B > gpurun_out/r2ag_base.json 2> gpurun_out/r2ag_base.err; echo "base rc=$?"
FZB_LIB_PATH=$PWD/frankenz_b200/lib/libfzb200_exp.so timeout 400 $B > gpurun_out/r2ag_exp.json 2> gpurun_out/r2ag_exp.err; echo "exp rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2ag_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); r=d['roofline']
            print(f, '%.4g'%d['value'], r['ms'], 'fp32grid', d.get('fp32_rounded_grid',{}).get('ms'))
PY
