#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --objects 196608 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs --grid fp32"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tc -s 3 -c 3 -f -o gpurun_out/prof_fuse_r2 $B > gpurun_out/r2q_ncu_fuse.log 2>&1; echo "ncu fuse rc=$?"
FZB_NO_FUSE=1 timeout 700 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tc -s 2 -c 1 -f -o gpurun_out/prof_nofuse_r2 $B > gpurun_out/r2q_ncu_nofuse.log 2>&1; echo "ncu nofuse rc=$?"
B1="python bench.py --objects 1000000 --steps 2 --warmup 2 --no-cpu --no-legs --no-e2e --grid fp32"
for band in 0.003 0.03; do
  FZB_FUSE_BAND=$band timeout 300 $B1 > gpurun_out/r2q_bench_band$band.json 2> gpurun_out/r2q_bench_band$band.err; echo "band $band rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2q_bench_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); r=d['roofline']
            print(f, '%.4g'%d['value'], r['ms'], r.get('pass2_pairs_evaluated_frac'))
PY
