#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --objects 1000000 --steps 3 --warmup 3 --no-cpu --no-legs --no-e2e --grid float64"
for cfg in "FZB_FUSE_BAND=0.004" "FZB_FUSE_BAND=0.009" "FZB_TC_MLO_SNR=24" "FZB_TC_MLO_SNR=48" "FZB_TC_MLO_SNR=64"; do
  env $cfg timeout 300 $B > gpurun_out/r2ah_$cfg.json 2> /dev/null
  python - "$cfg" <<'PY'
import json,sys
cfg=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/r2ah_%s.json'%cfg) if l.startswith('{')][-1]); r=d['roofline']
print(cfg, '%.4g'%d['value'], {k:round(v,1) for k,v in r['ms'].items()}, 'fused frac %.3f' % r['objects_completed_by_the_fused_pass_frac'], 'pass2 frac %.3f' % r['pass2_pairs_evaluated_frac'])
PY
done
