#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_calibrate.py tests/test_gpu_scale.py -m gpu -q -s > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ak_pytest.log
grep -E "max L1|passed|failed|FAILED|Error|error|assert|rc=" gpurun_out/r2ak_pytest.log | tail -25
B="python bench.py --objects 1000000 --steps 3 --warmup 3 --no-cpu --no-legs --no-e2e"
timeout 400 $B > gpurun_out/r2ak_bench_fused.json 2> gpurun_out/r2ak_bench_fused.err; echo "bench fused rc=$?"
FZB_FUSE_BAND=0.005 timeout 400 $B --grid fp32 > gpurun_out/r2ak_bench_band005.json 2> gpurun_out/r2ak_bench_band005.err; echo "bench band rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2ak_bench_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); r=d['roofline']
            print(f, '%.4g'%d['value'], r['ms'], r.get('pass2_pairs_evaluated_frac'), 'fp32grid', d.get('fp32_rounded_grid',{}).get('ms'))
PY
tail -3 gpurun_out/r2ak_bench_fused.err
