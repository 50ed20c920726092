#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_calibrate.py tests/test_gpu_scale.py -m gpu -q -s -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
grep -E "max L1|passed|failed|FAILED|Error|error|assert|rc=" gpurun_out/r2p_pytest.log | tail -25
B="python bench.py --objects 1000000 --steps 3 --warmup 3 --no-cpu --no-legs --no-e2e"
timeout 400 $B > gpurun_out/r2p_bench_fused.json 2> gpurun_out/r2p_bench_fused.err; echo "bench fused rc=$?"
FZB_NO_FUSE=1 timeout 400 $B > gpurun_out/r2p_bench_nofuse.json 2> gpurun_out/r2p_bench_nofuse.err; echo "bench nofuse rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2p_bench_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); r=d['roofline']
            print(f, '%.4g'%d['value'], r['ms'], r.get('pass2_pairs_evaluated_frac'), 'fp32grid', d.get('fp32_rounded_grid',{}).get('ms'))
PY
tail -3 gpurun_out/r2p_bench_fused.err
FZB_KNN_FORM=dot timeout 600 python tools/bench_knn.py 1000000 65536 20 25 > gpurun_out/r2p_knn.log 2>&1; tail -3 gpurun_out/r2p_knn.log | cut -c1-260
timeout 600 python tools/knn_e2e_breakdown.py > gpurun_out/r2p_knn_e2e.log 2>&1; tail -3 gpurun_out/r2p_knn_e2e.log
