#!/bin/bash
# round 2, session m: the threshold-filter kNN scan (both forms) + A/B of the C3 step (exact cut on/off, pruning on/off)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2m_gpu.txt
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -k "knn" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -15 gpurun_out/r2m_pytest.log
for form in dot diff; do
  FZB_KNN_FORM=$form timeout 600 python tools/bench_knn.py 1000000 65536 20 25 > gpurun_out/r2m_knn_$form.log 2>&1; echo "knn $form rc=$?"
  tail -4 gpurun_out/r2m_knn_$form.log | cut -c1-260
done
B="python bench.py --objects 1000000 --steps 3 --warmup 3 --no-cpu --no-legs --no-e2e"
timeout 400 $B > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err; echo "bench default rc=$?"
FZB_NO_EXACT_CUT=1 timeout 400 $B > gpurun_out/r2m_bench_nocut.json 2> gpurun_out/r2m_bench_nocut.err; echo "bench nocut rc=$?"
FZB_NO_EXACT_CUT=1 FZB_NO_PRUNE=1 timeout 400 $B > gpurun_out/r2m_bench_nocut_noprune.json 2> gpurun_out/r2m_bench_nocut_noprune.err; echo "bench nocut noprune rc=$?"
timeout 400 $B > gpurun_out/r2m_bench_default2.json 2> gpurun_out/r2m_bench_default2.err; echo "bench default2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_bench_*.json')):
    for line in open(f):
        if line.startswith('{'):
            d=json.loads(line); r=d['roofline']
            print(f, '%.4g'%d['value'], r['ms'], r.get('pass2_pairs_evaluated_frac'), d['clocks'], 'fp32grid', d.get('fp32_rounded_grid',{}).get('ms'))
PY
