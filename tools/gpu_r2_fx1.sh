#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2am_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2am_pytest.log
tail -6 gpurun_out/r2am_pytest.log
timeout 900 python bench.py --no-e2e --no-cpu --grid float64 --steps 2 > gpurun_out/r2am_bench.json 2> gpurun_out/r2am_bench.err; echo "bench rc=$?"
FZB_NO_PRUNE=1 timeout 900 python bench.py --no-e2e --no-cpu --grid float64 --steps 2 > gpurun_out/r2am_bench_noprune.json 2> gpurun_out/r2am_bench_noprune.err; echo "bench noprune rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2am_bench.json','gpurun_out/r2am_bench_noprune.json'):
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    x=d['default_likelihood']
    print(f, 'fx1 %.4g pairs/s, %.1f ms/step, fp64 objs %d' % (x['value'], x['ms_per_step'], x['objects_routed_to_fp64']), 'headline %.4g' % d['value'])
PY
