#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ae_pytest.log
tail -4 gpurun_out/r2ae_pytest.log
timeout 300 python tools/e2e_breakdown.py 1000000 > gpurun_out/r2ae_e2e.log 2>&1; tail -2 gpurun_out/r2ae_e2e.log
timeout 900 python bench.py --no-legs > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ae_bench.json') if l.startswith('{')][-1])
r=d['roofline']
print('value %.4g e2e %.4g (%.3f s) e2e_summ %.4g (%.3f s) frac %.3f' % (d['value'], d['e2e']['value'], d['e2e']['seconds_per_step'], d['e2e_summaries']['value'], d['e2e_summaries']['seconds_per_step'], r['frac']), r['ms'])
PY
