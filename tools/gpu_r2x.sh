#!/bin/bash
# round-end style run: all GPU tests, default bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -6 gpurun_out/r2x_pytest.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$? in $SECONDS s"
tail -c 600 gpurun_out/r2x_bench.err
SECONDS=0
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2x_ref.json 2> gpurun_out/r2x_ref.err; echo "ref rc=$? in $SECONDS s"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2x_bench.json') if l.startswith('{')][-1])
r=d['roofline']
print('value %.4g e2e %.4g e2e_summ %.4g frac %.3f whole %.3f' % (d['value'], d['e2e']['value'], d['e2e_summaries']['value'], r['frac'], r['whole_step_frac']), r['ms'], d['clocks'])
print('cpu', d['cpu_baseline'])
print('fp32grid', d['fp32_rounded_grid']['value'], d['fp32_rounded_grid']['ms'])
print('fx1', json.dumps(d['default_likelihood'])[:700])
print('knn', json.dumps(d['knn'])[:900])
print(open('gpurun_out/r2x_ref.json').read()[:400])
PY
