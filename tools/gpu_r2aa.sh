#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -k "knn" > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log
tail -5 gpurun_out/r2aa_pytest.log
timeout 600 python tools/knn_e2e_breakdown.py > gpurun_out/r2aa_knn_e2e.log 2>&1; tail -2 gpurun_out/r2aa_knn_e2e.log
FZB_KNN_NO_SHARE=1 timeout 600 python tools/knn_e2e_breakdown.py > gpurun_out/r2aa_knn_e2e_noshare.log 2>&1; tail -1 gpurun_out/r2aa_knn_e2e_noshare.log
timeout 900 python bench.py > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2aa_bench.json') if l.startswith('{')][-1])
r=d['roofline']
print('value %.4g e2e %.4g (%.3f s) e2e_summ %.4g (%.3f s) frac %.3f' % (d['value'], d['e2e']['value'], d['e2e']['seconds_per_step'], d['e2e_summaries']['value'], d['e2e_summaries']['seconds_per_step'], r['frac']), r['ms'])
print('knn', json.dumps(d['knn'])[:700])
PY
