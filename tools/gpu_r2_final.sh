#!/bin/bash
# round-2 closing run on one B200: all GPU tests, default bench + reference arm, refreshed kNN profile and launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_filter -s 7 -c 1 -f -o gpurun_out/prof_knn_r2c python tools/bench_knn.py 1000000 16384 4 25 > gpurun_out/r2f_ncu_knn.log 2>&1; echo "ncu knn rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 170 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --objects 262144 --steps 2 --warmup 1 --no-e2e --no-cpu --no-legs --grid float64 > gpurun_out/r2f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_knn -c 40 --csv --log-file gpurun_out/r2f_knn_launches.csv python tools/bench_knn.py 1000000 65536 20 25 > gpurun_out/r2f_knn_launches.log 2>&1; echo "ncu knn launches rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_bench.json') if l.startswith('{')][-1])
r=d['roofline']
print('value %.4g e2e %.4g (%.3f s) e2e_summ %.4g frac %.3f whole %.3f' % (d['value'], d['e2e']['value'], d['e2e']['seconds_per_step'], d['e2e_summaries']['value'], r['frac'], r['whole_step_frac']), r['ms'], d['clocks'])
print('cpu', d['cpu_baseline']['value'], 'fp32grid', d['fp32_rounded_grid']['value'], 'fx1', d['default_likelihood']['value'], 'knn', d['knn']['distance_evaluations_per_s'], d['knn']['e2e_queries_per_s'])
PY
