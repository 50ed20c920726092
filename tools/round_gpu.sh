# round-end style GPU run: tests, bench (1 GPU), reference arm, ncu launch list
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r1_pytest_gpu.log; tail -4 gpurun_out/r1_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_r1_tc.json 2> gpurun_out/bench_r1_tc.err; echo "bench exit=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; echo "ref exit=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1_tc.csv python bench.py --objects 262144 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches_tc.log 2>&1; echo "ncu exit=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r1_tc.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g frac %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']), d['roofline']['ms'], d['roofline']['kernel'], d['cpu_baseline']['value'], d['clocks'])
"
cat gpurun_out/bench_r1_ref.json | cut -c1-300
timeout 400 python tests/scripts/bench_c1.py > gpurun_out/r1_c1_c2.log 2>&1; tail -8 gpurun_out/r1_c1_c2.log | cut -c1-330
timeout 500 python tools/bench_knn.py 1000000 65536 20 25 > gpurun_out/r1_knn.log 2>&1; tail -7 gpurun_out/r1_knn.log | cut -c1-250
