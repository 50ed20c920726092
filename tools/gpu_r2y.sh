#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2y_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -4 gpurun_out/r2y_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2y_bench2.json 2> gpurun_out/r2y_bench2.err; echo "bench2 rc=$?"
tail -c 400 gpurun_out/r2y_bench2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2y_bench2.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g' % (d['value'], d['e2e']['value']), d['roofline']['ms'])
print('model_sharded', json.dumps(d['model_sharded'])[:1500])
print('knn', json.dumps(d['knn'])[:500])
PY
