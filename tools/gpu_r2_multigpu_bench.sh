#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench$N.json 2> gpurun_out/r2_bench$N.err; echo "bench$N rc=$?"
tail -c 300 gpurun_out/r2_bench$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads([l for l in open('gpurun_out/r2_bench%s.json'%N) if l.startswith('{')][-1])
print('value %.4g e2e %.4g e2e_summ %.4g' % (d['value'], d['e2e']['value'], d['e2e_summaries']['value']), d['roofline']['ms'])
ms=d['model_sharded']; print('model_sharded %.4g eff %.3f' % (ms['value'], ms['efficiency_vs_independent_gpus']), ms['parity_vs_unsharded'])
k=d['knn']; print('knn %.4g e2e %.4g' % (k['distance_evaluations_per_s'], k['e2e_queries_per_s']), k['index_check'], k['redo_searches'])
print('fx1 %.4g' % d['default_likelihood']['value'])
PY
