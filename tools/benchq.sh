timeout 300 python bench.py --objects 262144 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/tc_bench.log 2>&1; echo exit=$? >> gpurun_out/tc_bench.log
python - <<PY
import json
l=[x for x in open("gpurun_out/tc_bench.log") if x.startswith("{")]
if not l: print(open("gpurun_out/tc_bench.log").read()[-2000:])
else:
    d=json.loads(l[-1]); print(d["value"], d["roofline"]["ms"], d["roofline"]["objects_routed_to_fp64"], d["roofline"]["pairs_per_s_kernel"], d["roofline"].get("fit_only_pairs_per_s"))
PY
