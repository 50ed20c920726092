#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --objects 1000000 --steps 2 --warmup 3 --no-cpu --no-legs --no-e2e --grid both > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2d_bench.err
FZB_NO_PRUNE=1 timeout 900 python bench.py --objects 1000000 --steps 2 --warmup 3 --no-cpu --no-legs --no-e2e --grid fp32 > gpurun_out/r2d_bench_noprune.json 2> gpurun_out/r2d_bench_noprune.err; echo "bench rc=$?"
