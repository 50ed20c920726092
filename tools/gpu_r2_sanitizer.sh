#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -k "knn" > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -4 gpurun_out/r2_pytest.log
timeout 900 compute-sanitizer --tool memcheck python tools/tc_small.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/r2_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/tc_small.py > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r2_racecheck.log
