#!/bin/bash
# default-likelihood leg: phases; parity tests of the float64 sweep route
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/bench_fx1.py > gpurun_out/r2_fx1.log 2>&1; tail -2 gpurun_out/r2_fx1.log
