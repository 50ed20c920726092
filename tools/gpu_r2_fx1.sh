#!/bin/bash
# default-likelihood leg: parity tests, phases with the fused single pass of the packed sweep on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_gpu_multirank.py -m gpu -q -x 2>&1 | tail -12
timeout 300 python tools/bench_fx1.py > gpurun_out/r2_fx1.log 2>&1; tail -1 gpurun_out/r2_fx1.log
FZB_NO_FUSE=1 timeout 300 python tools/bench_fx1.py > gpurun_out/r2_fx1_nofuse.log 2>&1; tail -1 gpurun_out/r2_fx1_nofuse.log
