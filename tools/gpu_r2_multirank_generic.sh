#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
timeout 600 python tests/scripts/bench_generic_modes.py > gpurun_out/r2_generic.log 2>&1; echo "generic rc=$?"; tail -8 gpurun_out/r2_generic.log | cut -c1-250
