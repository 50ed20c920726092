#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "knn" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -30 gpurun_out/r2e_pytest.log
timeout 900 python tools/bench_knn.py 1000000 65536 20 25 > gpurun_out/r2e_knn.log 2>&1; echo "knn rc=$?"
tail -12 gpurun_out/r2e_knn.log
