#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "knn" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
timeout 600 python tools/bench_knn.py 1000000 65536 20 25 > gpurun_out/r2h_knn.log 2>&1; echo "knn rc=$?"
tail -4 gpurun_out/r2h_knn.log
