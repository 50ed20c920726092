#!/bin/bash
# round 2, first GPU session: parity tests, short bench (both grids + kNN leg), ncu capture of the kNN scan
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --objects 262144 --steps 2 --warmup 3 --cpu-objects-per-core 4 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_scan -c 1 -f -o gpurun_out/prof_knn_r2a python tools/bench_knn.py 200000 16384 4 25 > gpurun_out/r2a_ncu_knn.log 2>&1; echo "ncu rc=$?"
