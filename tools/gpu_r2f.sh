#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_knn.py 1000000 16384 20 25 > gpurun_out/r2f_knn.log 2>&1; echo "knn rc=$?"
tail -4 gpurun_out/r2f_knn.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_knn_launches.csv python tools/bench_knn.py 1000000 16384 20 25 > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2f_knn_launches.csv 2>/dev/null | head -30
