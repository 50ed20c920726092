#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_scan_tc -c 1 -f -o gpurun_out/prof_knn_tc_r2 python tools/bench_knn.py 400000 8192 4 25 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2g_ncu.log
