#!/bin/bash
# profiles of the round-2 kernels at HEAD + e2e breakdown
mkdir -p gpurun_out
timeout 300 python tools/e2e_breakdown.py 1000000 > gpurun_out/r2z_e2e.log 2>&1; tail -3 gpurun_out/r2z_e2e.log
B="python bench.py --objects 196608 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs --grid float64"
# launches per step: prepass A, prepass B, fused A, seeded B, pass 2  -> skip the warm-up (2 steps x 5), take 5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tc -s 10 -c 5 -f -o gpurun_out/prof_tc_r2b $B > gpurun_out/r2z_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_filter -s 3 -c 1 -f -o gpurun_out/prof_knn_r2b python tools/bench_knn.py 1000000 16384 4 25 > gpurun_out/r2z_ncu_knn.log 2>&1; echo "ncu knn rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --objects 262144 --steps 2 --warmup 1 --no-e2e --no-cpu --no-legs --grid float64 > gpurun_out/r2z_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
