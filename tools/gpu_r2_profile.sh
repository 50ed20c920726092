#!/bin/bash
# round 2 profiling session (one B200): ncu --set full of the dominant kernels, launch list of a bench step, sanitizers
mkdir -p gpurun_out
timeout 700 ncu --set full --clock-control none --import-source on -k regex:k_sweep_tc -s 2 -c 2 -f -o gpurun_out/prof_tc_r2 python bench.py --objects 196608 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs --grid float64 > gpurun_out/r2_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --objects 262144 --steps 2 --warmup 1 --no-e2e --no-cpu --no-legs --grid float64 > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_scan -c 1 -f -o gpurun_out/prof_knn_r2 python tools/bench_knn.py 1000000 8192 4 25 > gpurun_out/r2_ncu_knn.log 2>&1; echo "ncu knn rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:k_sweep2 -s 2 -c 1 -f -o gpurun_out/prof_fx1_r2 python bench.py --objects 65536 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs --grid fp32 --lprob '{}' > gpurun_out/r2_ncu_fx1.log 2>&1; echo "ncu fx1 rc=$?"
timeout 600 compute-sanitizer --tool memcheck python tools/tc_small.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck python tools/tc_small.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r2_sanitizer_memcheck.log gpurun_out/r2_sanitizer_racecheck.log
