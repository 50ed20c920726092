#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "weight_thresholds or fused" > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ai_pytest.log
tail -4 gpurun_out/r2ai_pytest.log
B="python bench.py --objects 196608 --steps 1 --warmup 1 --no-e2e --no-cpu --no-legs --grid float64"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_finish|k_fuse_collect" -s 4 -c 3 -f -o gpurun_out/prof_finish_r2 $B > gpurun_out/r2ai_ncu.log 2>&1; echo "ncu rc=$?"
