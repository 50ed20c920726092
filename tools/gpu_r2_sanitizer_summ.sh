#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/summ_small.py > gpurun_out/r2_memcheck_summ.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r2_memcheck_summ.log
timeout 900 compute-sanitizer --tool racecheck python tools/summ_small.py > gpurun_out/r2_racecheck_summ.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r2_racecheck_summ.log
