#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
grep -E "max L1|passed|failed|FAILED|rc=" gpurun_out/r2k_pytest.log | tail -12
timeout 900 python bench.py --objects 1000000 --steps 3 --warmup 3 --cpu-objects-per-core 8 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2k_bench.err
python tools/measure_peaks.py > gpurun_out/peaks_r2.json 2>/dev/null; echo "peaks rc=$?"
