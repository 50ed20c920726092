#!/bin/bash
# end-to-end breakdown of BruteForce.fit_predict (1M objects): recycled output buffers on / off, chunk schedules
mkdir -p gpurun_out
for cfg in "default" "FZB_HOST_POOL_BYTES=0" "FZB_E2E_GEOM=1" "FZB_E2E_CHUNK=524288"; do
  echo "== $cfg"
  if [ "$cfg" = "default" ]; then timeout 300 python tools/e2e_breakdown.py 1000000 2>&1 | tail -2; else env $cfg timeout 300 python tools/e2e_breakdown.py 1000000 2>&1 | tail -2; fi
done > gpurun_out/r2_e2e.log 2>&1
cat gpurun_out/r2_e2e.log | cut -c1-250
