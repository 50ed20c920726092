#!/bin/bash
mkdir -p gpurun_out
for cfg in "default" "FZB_E2E_GEOM=1" "FZB_PINNED_POOL_BYTES=25769803776" "FZB_PINNED_POOL_BYTES=25769803776 FZB_E2E_HALVING=1"; do
  echo "== $cfg"
  if [ "$cfg" = "default" ]; then timeout 300 python tools/e2e_breakdown.py 1000000 2>&1 | tail -2; else env $cfg timeout 300 python tools/e2e_breakdown.py 1000000 2>&1 | tail -2; fi
done > gpurun_out/r2af_e2e.log 2>&1
cat gpurun_out/r2af_e2e.log | cut -c1-250
