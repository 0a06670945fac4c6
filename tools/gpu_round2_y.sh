#!/bin/bash
# GPU call Y: the AR steps' last rows on the tensor cores (one half-filled tile, split-K) vs the exact FFMA engine
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x 2>&1 | tail -3
for rep in 1 2 3; do for v in 0 1; do
  echo "== MEGATTS2_LAST_ROW_TC=$v"
  MEGATTS2_LAST_ROW_TC=$v timeout 600 python tools/time_stages.py --reps 2 2>&1 | grep -A13 "pass 1" | grep -E "adm|plm|full"
done; done 2>&1 | tee gpurun_out/r2y_last_row_tc_ab.log
