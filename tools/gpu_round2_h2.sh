#!/bin/bash
# GPU call H2: LayerNorm fused into the split-K reduction of the preceding dense layer: parity, A/B in the step
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x 2>&1 | tail -4
for rep in 1 2 3; do for v in 0 1; do
  echo "== MEGATTS2_LN_FUSE=$v"
  MEGATTS2_LN_FUSE=$v timeout 600 python tools/time_stages.py --reps 3 2>&1 | grep -A13 "pass 2" | grep -E "adm|plm|launches|full"
done; done 2>&1 | tee gpurun_out/r2h2_ln_fuse_ab.log
