#!/bin/bash
# GPU call G2: CUDA-graph replay of the AR drivers: parity, A/B in the step
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x 2>&1 | tail -4
for rep in 1 2 3; do for v in 0 1; do
  echo "== MEGATTS2_GRAPHS=$v"
  MEGATTS2_GRAPHS=$v timeout 600 python tools/time_stages.py --reps 3 2>&1 | grep -A13 "pass 2" | grep -E "adm|plm|full"
done; done 2>&1 | tee gpurun_out/r2g2_graphs_ab.log
