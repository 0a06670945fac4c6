#!/bin/bash
# GPU call V: transpose-free planes-only epilogue of the halo form (first conv of a ResBlock step): parity + A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x -k "tc or hifigan or e2e or conv or revocode" 2>&1 | tail -3
for rep in 1 2; do for d in 0 1; do
  echo "== MEGATTS2_TC_DIRECT=$d"
  MEGATTS2_TC_DIRECT=$d timeout 600 python tools/time_stages.py --reps 2 2>&1 | grep -A13 "pass 1" | grep -E "hifigan|full"
done; done 2>&1 | tee gpurun_out/r2v_direct_planes_ab.log
