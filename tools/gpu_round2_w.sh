#!/bin/bash
# GPU call W: deeper weight-tile ring in the halo form (24 stages at C = 32; C = 64: 2 tile buffers, 6 / 13 stages)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x -k "tc or hifigan or e2e or conv or revocode" 2>&1 | tail -3
for hp in 1 2 0; do
  echo "== shapes MEGATTS2_TC_HALO_PAIR=$hp"
  MEGATTS2_TC_HALO_PAIR=$hp timeout 600 python tools/bench_tc_shapes.py --fmt f16x2 --reps 10 --variants 0 --shapes 6,7 2>&1 | grep hifi
done 2>&1 | tee gpurun_out/r2w_tc_shapes.log
for rep in 1 2; do for hp in 1 2; do
  echo "== MEGATTS2_TC_HALO_PAIR=$hp"
  MEGATTS2_TC_HALO_PAIR=$hp timeout 600 python tools/time_stages.py --reps 2 2>&1 | grep -A13 "pass 1" | grep -E "hifigan|full"
done; done 2>&1 | tee gpurun_out/r2w_stages.log
