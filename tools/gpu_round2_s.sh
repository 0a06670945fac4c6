#!/bin/bash
# GPU call S: halo-pair modes in the step (0 off, 1 C64 only, 2 both), alternating
set -u
mkdir -p gpurun_out
for rep in 1 2; do for hp in 0 1 2; do
  echo "== MEGATTS2_TC_HALO_PAIR=$hp"
  MEGATTS2_TC_HALO_PAIR=$hp timeout 600 python tools/time_stages.py --reps 2 2>&1 | grep -A13 "pass 1" | grep -E "hifigan|full"
done; done 2>&1 | tee gpurun_out/r2s_stages_halo_pair_modes.log
