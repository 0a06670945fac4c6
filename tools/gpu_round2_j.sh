#!/bin/bash
# GPU call J: prompt re-vocode on a side stream with an SM budget: correctness + sweep of the budget
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q --timeout 600 -p no:randomly -k "e2e or revocode or hifigan" 2>&1 | tail -5
for s in 0 40 56 64 74 96; do
  echo "== REVOCODE_SMS=$s"
  MEGATTS2_REVOCODE_SMS=$s timeout 600 python tools/time_stages.py --reps 2 2>&1 | grep -E "full gpu_step|samples/s"
done 2>&1 | tee gpurun_out/r2j_revocode_overlap_sweep.log
