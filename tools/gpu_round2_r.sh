#!/bin/bash
# GPU call R: halo form as a CTA pair (cta_group::2, M = 256 per MMA) for the C = 32 / 64 vocoder stages: parity + A/B
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x 2>&1 | tee gpurun_out/r2r_pytest.log | tail -5
for hp in 1 0 1 0; do
  echo "== MEGATTS2_TC_HALO_PAIR=$hp"
  MEGATTS2_TC_HALO_PAIR=$hp timeout 600 python tools/time_stages.py --reps 2 2>&1 | grep -A13 "pass 1" | grep -E "hifigan|full" | tee -a gpurun_out/r2r_stages_halo_pair.log
done
for hp in 1 0; do
  echo "== shapes MEGATTS2_TC_HALO_PAIR=$hp"
  MEGATTS2_TC_HALO_PAIR=$hp timeout 600 python tools/bench_tc_shapes.py --fmt f16x2 --reps 10 --variants 0 --shapes 6,7 2>&1 | tee -a gpurun_out/r2r_tc_shapes.log
done
