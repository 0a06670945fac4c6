#!/bin/bash
# GPU call D: halo form of the small-channel convs: parity under both base-offset modes, then timing; attention A/B sweep
set -u
mkdir -p gpurun_out
for bo in 1 0; do
  echo "== conv tests HALO_BO=$bo"
  MEGATTS2_TC_HALO_BO=$bo timeout 900 python -m pytest tests/test_gpu_tc.py -q --timeout 600 -p no:randomly -k "conv_tc_vs_fp64 or hifigan" 2>&1 | tee gpurun_out/r2d_pytest_bo$bo.log | tail -12
done
echo "== attention sweep"
for m in 0 1; do MEGATTS2_ATTN_TC=$m MEGATTS2_ATTN_TC_MIN=1 timeout 300 python tools/bench_attention.py; done 2>&1 | tee gpurun_out/r2d_attention_ab.log
for h in 1 0; do
  echo "== stages HALO=$h"
  MEGATTS2_TC_HALO=$h timeout 600 python tools/time_stages.py --reps 2 2>&1 | tee gpurun_out/r2d_stages_halo$h.log | grep -A12 "pass 1"
done
