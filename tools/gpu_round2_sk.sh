#!/bin/bash
# GPU call SK: parity suite on the refactored dispatcher (tc_plan) + the split-K cost-model constants (never swept on a GPU before)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x 2>&1 | tee gpurun_out/r3c_pytest_gpu.log | tail -3
run() {
  echo "== $*"
  env "$@" timeout 300 python tools/ar_step_curve.py --steps $(seq 1 64) --reps 3 --infer --quiet 2>&1 | grep sum_ms
}
{
  run A=0
  run MEGATTS2_TC_SPLITK_MARGIN=1.0
  run MEGATTS2_TC_SPLITK_MARGIN=0.7
  run MEGATTS2_TC_SPLITK_MAX=4
  run MEGATTS2_TC_SPLITK_MAX=16
  run MEGATTS2_TC_SPLITK_MARGIN=1.2 MEGATTS2_TC_SPLITK_MAX=16
  run A=0
} | tee gpurun_out/r3c_splitk_constants.log
