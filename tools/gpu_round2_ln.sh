#!/bin/bash
# GPU call LN: modelled cost of a split-K reduction that replaces the LayerNorm launch (9000 cycles = the plain reduction's, vs 2000):
# per-step curves A/B, then the parity suite under the cheaper setting
set -u
mkdir -p gpurun_out
for c in 9000 2000 9000 2000; do
  echo "== MEGATTS2_TC_SPLITK_LNCOST=$c"
  MEGATTS2_TC_SPLITK_LNCOST=$c timeout 200 python tools/ar_step_curve.py --steps $(seq 1 64) --reps 3 --infer 2>&1 | grep -v Warning
done > gpurun_out/r3d_lncost_curves.log 2>&1
grep -E "==|sum_ms" gpurun_out/r3d_lncost_curves.log
MEGATTS2_TC_SPLITK_LNCOST=2000 timeout 400 python -m pytest tests -m gpu -q --timeout 300 -p no:randomly -x 2>&1 | tee gpurun_out/r3d_pytest_gpu_lncost2000.log | tail -3
