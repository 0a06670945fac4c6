#!/bin/bash
# GPU call TM: cost-model dispatch of the dense layers (tile width / K split / pairing by modelled launch cost) vs the fill
# heuristics: parity suite on the new default, per-step curves and graph-replayed infer() under both, full step A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x 2>&1 | tee gpurun_out/r2tm_pytest_gpu.log | tail -4
for m in 1 0 1 0; do
  echo "== MEGATTS2_TC_MODEL=$m"
  MEGATTS2_TC_MODEL=$m timeout 300 python tools/ar_step_curve.py --steps $(seq 1 64) --reps 4 --infer 2>&1 | grep -v Warning
done | tee gpurun_out/r2tm_ar_curves.log | grep -E "==|sum_ms"
for m in 1 0 1 0; do
  echo "== MEGATTS2_TC_MODEL=$m"
  MEGATTS2_TC_MODEL=$m timeout 300 python tools/sweep_overlap.py --configs 0 --steps 5 2>&1 | grep config
done | tee gpurun_out/r2tm_step_ab.log
