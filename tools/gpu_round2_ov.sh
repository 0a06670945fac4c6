#!/bin/bash
# GPU call OV: overlapped prompt re-vocode with complementary SM budgets / no PDL on the vocoder stream / no pairs on the AR
# stream: parity (whole GPU suite), sweep of (budget, share, start), AR per-step curve
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x 2>&1 | tee gpurun_out/r2ov_pytest_gpu.log | tail -6
timeout 600 python tools/sweep_overlap.py 2>&1 | tee gpurun_out/r2ov_overlap_sweep.log
timeout 300 python tools/ar_step_curve.py 2>&1 | grep -v Warning | tee gpurun_out/r2ov_ar_step_curve.log
