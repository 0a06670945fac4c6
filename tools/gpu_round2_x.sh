#!/bin/bash
# GPU call X2: vocoder-only A/B on one box: committed build vs compact per-tap MMA issue (tc_tap_f16x2)
set -u
mkdir -p gpurun_out
L=megatts2_b200/lib
cp $L/libmegatts2_b200.so $L/new.keep
run() { echo "== $1"; timeout 300 python tools/bench_hifigan.py --reps 10 2>&1 | tail -1; }
for rep in 1 2 3; do
  cp $L/base.keep $L/libmegatts2_b200.so; run "base"
  cp $L/new.keep $L/libmegatts2_b200.so; run "new"
done 2>&1 | tee gpurun_out/r2x2_hifigan_tap_issue_ab.log
cp $L/new.keep $L/libmegatts2_b200.so
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x -k "tc or hifigan or e2e or conv or revocode" 2>&1 | tail -3
