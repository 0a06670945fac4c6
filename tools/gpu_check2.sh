#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== stage timing B=8"
timeout 300 python tools/time_stages.py --batch 8 --reps 1 2>&1 | tee gpurun_out/stages_b8.log | tail -20
echo "== stage timing B=64"
timeout 600 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_b64.log | tail -30
echo "== TC tests (one process each)"
: > gpurun_out/tc_tests.log
for id in $(python -m pytest tests/test_gpu_tc.py --co -q 2>/dev/null | grep '::'); do
  echo "#### $id" | tee -a gpurun_out/tc_tests.log
  timeout 240 python -m pytest "$id" -q -x -s --timeout 200 2>&1 | tail -15 >> gpurun_out/tc_tests.log
  tail -3 gpurun_out/tc_tests.log
done
