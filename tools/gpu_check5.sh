#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== TC conv tests (one process each)"
: > gpurun_out/tc_conv_tests.log
for id in $(python -m pytest tests/test_gpu_tc.py --co -q 2>/dev/null | grep '::' | grep -E "conv|hifigan"); do
  echo "#### $id" >> gpurun_out/tc_conv_tests.log
  timeout 240 python -m pytest "$id" -q -x -s --timeout 200 2>&1 | grep -E "err|passed|failed|Error|assert" | tail -8 >> gpurun_out/tc_conv_tests.log
done
cat gpurun_out/tc_conv_tests.log | tail -60
echo "== stage timing, TC engine"
timeout 600 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_b64_tc2.log | tail -13
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tee gpurun_out/bench_tc2.log | tail -8 | cut -c1-2500
