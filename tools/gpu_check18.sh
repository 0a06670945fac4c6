#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== halo tests (each in its own process)"
: > gpurun_out/halo_tests.log
for id in $(python -m pytest tests/test_gpu_tc.py --co -q 2>/dev/null | grep '::' | grep -E "T5000|T5001|T4999|halo"); do
  echo "#### $id" >> gpurun_out/halo_tests.log
  timeout 200 python -m pytest "$id" -q -x -s --timeout 150 2>&1 | grep -E "err|passed|failed|Error|assert" | tail -5 >> gpurun_out/halo_tests.log
done
cat gpurun_out/halo_tests.log | tail -40
echo "== stage timing + trace"
timeout 300 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1l.log | grep -E "pass 1|mrte|adm.infer|plm.infer|decode_mel|hifigan|full gpu|samples/s|_launch|TOTAL|finite"
