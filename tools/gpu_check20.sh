#!/bin/bash
set -u
mkdir -p gpurun_out
for mode in 1 2; do
  echo "== TC tests PAIR=$mode"
  MEGATTS2_TC_PAIR=$mode timeout 400 python -m pytest tests/test_gpu_tc.py -q -x --timeout 120 2>&1 | tail -8 | tee gpurun_out/tc_tests_pair$mode.log
done
for mode in 0 1 2; do
  echo "== stage timing PAIR=$mode"
  MEGATTS2_TC_PAIR=$mode timeout 300 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_pair$mode.log | grep -E "mrte|adm.infer|plm.infer|decode_mel|hifigan|full gpu|samples/s|conv_tc_launch|TOTAL" | tail -9
done
