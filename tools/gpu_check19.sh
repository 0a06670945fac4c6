#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== TC tests"
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x --timeout 300 2>&1 | tail -5 | tee gpurun_out/tc_tests_r1m.log
echo "== stage timing BN256=0"
MEGATTS2_TC_BN256=0 timeout 300 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1m_bn128.log | grep -E "pass 1|mrte|adm.infer|plm.infer|decode_mel|hifigan|full gpu|samples/s|finite"
echo "== stage timing BN256=1"
timeout 300 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1m.log | grep -E "pass 1|mrte|adm.infer|plm.infer|decode_mel|hifigan|full gpu|samples/s|_launch|TOTAL|finite"
