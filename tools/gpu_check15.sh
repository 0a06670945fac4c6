#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== TC tests (2-CTA multicast variant active for big shapes)"
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x --timeout 120 2>&1 | tail -6 | tee gpurun_out/pytest_r1i.log
echo "== parity subset"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "hifigan or e2e or plm or adm or encoder or linearity" 2>&1 | tail -6 | tee -a gpurun_out/pytest_r1i.log
echo "== stage timing + trace (pairs on)"
timeout 300 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1i.log | grep -E "pass 1|mrte|adm.infer|plm.infer|decode_mel|hifigan|full gpu|samples/s|conv_tc_launch|TOTAL"
echo "== stage timing (pairs off)"
MEGATTS2_TC_PAIRS=0 timeout 300 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1i_nopairs.log | grep -E "pass 1|mrte|adm.infer|plm.infer|decode_mel|hifigan|full gpu|samples/s|conv_tc_launch|TOTAL"
