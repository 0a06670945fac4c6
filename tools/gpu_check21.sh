#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== parity tests (attention, encoder, plm, adm, e2e)"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 2>&1 | tail -5 | tee gpurun_out/parity_r1n.log
echo "== TC tests"
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x --timeout 300 2>&1 | tail -3 | tee gpurun_out/tc_tests_r1n.log
echo "== stage timing"
timeout 300 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1n.log | grep -E "pass 1|mrte|adm.infer|plm.infer|decode_mel|hifigan|full gpu|samples/s|_launch|layernorm|TOTAL|finite" | tail -24
