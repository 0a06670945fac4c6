#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_tc.py -q --timeout 600 -p no:randomly -k "train or hifigan or oracle_autograd or functions" 2>&1 | tee gpurun_out/r2h2_pytest.log | tail -12
MEGATTS2_PDL=1 timeout 600 python tools/time_stages.py --reps 2 2>&1 | tee gpurun_out/r2h2_stages.log | grep -A13 "pass 1"
bash tools/gpu_round2_i.sh
