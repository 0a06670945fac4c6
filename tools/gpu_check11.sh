#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== targeted tests"
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "hifigan or conv or e2e or plm or adm or tc or encoder" 2>&1 | tail -8 | tee gpurun_out/pytest_r1g.log
echo "== stage timing + trace"
timeout 600 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1g.log | tail -34
