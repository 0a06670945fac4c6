#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== targeted tests"
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "hifigan or conv or e2e or plm or adm or tc or encoder" 2>&1 | tail -8 | tee gpurun_out/pytest_r1h.log
echo "== stage timing + trace"
timeout 600 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1h.log | tail -22
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_hifigan2.csv python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/ncu_l4.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_hifigan2.csv gpurun_out/launch_summary_hifigan2.md | tail -12
