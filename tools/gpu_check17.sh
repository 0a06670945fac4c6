#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "hifigan or e2e or plm or adm or encoder or tc or conv" 2>&1 | tail -6 | tee gpurun_out/pytest_r1k.log
echo "== stage timing + trace"
timeout 300 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1k.log | grep -E "pass 1|mrte|adm.infer|plm.infer|decode_mel|hifigan|full gpu|samples/s|_launch|layernorm|launch_cfg|conv1d_ffma|conv_tc |TOTAL"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_hifigan3.csv python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/ncu_l5.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_hifigan3.csv gpurun_out/launch_summary_hifigan3.md | tail -10
