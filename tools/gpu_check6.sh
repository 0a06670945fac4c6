#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_r1b_b64.csv python tools/profile_step.py --batch 64 > gpurun_out/ncu_launch2.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r1b_b64.csv gpurun_out/launch_summary_r1b_b64.md | tail -40
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_bf16x3 -s 30 -c 8 \
   -o gpurun_out/prof_conv_tc -f python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/ncu_full4.log 2>&1
ls -la gpurun_out/prof_conv_tc.ncu-rep
