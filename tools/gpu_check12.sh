#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_hifigan.csv python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/ncu_l3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_hifigan.csv gpurun_out/launch_summary_hifigan.md | tail -14
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_bf16x3 -s 72 -c 3 \
   -o gpurun_out/prof_conv_tc_c64 -f python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/ncu_full5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_bf16x3 -s 108 -c 3 \
   -o gpurun_out/prof_conv_tc_c32 -f python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/ncu_full6.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
