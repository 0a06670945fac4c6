#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_r1c_b64.csv python tools/profile_step.py --batch 64 > gpurun_out/ncu_launch3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r1c_b64.csv gpurun_out/launch_summary_r1c_b64.md | tail -32
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_bf16x3 -s 3000 -c 4 \
   -o gpurun_out/prof_tc_plm -f python tools/profile_step.py --batch 64 --stage plm > gpurun_out/ncu_full7.log 2>&1
ls -la gpurun_out/prof_tc_plm.ncu-rep
