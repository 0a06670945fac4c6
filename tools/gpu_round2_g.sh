#!/bin/bash
# GPU call G: ncu launch list of one step on the uniform-issue build + bench line
set -u
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2g_launches.csv python tools/profile_step.py --batch 64 > gpurun_out/r2g_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2g_launches.csv)"
python tools/summarize_launches.py gpurun_out/r2g_launches.csv gpurun_out/r2g_launch_summary.md | head -30
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -4 gpurun_out/r2g_bench.err; cut -c1-400 gpurun_out/r2g_bench.json
