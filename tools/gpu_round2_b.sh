#!/bin/bash
# GPU call B: where does the f16x2 step spend its time (ncu launch list), per-shape tap-GEMM rates, CPU-arm thread sweep
set -u
mkdir -p gpurun_out
echo "== shapes f16x2"
timeout 600 python tools/bench_tc_shapes.py --fmt f16x2 --reps 10 2>&1 | tee gpurun_out/r2b_tc_shapes_f16x2.log | tail -60
echo "== ncu launch list (one step, f16x2)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2b_launches.csv python tools/profile_step.py --batch 64 > gpurun_out/r2b_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2b_launches.csv)"
python tools/summarize_launches.py gpurun_out/r2b_launches.csv gpurun_out/r2b_launch_summary.md | head -50
echo "== CPU arm sweep (workers x threads), one utterance per worker"
for wt in "1 32" "1 64" "2 32" "2 64" "4 16" "4 32" "8 16"; do
  set -- $wt
  timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --cpu-workers $1 --cpu-threads $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('workers x threads = $1 x $2:', d['value'], 'samples/s; single-stream', d['cpu_baseline']['single_stream']['value'])"
done 2>&1 | tee gpurun_out/r2b_cpu_sweep.log
