#!/bin/bash
# Final evidence of round 2 (session 3 build: launch policy, cost-chosen tile width) on one box: full GPU suite, smoke, bench (default line), reference arm, C2, stage times,
# ncu launch list of one step (per-launch durations) and the SASS summary inputs
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly 2>&1 | tee gpurun_out/r3a_pytest_gpu.log | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r3a_bench_n1.json 2> gpurun_out/r3a_bench_n1.err; tail -4 gpurun_out/r3a_bench_n1.err; cut -c1-300 gpurun_out/r3a_bench_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3a_bench_reference.json 2> gpurun_out/r3a_bench_reference.err; cut -c1-400 gpurun_out/r3a_bench_reference.json
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/r3a_bench_c2.json 2> gpurun_out/r3a_bench_c2.err; cut -c1-200 gpurun_out/r3a_bench_c2.json
timeout 600 python tools/time_stages.py --reps 3 2>&1 | tee gpurun_out/r3a_stage_times.log | grep -A13 "pass 2"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r3a_launches.csv python tools/profile_step.py --batch 64 > gpurun_out/r3a_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r3a_launches.csv)"
python tools/summarize_launches.py gpurun_out/r3a_launches.csv gpurun_out/r3a_launch_summary.md | head -24
rm -f gpurun_out/r3a_launches.csv
for m in 1 0 2 1 0 2; do
  echo "== MEGATTS2_TC_MODEL=$m"
  MEGATTS2_TC_MODEL=$m timeout 300 python tools/ar_step_curve.py --steps $(seq 1 64) --reps 4 --infer 2>&1 | grep -v Warning
done | tee gpurun_out/r3a_ar_curves.log | grep -E "==|sum_ms"
