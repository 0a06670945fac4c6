#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== TC tests"
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x -s --timeout 200 2>&1 | grep -E "tc err|passed|failed|Error" | tee gpurun_out/tc_tests2.log | tail -15
echo "== reference arm on this box"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tee gpurun_out/bench_reference.log | tail -4
echo "== bench FFMA engine"
timeout 900 python bench.py --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ffma.log | tail -12
echo "== stage timing, TC engine"
MEGATTS2_ENGINE=tc timeout 600 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_b64_tc.log | tail -14
echo "== bench TC engine"
MEGATTS2_ENGINE=tc timeout 900 python bench.py --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_tc.log | tail -12
