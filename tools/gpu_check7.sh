#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15 | tee gpurun_out/pytest_r1c.log
echo "== stage timing"
timeout 600 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1c.log | tail -13
echo "== mel C2"
timeout 300 python tools/bench_mel.py 10000 2>&1 | tee gpurun_out/bench_mel_c2_r1c.log | tail -2
echo "== bench"
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tee gpurun_out/bench_r1c.log | tail -6 | cut -c1-3000
