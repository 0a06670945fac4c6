#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_r1d.log
echo "== stage timing + trace"
timeout 600 python tools/time_stages.py --batch 64 --reps 2 2>&1 | tee gpurun_out/stages_r1d.log | tail -45
