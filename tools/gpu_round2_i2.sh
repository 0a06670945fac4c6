#!/bin/bash
# GPU call I2: tiled conv_post kernel: parity (vocoder tests) + vocoder timing
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x -k "hifigan or e2e or revocode or conv or tc" 2>&1 | tail -3
timeout 300 python tools/bench_hifigan.py --reps 10 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"conv_cout1" --csv python tools/profile_step.py --batch 64 --stage hifigan 2>&1 | grep -i "conv_cout1" | cut -c1-300 | tail -3
