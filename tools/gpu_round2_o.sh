#!/bin/bash
# GPU call O: leaner epilogue (mixed-radix tile walk, hoisted address bases, halo-form constants): full parity, stage times, shapes
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x 2>&1 | tee gpurun_out/r2o_pytest.log | tail -5
for rep in 1 2; do timeout 600 python tools/time_stages.py --reps 2 2>&1 | tee -a gpurun_out/r2o_stages.log | grep -A13 "pass 1" | grep -E "mrte|adm|plm|decode|hifigan|full"; done
timeout 600 python tools/bench_tc_shapes.py --fmt f16x2 --reps 10 --variants 0 --shapes 0,1,2,3,4,5,6,7 2>&1 | tee -a gpurun_out/r2o_tc_shapes.log
