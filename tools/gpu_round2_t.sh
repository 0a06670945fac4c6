#!/bin/bash
# GPU call T: four accumulator buffers for the narrow tiles (BN <= 64) vs two: A/B of two builds on one box
set -u
mkdir -p gpurun_out
L=megatts2_b200/lib
cp $L/libmegatts2_b200.so $L/new.keep
for rep in 1 2; do for v in base new; do
  if [ $v = base ]; then cp $L/libmegatts2_b200.base.so.keep $L/libmegatts2_b200.so; else cp $L/new.keep $L/libmegatts2_b200.so; fi
  echo "== build $v"
  timeout 600 python tools/bench_tc_shapes.py --fmt f16x2 --reps 10 --variants 0 --shapes 6,7 2>&1 | grep hifi
  timeout 600 python tools/time_stages.py --reps 2 2>&1 | grep -A13 "pass 1" | grep -E "mrte|adm|plm|hifigan|full"
done; done 2>&1 | tee gpurun_out/r2t_nacc4_ab.log
cp $L/new.keep $L/libmegatts2_b200.so
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly -x -k "tc or hifigan or e2e or conv" 2>&1 | tail -3
