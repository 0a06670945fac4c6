#!/bin/bash
# GPU call AT: fp32 attention kernel capped at 80 registers (3 CTAs per SM instead of 2): A/B of two builds
set -u
mkdir -p gpurun_out
L=megatts2_b200/lib
cp $L/libmegatts2_b200.so $L/new.keep
for rep in 1 2; do for v in base new; do
  cp $L/$v.keep $L/libmegatts2_b200.so
  echo "== build $v"
  BENCH_S=16,32,48,64 timeout 300 python tools/bench_attention.py 2>&1 | grep -E "B64"
  timeout 600 python tools/time_stages.py --reps 5 2>&1 | grep -A13 "pass 4" | grep -E "mrte|adm|plm|full"
done; done 2>&1 | tee gpurun_out/r2at_attention_regs_ab.log
cp $L/new.keep $L/libmegatts2_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "attention or plm or adm or encoder or mrte" --timeout 300 -p no:randomly 2>&1 | tail -3
