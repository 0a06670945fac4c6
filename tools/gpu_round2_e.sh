#!/bin/bash
# GPU call E: PDL correctness + timing; ncu --set full of the C = 32 conv (plain and halo form)
set -u
mkdir -p gpurun_out
echo "== tests (PDL on)"
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly 2>&1 | tee gpurun_out/r2e_pytest.log | tail -8
for m in 1 0; do
  echo "== stages PDL=$m"
  MEGATTS2_PDL=$m timeout 600 python tools/time_stages.py --reps 2 2>&1 | tee gpurun_out/r2e_stages_pdl$m.log | grep -A12 "pass 1"
done
for h in 0 1; do
  echo "== ncu full C32 k7 HALO=$h"
  MEGATTS2_TC_HALO=$h timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f \
     -o gpurun_out/r2e_c32k7_halo$h python tools/bench_tc_shapes.py --fmt f16x2 --shapes 7 --variants 0 --reps 2 > gpurun_out/r2e_ncu_halo$h.log 2>&1
  tail -2 gpurun_out/r2e_ncu_halo$h.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
