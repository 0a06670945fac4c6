#!/bin/bash
# GPU call L: epilogue prefetch + deeper halo buffering: parity, stage times, shape variants; PLM GEMM ncu capture
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly 2>&1 | tee gpurun_out/r2l_pytest.log | tail -5
timeout 600 python tools/time_stages.py --reps 2 2>&1 | tee gpurun_out/r2l_stages.log | grep -A13 "pass 1"
for v in "X=1" "MEGATTS2_TC_SWB64=1" "MEGATTS2_TC_PAIR=0"; do
  echo "== shapes $v"
  env $v timeout 600 python tools/bench_tc_shapes.py --fmt f16x2 --reps 10 --variants 0 --shapes 0,1,2,3,4,5,6,7 2>&1 | tee -a gpurun_out/r2l_tc_shapes.log
done
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc_kernel --launch-skip 2700 -c 8 -f \
  -o gpurun_out/r2k_plm_gemm python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2k_ncu_1.log 2>&1; tail -1 gpurun_out/r2k_ncu_1.log
