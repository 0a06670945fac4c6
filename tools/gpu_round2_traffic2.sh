#!/bin/bash
set -u
mkdir -p gpurun_out
export MEGATTS2_GRAPHS=0
timeout 900 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled \
  -k regex:"conv_tc_kernel<\(int\)128, \(int\)128, \(int\)1" --launch-skip 2000 -c 6 -f -o gpurun_out/r2f_plm_gemm \
  python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2f_ncu_1.log 2>&1; tail -1 gpurun_out/r2f_ncu_1.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"attn_kernel|layernorm_reg_kernel" \
  --launch-skip 1800 -c 6 -f -o gpurun_out/r2f_plm_attn_ln python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2f_ncu_2.log 2>&1; tail -1 gpurun_out/r2f_ncu_2.log
ls -la gpurun_out/r2f_plm*.ncu-rep
