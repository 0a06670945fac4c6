#!/bin/bash
# ncu --set full captures of the top kernel classes on the final build (traffic table, profiles/r2_traffic.*)
set -u
mkdir -p gpurun_out
export MEGATTS2_GRAPHS=0      # eager enqueue: kernel names / launch-skip counts as in the earlier captures
timeout 900 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled \
  -k regex:"conv_tc_kernel<\(int\)128, \(int\)128, \(int\)1" --launch-skip 2700 -c 6 -f -o gpurun_out/r2f_plm_gemm \
  python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2f_ncu_1.log 2>&1; tail -1 gpurun_out/r2f_ncu_1.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"attn_kernel|layernorm_reg_kernel" \
  --launch-skip 2900 -c 6 -f -o gpurun_out/r2f_plm_attn_ln python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2f_ncu_2.log 2>&1; tail -1 gpurun_out/r2f_ncu_2.log
i=0
for pat in "conv_tc_kernel<\(int\)32, \(int\)64" "conv_tc_kernel<\(int\)64, \(int\)128, \(int\)1, \(int\)2, \(int\)1" "conv_tc_kernel<\(int\)128, \(int\)128, \(int\)0"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled \
    -k regex:"$pat" --launch-skip 6 -c 4 -f -o gpurun_out/r2f_hifigan_k$i \
    python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/r2f_ncu_h$i.log 2>&1; tail -1 gpurun_out/r2f_ncu_h$i.log
done
ls -la gpurun_out/r2f_*.ncu-rep
