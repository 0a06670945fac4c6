#!/bin/bash
# ncu evidence for profiles/: launch list of ONE synthesis step (per-kernel share) and --set full captures of the
# attention / LayerNorm / tap-GEMM kernels inside the PLM stage.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/final_launches.csv python tools/profile_step.py --batch 64 > gpurun_out/final_launches.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/final_launches.csv)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"attn_kernel|layernorm_reg_kernel|conv_bf16x3_kernel" --launch-skip 2400 -c 14 -f -o gpurun_out/final_plm_kernels \
  python tools/profile_step.py --batch 64 --stage plm > gpurun_out/final_plm_kernels.log 2>&1
echo "set-full rc=$?"; ls -la gpurun_out/*.ncu-rep
