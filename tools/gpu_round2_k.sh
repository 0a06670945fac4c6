#!/bin/bash
# GPU call K: PLM-stage ncu captures (pair GEMM / attention / LayerNorm), reference arm, full suite
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc_kernel --launch-skip 2700 -c 8 -f \
  -o gpurun_out/r2k_plm_gemm python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2k_ncu_1.log 2>&1; tail -1 gpurun_out/r2k_ncu_1.log
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:attn_kernel --launch-skip 650 -c 3 -f \
  -o gpurun_out/r2k_plm_attn python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2k_ncu_2.log 2>&1; tail -1 gpurun_out/r2k_ncu_2.log
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:layernorm_reg_kernel --launch-skip 1300 -c 3 -f \
  -o gpurun_out/r2k_plm_ln python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2k_ncu_3.log 2>&1; tail -1 gpurun_out/r2k_ncu_3.log
ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly 2>&1 | tee gpurun_out/r2k_pytest.log | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
