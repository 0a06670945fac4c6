#!/bin/bash
# GPU call I: ncu --set full captures for the traffic table (top kernel classes), C2 / torch_gpu / reference arms
set -u
mkdir -p gpurun_out
echo "== ncu full: PLM stage, late steps (pair GEMMs, attention, LayerNorm)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"conv_tc_kernel|attn_kernel|layernorm_reg_kernel" --launch-skip 5200 -c 16 -f -o gpurun_out/r2i_plm_kernels \
  python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2i_ncu_plm.log 2>&1; tail -1 gpurun_out/r2i_ncu_plm.log
echo "== ncu full: HiFi-GAN stage (one conv per stage)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"conv_tc_kernel" --launch-skip 3 -c 72 -f -o gpurun_out/r2i_hifigan_kernels \
  python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/r2i_ncu_hifigan.log 2>&1; tail -1 gpurun_out/r2i_ncu_hifigan.log
echo "== bench c2"
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err; tail -2 gpurun_out/r2i_bench_c2.err; cut -c1-600 gpurun_out/r2i_bench_c2.json
echo "== bench torch_gpu"
timeout 900 python bench.py --impl torch_gpu --steps 2 --warmup 1 > gpurun_out/r2i_bench_torch_gpu.json 2> gpurun_out/r2i_bench_torch_gpu.err; tail -3 gpurun_out/r2i_bench_torch_gpu.err; cat gpurun_out/r2i_bench_torch_gpu.json
echo "== bench reference"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2i_bench_reference.json 2> gpurun_out/r2i_bench_reference.err; tail -2 gpurun_out/r2i_bench_reference.err; cut -c1-500 gpurun_out/r2i_bench_reference.json
