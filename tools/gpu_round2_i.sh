#!/bin/bash
# GPU call I: ncu --set full captures for the traffic table (top kernel classes), torch_gpu arm, training tests
set -u
mkdir -p gpurun_out
echo "== training tests"
timeout 900 python -m pytest tests/test_gpu_train.py -q --timeout 600 -p no:randomly 2>&1 | tee gpurun_out/r2i_pytest_train.log | tail -6
echo "== ncu full: PLM stage, late steps (pair GEMMs, attention, LayerNorm)"
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:"conv_tc_kernel|attn_kernel|layernorm_reg_kernel" --launch-skip 5200 -c 14 -f -o gpurun_out/r2i_plm_kernels \
  python tools/profile_step.py --batch 64 --stage plm > gpurun_out/r2i_ncu_plm.log 2>&1; tail -1 gpurun_out/r2i_ncu_plm.log
echo "== ncu full: HiFi-GAN stage, a few convs per template"
i=0
for pat in "conv_tc_kernel<\(int\)32, \(int\)64" "conv_tc_kernel<\(int\)64, \(int\)128, \(int\)0, \(int\)2, \(int\)1" "conv_tc_kernel<\(int\)128, \(int\)128, \(int\)0"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled \
    -k regex:"$pat" --launch-skip 6 -c 4 -f -o gpurun_out/r2i_hifigan_k$i \
    python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/r2i_ncu_hifigan_$i.log 2>&1; tail -1 gpurun_out/r2i_ncu_hifigan_$i.log
done
ls -la gpurun_out/*.ncu-rep
echo "== bench torch_gpu"
timeout 900 python bench.py --impl torch_gpu --steps 2 --warmup 1 > gpurun_out/r2i_bench_torch_gpu.json 2> gpurun_out/r2i_bench_torch_gpu.err; tail -3 gpurun_out/r2i_bench_torch_gpu.err; cat gpurun_out/r2i_bench_torch_gpu.json
echo "== bench c2"
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err; cut -c1-300 gpurun_out/r2i_bench_c2.json
bash tools/gpu_round2_j.sh
