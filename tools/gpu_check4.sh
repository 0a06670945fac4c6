#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== mel C2"
timeout 300 python tools/bench_mel.py 10000 2>&1 | tee gpurun_out/bench_mel_c2.log | tail -2
echo "== ncu launch list (TC engine, B=64, one step)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/launches_tc_b64.csv python tools/profile_step.py --batch 64 > gpurun_out/ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_tc_b64.csv gpurun_out/launch_summary_tc_b64.md | tail -25
echo "== ncu full: hifigan tapconv kernels"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tapconv_kernel -s 20 -c 6 \
   -o gpurun_out/prof_hifigan_tapconv -f python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/ncu_full1.log 2>&1
echo "== ncu full: gemm_bf16x3"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_bf16x3 -s 2000 -c 2 \
   -o gpurun_out/prof_gemm_tc -f python tools/profile_step.py --batch 64 --stage plm > gpurun_out/ncu_full2.log 2>&1
echo "== ncu full: mel kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 2 -c 1 \
   -o gpurun_out/prof_mel -f python tools/bench_mel.py 2000 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== full GPU test suite with the TC engine as default"
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 | tee gpurun_out/pytest_tc_default.log
