#!/bin/bash
# GPU call N: mel front end (16 frames per CTA, CTA-wide filterbank): tests, C2 bench, ncu full of mel_kernel; VQ training tests
set -u
mkdir -p gpurun_out
echo "== mel + vq-train tests"
timeout 900 python -m pytest tests -q -m gpu -k "mel or audio or e2e or bulk or vq_train or kmeans or expiry or uninitialised" --timeout 300 -p no:randomly 2>&1 | tail -15
echo "== bench c2"
timeout 600 python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/r2n_bench_c2.json 2> gpurun_out/r2n_bench_c2.err; tail -2 gpurun_out/r2n_bench_c2.err; cut -c1-400 gpurun_out/r2n_bench_c2.json
echo "== ncu mel"
timeout 600 ncu --set full --clock-control none -k regex:"mel_kernel" --launch-skip 2 -c 1 -f -o gpurun_out/r2n_mel \
  python bench.py --config c2 --steps 1 --warmup 3 > gpurun_out/r2n_ncu_mel.log 2>&1; tail -1 gpurun_out/r2n_ncu_mel.log
ls -la gpurun_out/r2n_mel.ncu-rep
