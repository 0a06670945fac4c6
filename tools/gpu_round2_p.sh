#!/bin/bash
# GPU call P: ncu (SASS-level) of the C = 32 / C = 64 halo-form vocoder convs after the epilogue clean-up
set -u
mkdir -p gpurun_out
i=0
for pat in "conv_tc_kernel<\(int\)32, \(int\)64" "conv_tc_kernel<\(int\)64, \(int\)128, \(int\)0, \(int\)2, \(int\)1"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --profile-from-start off --kernel-name-base demangled \
    -k regex:"$pat" --launch-skip 6 -c 2 -f -o gpurun_out/r2p_hifigan_k$i \
    python tools/profile_step.py --batch 64 --stage hifigan > gpurun_out/r2p_ncu_hifigan_$i.log 2>&1; tail -1 gpurun_out/r2p_ncu_hifigan_$i.log
done
ls -la gpurun_out/r2p_*.ncu-rep
