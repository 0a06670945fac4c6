#!/bin/bash
# GPU call F: uniform-issue build: full GPU suite, per-shape rates, stage times (PDL on/off, halo on/off)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly 2>&1 | tee gpurun_out/r2f_pytest.log | tail -8
echo "== shapes f16x2"
timeout 600 python tools/bench_tc_shapes.py --fmt f16x2 --reps 10 --variants 0 2>&1 | tee gpurun_out/r2f_tc_shapes_f16x2.log
for cfg in "1 1" "0 1" "1 0" "1 1"; do
  set -- $cfg
  echo "== stages PDL=$1 HALO=$2"
  MEGATTS2_PDL=$1 MEGATTS2_TC_HALO=$2 timeout 600 python tools/time_stages.py --reps 2 2>&1 | tee gpurun_out/r2f_stages_pdl$1_halo$2.log | grep -A13 "pass 1"
done
