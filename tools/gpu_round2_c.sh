#!/bin/bash
# GPU call C: tensor-core attention: parity tests, then stage times with it on / off
set -u
mkdir -p gpurun_out
echo "== attention + encoder tests"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_baseline_sizes.py -q --timeout 600 -p no:randomly -k "attention or encoder or plm or adm or e2e or golden" 2>&1 | tee gpurun_out/r2c_pytest.log | tail -30
for m in 1 0; do
  echo "== stages ATTN_TC=$m"
  MEGATTS2_ATTN_TC=$m timeout 600 python tools/time_stages.py --reps 2 2>&1 | tee gpurun_out/r2c_stages_attn$m.log | grep -A12 "pass 1"
  grep -E "attn_launch|attention" gpurun_out/r2c_stages_attn$m.log
done
