#!/bin/bash
# GPU call M: A/B of the epilogue load prefetch on one box (stage times x2 each, alternating)
set -u
mkdir -p gpurun_out
for rep in 1 2; do for v in 1 0; do
  echo "== EPI_PREFETCH=$v (rep $rep)"
  MEGATTS2_TC_EPI_PREFETCH=$v timeout 600 python tools/time_stages.py --reps 2 2>&1 | grep -A11 "pass 1" | grep -E "adm|plm|hifigan|full"
done; done 2>&1 | tee gpurun_out/r2m_epi_prefetch_ab.log
