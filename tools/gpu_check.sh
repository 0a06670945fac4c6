#!/bin/bash
# Runs on the GPU box (via gpurun): GPU parity tests, smoke, a short bench.  Everything is
# logged under gpurun_out/.  Failing tests are re-run one per process (a sticky CUDA error in
# one test must not mask the others).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu (single process)"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:randomly 2>&1 | tee gpurun_out/pytest_all.log | tail -40
python -m pytest tests -m gpu --lf --co -q 2>/dev/null | grep '::' > gpurun_out/failed_ids.txt
if grep -q "failed" gpurun_out/pytest_all.log; then
  echo "== re-running failures individually"
  : > gpurun_out/pytest_failed_individually.log
  while read -r id; do
    echo "#### $id" >> gpurun_out/pytest_failed_individually.log
    timeout 600 python -m pytest "$id" -q -x --timeout 500 2>&1 | tail -60 >> gpurun_out/pytest_failed_individually.log
  done < gpurun_out/failed_ids.txt
  tail -150 gpurun_out/pytest_failed_individually.log
fi
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | tail -5
if [ "${1:-}" != "nobench" ]; then
  echo "== bench (short)"
  timeout 1200 python bench.py --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_short.log | tail -3
fi
