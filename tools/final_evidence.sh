#!/bin/bash
# Round-end evidence on the GPU box: full GPU suite, smoke, the default bench line, the reference arm,
# and the ncu launch list of one synthesis step (per-kernel share).  Logs go to gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 | tee gpurun_out/final_pytest_gpu.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
echo "== bench (default flags)"
timeout 1500 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 1500 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 600 gpurun_out/final_bench_reference.json
echo "== stage times + trace"
timeout 300 python tools/time_stages.py --batch 64 --reps 2 > gpurun_out/final_stages.log 2>&1; grep -E "full gpu|samples/s" gpurun_out/final_stages.log
