#!/bin/bash
# GPU call A of round 2: the whole GPU suite on the new build (both operand formats), stage times and a bench of each engine.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_rates.jsonl
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:randomly 2>&1 | tee gpurun_out/r2a_pytest.log | tail -60
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/r2a_smoke.log | tail -5
for eng in f16x2 bf16x3; do
  echo "== stages $eng"
  MEGATTS2_ENGINE=$eng timeout 600 python tools/time_stages.py --reps 2 2>&1 | tee gpurun_out/r2a_stages_$eng.log | tail -45
done
echo "== bench f16x2"
MEGATTS2_ENGINE=f16x2 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2a_bench_f16x2.json 2> gpurun_out/r2a_bench_f16x2.err; tail -5 gpurun_out/r2a_bench_f16x2.err; cat gpurun_out/r2a_bench_f16x2.json
echo "== bench bf16x3 (no cpu leg)"
MEGATTS2_ENGINE=bf16x3 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2a_bench_bf16x3.json 2> gpurun_out/r2a_bench_bf16x3.err; tail -3 gpurun_out/r2a_bench_bf16x3.err; cat gpurun_out/r2a_bench_bf16x3.json
