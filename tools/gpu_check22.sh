#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== causal decode tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --timeout 300 -k "causal" -s 2>&1 | tail -15 | tee gpurun_out/causal_tests.log
echo "== decode bench"
timeout 600 python tools/bench_decode.py 2>&1 | tee gpurun_out/bench_decode.log | tail -8
