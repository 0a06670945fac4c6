#!/bin/bash
# GPU call H: new rows (audio front/back ends 8f-3, stage-2 latent dump 8f-2) + full suite
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:randomly 2>&1 | tee gpurun_out/r2h_pytest.log | tail -15
