#!/usr/bin/env python
"""One bench step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--batch 64]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--stage", default="all", choices=["all", "mel", "hifigan", "plm"])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    tts = bench.build_product(dev)
    wav, phone, forced = bench.make_inputs(a.batch, 1234)
    wav, phone, forced = wav.to(dev), phone.to(dev), forced.to(dev)
    bench.gpu_step(tts, wav, phone, forced)          # warm-up: plans, workspace, attributes
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    bench.gpu_step(tts, wav, phone, forced)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
