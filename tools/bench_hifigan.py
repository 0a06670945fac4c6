#!/usr/bin/env python
"""Vocoder stage alone (diagnostics, GPU only): HiFi-GAN on the batch-64 bench shape (512 mel frames), CUDA events over
`--reps` runs with an L2 flush in between; prints min / median.  Used for A/B of two builds on one box (the full step's
box-to-box and run-to-run variance, +-4 %, hides vocoder changes of that size)."""
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
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--reps", type=int, default=12)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    tts = bench.build_product(dev)
    mel = (torch.randn(a.batch, a.frames, 80, generator=torch.Generator().manual_seed(5)) * 2 - 4).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        tts.hifi_gan.decode_batch_cl(mel)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tts.hifi_gan.decode_batch_cl(mel)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"hifigan B{a.batch} T{a.frames}: min {ts[0]:.2f} ms  median {ts[len(ts) // 2]:.2f} ms  max {ts[-1]:.2f} ms", flush=True)


if __name__ == "__main__":
    main()
