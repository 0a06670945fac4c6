#!/usr/bin/env python
"""A/B sweep of the overlapped prompt re-vocode (Megatts._synthesize): side-stream SM budget V, share of the batch re-vocoded
there, and the first main-stream stage that runs beside it.  One process, one model; every configuration gets its own set-up
passes (plans, graph capture) before the timed ones.  GPU only (diagnostics; bench.py is the contract's measurement).

    python tools/sweep_overlap.py [--configs 0 100:1:mrte 100:0.5:adm ...] [--steps 3]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--setup", type=int, default=5)
    ap.add_argument("--configs", nargs="*", default=["0", "100:1:mrte", "90:1:mrte", "110:1:mrte", "100:0.5:adm", "116:0.5:adm",
                                                      "124:0.5:adm", "108:0.75:mrte", "0"])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    t0 = time.perf_counter()
    tts = bench.build_product(dev)
    print(f"build_product: {time.perf_counter() - t0:.1f} s", flush=True)
    wav, phone, forced = bench.make_inputs(range(a.batch))
    wav, phone, forced = wav.to(dev), phone.to(dev), forced.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    for cfg in a.configs:
        parts = cfg.split(":")
        os.environ["MEGATTS2_REVOCODE_SMS"] = parts[0]
        os.environ["MEGATTS2_REVOCODE_FRAC"] = parts[1] if len(parts) > 1 else "1.0"
        os.environ["MEGATTS2_REVOCODE_FROM"] = parts[2] if len(parts) > 2 else "mrte"
        for _ in range(a.setup):
            out = bench.gpu_step(tts, wav, phone, forced, intermediates=True)
        torch.cuda.synchronize()
        ms = []
        for _ in range(a.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = bench.gpu_step(tts, wav, phone, forced, intermediates=True)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        row = {"config": cfg, "ms_min": round(min(ms), 2), "ms_median": round(sorted(ms)[len(ms) // 2], 2), "ms_max": round(max(ms), 2),
               "samples_per_s": round(a.batch * bench.SAMPLES_PER_UTT / (sorted(ms)[len(ms) // 2] * 1e-3))}
        if ref is None:
            ref = {k: out[k].clone() for k in ("wav", "p_codes", "dt", "tc_latent")}
        else:
            row.update(wav_equal=bool(torch.equal(out["wav"], ref["wav"])),
                       wav_maxdiff=float((out["wav"] - ref["wav"]).abs().max()),
                       ids_equal=float((out["p_codes"] == ref["p_codes"]).float().mean()),
                       dt_equal=float((out["dt"] == ref["dt"]).float().mean()),
                       tc_latent_maxdiff=float((out["tc_latent"] - ref["tc_latent"]).abs().max()))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
