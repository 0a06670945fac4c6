"""PLM / ADM decode throughput: the reference-faithful infer() (non-causal full recompute, O(T^2)) next to the opt-in
causal KV-cache decode.  GPU only; reported separately from bench.py's headline metric.

    python tools/bench_decode.py [--full-c3]      # --full-c3 also runs infer() at BASELINE config C3 (B=16, T=512)
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers  # noqa: E402
from oracle import weights  # noqa: E402  (seeded weight specs only; nothing of the oracle is timed here)

DEV = "cuda:0"


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-c3", action="store_true")
    args = ap.parse_args()
    t0 = time.time()
    plm = helpers.build_plm(weights.plm_state_dict(), DEV)
    adm = helpers.build_adm(weights.adm_state_dict(), DEV)
    print(f"models built in {time.time() - t0:.1f} s", file=sys.stderr)
    rows = []
    g = torch.Generator().manual_seed(7)
    for name, B, T, run_full in (("C4 PLM stage", 64, 64, True), ("PLM B16 T128", 16, 128, True),
                                 ("C3 PLM (B16 T512)", 16, 512, args.full_c3)):
        tc = F.relu(torch.randn(B, T, 512, generator=g)).to(DEV)
        ms_c, ids_c = timed(lambda: plm.infer_causal(tc), 2)
        row = {"config": name, "B": B, "T": T, "causal_ms": round(ms_c, 2), "causal_tok_s": round(B * T / ms_c * 1e3, 1)}
        if run_full:
            ms_f, ids_f = timed(lambda: plm.infer(tc), 1)
            row.update(infer_ms=round(ms_f, 2), infer_tok_s=round(B * T / ms_f * 1e3, 1), speedup=round(ms_f / ms_c, 1),
                       ids_equal_rate=round((ids_c == ids_f).float().mean().item(), 3))
        rows.append(row)
        print(json.dumps(row), flush=True)
    tcl = F.relu(torch.randn(64, 64, 512, generator=g)).to(DEV)
    ms_c, _ = timed(lambda: adm.infer_causal(tcl), 2)
    ms_f, _ = timed(lambda: adm.infer(tcl), 1)
    row = {"config": "C4 ADM stage", "B": 64, "T": 64, "causal_ms": round(ms_c, 2), "infer_ms": round(ms_f, 2),
           "speedup": round(ms_f / ms_c, 1)}
    print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
