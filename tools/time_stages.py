#!/usr/bin/env python
"""Per-stage wall/GPU timing of one C4 step with flushed progress prints (diagnostics)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def say(*a):
    print(*a, flush=True)


def timed(name, fn, reps=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    say(f"  {name:28s} {dt * 1e3:10.2f} ms")
    return out, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    from megatts2_b200 import ops
    from megatts2_b200.modules.tokenizer import extract_mel_spec
    dev = torch.device("cuda", 0)
    t0 = time.perf_counter()
    tts = bench.build_product(dev)
    say(f"build_product: {time.perf_counter() - t0:.1f} s")
    wav, phone, forced = bench.make_inputs(range(a.batch))
    wav, phone, forced = wav.to(dev), phone.to(dev), forced.to(dev)
    for rep in range(a.reps):
        say(f"== pass {rep} (batch {a.batch})")
        n0 = ops.launch_count()
        mel, _ = timed("mel front end", lambda: extract_mel_spec(wav, frames_major=True))
        tc, _ = timed("mrte.tc_latent", lambda: tts.generator.mrte.tc_latent(phone, mel))
        dt, _ = timed("adm.infer", lambda: tts.adm.infer(tc))
        exp, _ = timed("length regulator", lambda: tts.lr(tc, forced))
        tc8, _ = timed("maxpool", lambda: ops.maxpool_time(exp, 8))
        codes, _ = timed("plm.infer", lambda: tts.plm.infer(tc8))
        melo, _ = timed("decode_mel", lambda: tts.generator.decode_mel_cl(exp, codes))
        wv, _ = timed("hifigan", lambda: tts.hifi_gan.decode_batch_cl(melo))
        _, _ = timed("hifigan (prompt re-vocode)", lambda: tts.hifi_gan.decode_batch_cl(mel))
        say(f"  launches this pass: {ops.launch_count() - n0}; wav {tuple(wv.shape)} finite={bool(torch.isfinite(wv).all())}")
    _, dt = timed("full gpu_step", lambda: bench.gpu_step(tts, wav, phone, forced))
    say(f"samples/s = {a.batch * bench.SAMPLES_PER_UTT / dt:.0f}")
    import ctypes as C
    from megatts2_b200 import _lib as L
    lib = L.lib()
    from megatts2_b200 import graphs
    lib.mtts_trace_begin(ops._stream())
    with graphs.disabled():                 # the trace records events between launches
        bench.gpu_step(tts, wav, phone, forced)
    buf = C.create_string_buffer(16384)
    lib.mtts_trace_end(buf, 16384)
    say("== per-launcher event trace of one step (event overhead included)")
    say(buf.value.decode())


if __name__ == "__main__":
    main()
