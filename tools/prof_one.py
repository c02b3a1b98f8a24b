#!/usr/bin/env python3
"""Run ONE shape a few times (the command line profiled under ncu; also prints the event-timed ms).
    python tools/prof_one.py --shape 100,64,64,64 [--mode c2c|real|half] [--steps 5] [--inverse]
Environment knobs (B200FFT_FUSED, B200FFT_CHUNK_MB, B200FFT_FUSED_PREFER, B200FFT_PREFER) select the kernels."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch

import b200fft


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", required=True)
    ap.add_argument("--mode", default="c2c")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--inverse", action="store_true")
    a = ap.parse_args()
    shape = tuple(int(v) for v in a.shape.split(","))
    comps = 2 if a.mode == "c2c" else 1
    x = torch.randn(shape + (comps,), device="cuda")
    oshape = shape[:-1] + (shape[-1] // 2 + 1, 2) if a.mode == "half" else shape + (2,)
    out = torch.empty(oshape, device="cuda")
    plan = b200fft.plan_fft("float32", "float32", x.shape, oshape, inverse=a.inverse,
                            real_mode=b200fft.REAL_HALF if a.mode == "half" else b200fft.REAL_FULL)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.exec(out, x, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        plan.exec(out, x, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    ab = x.numel() * 4 + out.numel() * 4
    print(json.dumps({"shape": shape, "ms": ms, "algorithmic_gbs": ab / ms / 1e6, "plan": plan.describe().strip().split("\n")}))


if __name__ == "__main__":
    main()
