#!/usr/bin/env python3
"""Tuning aid: time one shape (optionally one axis via --mask) for each B200FFT_PREFER key.

    python tools/sweep.py --shape 100000,1024 --prefer rows1024_16x16x4,rows1024_32x32
    python tools/sweep.py --shape 100,640,480 --mask 2 --prefer cols640_      # axis 1 only (bit per axis)
One subprocess per key (the registry reads B200FFT_PREFER at plan time, but keeping runs
isolated also keeps L2 / clocks comparable). Prints ms, algorithmic GB/s and the kernels used.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]


def child(shape, mask, steps, real):
    import torch
    import b200fft
    comps = 1 if real else 2
    x = torch.randn(tuple(shape) + (comps,), device="cuda", dtype=torch.float32)
    out = torch.empty(tuple(shape) + (2,), device="cuda", dtype=torch.float32)
    src = x
    if mask and not real:
        src = out  # single-axis pass in place on a complex buffer (what an inner pass of an N-d plan does)
        out.copy_(x)
    plan = b200fft.plan_fft("float32", "float32", src.shape, out.shape, axis_mask=mask)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.exec(out, src, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(steps):
            plan.exec(out, src, st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    import numpy as np
    pts = int(np.prod(shape))
    print(json.dumps({"ms": best, "gbs": pts * (8 + 4 * comps) / best / 1e6,
                      "kernels": [k.split(" n=")[0].split(": ")[1] for k in plan.describe().strip().split("\n") if ": " in k]}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", required=True)
    ap.add_argument("--mask", type=int, default=0)
    ap.add_argument("--prefer", default="")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--real", action="store_true")
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    shape = [int(v) for v in a.shape.split(",")]
    if a.child:
        child(shape, a.mask, a.steps, a.real)
        sys.exit(0)
    for key in (a.prefer.split(",") if a.prefer else [""]):
        env = dict(os.environ, B200FFT_PREFER=key)
        cmd = [sys.executable, __file__, "--child", "--shape", a.shape, "--mask", str(a.mask), "--steps", str(a.steps)]
        if a.real:
            cmd.append("--real")
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        line = r.stdout.strip().split("\n")[-1] if r.stdout.strip() else r.stderr[-300:]
        print("%-28s %s" % (key or "(default)", line), flush=True)
