#!/usr/bin/env python3
"""Time the fused N-d kernel against the per-axis passes and cuFFT on the BASELINE N-d shapes.
    python tools/fused_bench.py [--chunks 8,16,32] [--prefer plane,rows] > gpurun_out/fused_bench.jsonl
Every line is one (shape, configuration) timing: CUDA events, 20 calls after 5 warm-ups."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import numpy as np
import torch

import b200fft
from bench import CuFFT, time_gpu, measured_peak

SHAPES = [("3d_100x64^3", (100, 64, 64, 64), "c2c"), ("3d_10x128^3", (10, 128, 128, 128), "c2c"),
          ("3d_1x256^3", (1, 256, 256, 256), "c2c"), ("3d_1x512^3", (1, 512, 512, 512), "c2c"),
          ("2d_100x640x480", (100, 640, 480), "c2c"), ("2d_100x640x480_real", (100, 640, 480), "real"),
          ("2d_100x640x480_half", (100, 640, 480), "half")]


def run(name, shape, mode, env, peak, steps=20):
    for k in ("B200FFT_FUSED", "B200FFT_CHUNK_MB", "B200FFT_FUSED_PREFER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    comps = 2 if mode == "c2c" else 1
    x = torch.randn(tuple(shape) + (comps,), device="cuda")
    oshape = tuple(shape[:-1]) + (shape[-1] // 2 + 1, 2) if mode == "half" else tuple(shape) + (2,)
    out = torch.empty(oshape, device="cuda")
    rm = b200fft.REAL_HALF if mode == "half" else b200fft.REAL_FULL
    plan = b200fft.plan_fft("float32", "float32", x.shape, oshape, real_mode=rm)
    st = torch.cuda.current_stream().cuda_stream
    ms = time_gpu(lambda: plan.exec(out, x, st), 5, steps, torch)
    ab = x.numel() * 4 + out.numel() * 4
    row = {"shape": name, "env": env, "ms": round(ms, 5), "hbm_frac": round(ab / ms / 1e6 / peak, 4),
           "plan": plan.describe().strip().split("\n")[0][:160]}
    plan.destroy()
    del x, out
    torch.cuda.empty_cache()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", default="8,16,32")
    ap.add_argument("--prefer", default="+32_plane,+32_rows,t256_plane")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    peak, _ = measured_peak()
    for name, shape, mode in SHAPES:
        if a.only and a.only not in name:
            continue
        print(json.dumps(run(name, shape, mode, {"B200FFT_FUSED": "0"}, peak)), flush=True)
        prefs = a.prefer.split(",") if "64^3" in name else ["ndA", "nd" + name.split("x")[1].split("^")[0].split("_")[0]]
        for pref in prefs:
            for c in a.chunks.split(","):
                env = {"B200FFT_FUSED": "1", "B200FFT_CHUNK_MB": c}
                if pref:
                    env["B200FFT_FUSED_PREFER"] = pref
                try:
                    print(json.dumps(run(name, shape, mode, env, peak)), flush=True)
                except Exception as e:
                    print(json.dumps({"shape": name, "env": env, "error": str(e)}), flush=True)
        try:
            x = torch.randn(tuple(shape) + ((2,) if mode == "c2c" else ()), device="cuda")
            oshape = tuple(shape) + (2,) if mode == "c2c" else tuple(shape[:-1]) + (shape[-1] // 2 + 1, 2)
            out = torch.empty(oshape, device="cuda")
            cf = CuFFT(shape, r2c=mode != "c2c")
            st = torch.cuda.current_stream().cuda_stream
            cms = time_gpu(lambda: cf.exec(x, out, st), 5, 20, torch)
            cf.destroy()
            print(json.dumps({"shape": name, "cufft_ms": round(cms, 5)}), flush=True)
            del x, out
        except Exception as e:
            print(json.dumps({"shape": name, "cufft_error": str(e)}), flush=True)


if __name__ == "__main__":
    main()
