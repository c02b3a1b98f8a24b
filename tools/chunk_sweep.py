#!/usr/bin/env python3
"""Sweep B200FFT_PASS_CHUNK_MB (L2-resident pass groups, csrc/api.cu) over the N-d BASELINE shapes: ms per transform."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch

import b200fft

SHAPES = [("2d_100x640x480", (100, 640, 480), False, 0), ("2d_100x640x480_r2c", (100, 640, 480), "half", 0),
          ("3d_1x256^3", (1, 256, 256, 256), False, 0), ("3d_1x512^3", (1, 512, 512, 512), False, 0),
          ("3d_10x128^3_nofused", (10, 128, 128, 128), False, b200fft.FLAG_NO_FUSED),
          ("3d_100x64^3_nofused", (100, 64, 64, 64), False, b200fft.FLAG_NO_FUSED),
          ("3d_1x256^3_r2c", (1, 256, 256, 256), "half", 0), ("3d_10x128^3_r2c", (10, 128, 128, 128), "half", 0),
          ("3d_100x64^3_r2c", (100, 64, 64, 64), "half", 0), ("2d_10x1920x1080", (10, 1920, 1080), False, 0)]


def time_ms(fn, warm=3, steps=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    budgets = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,8,16,24,32,48,64").split(",")]
    st = torch.cuda.current_stream().cuda_stream
    for name, shape, real, flags in SHAPES:
        half = real == "half"
        x = torch.randn(tuple(shape) + (1 if real else 2,), device="cuda")
        oshape = tuple(shape[:-1]) + (shape[-1] // 2 + 1, 2) if half else tuple(shape) + (2,)
        out = torch.empty(oshape, device="cuda")
        row = {"shape": name}
        for mb in budgets:
            os.environ["B200FFT_PASS_CHUNK_MB"] = str(mb)
            plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape, flags=flags,
                                    real_mode=b200fft.REAL_HALF if half else b200fft.REAL_FULL)
            row["%dMB" % mb] = round(time_ms(lambda: plan.exec(out, x, st)), 4)
            row["launches_%dMB" % mb] = plan.launches
            plan.destroy()
        print(json.dumps(row), flush=True)
        del x, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
