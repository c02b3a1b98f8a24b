#!/usr/bin/env python3
"""Every shape the reference publishes (fft/bench.mojo:107-122, README.md:21-74), ours vs cuFFT on one B200:
    python tools/published_shapes.py > gpurun_out/published_shapes.jsonl
CUDA events, 20 calls after 5 warm-ups, forward C2C fp32; parity of batch item 0 against torch float64."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import numpy as np
import torch

import b200fft
from bench import CuFFT, time_gpu, measured_peak

SHAPES = [(500000, 128), (100000, 1024), (500000, 93), (1000000, 93), (100, 16384), (100, 640, 480), (10, 1920, 1080),
          (1, 3840, 2160), (1, 7680, 4320), (100, 64, 64, 64), (10, 128, 128, 128), (1, 256, 256, 256), (1, 512, 512, 512),
          (1, 64, 64, 64, 64), (1, 25, 160, 160, 48)]


def main():
    peak, _ = measured_peak()
    st = torch.cuda.current_stream().cuda_stream
    for shape in SHAPES:
        row = {"shape": list(shape)}
        try:
            x = torch.randn(tuple(shape) + (2,), device="cuda")
            out = torch.empty_like(x)
            plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape)
            ms = time_gpu(lambda: plan.exec(out, x, st), 5, 20, torch)
            n = int(np.prod(shape[1:]))
            ab = 2 * x.numel() * 4
            xc = torch.view_as_complex(x[0].double().contiguous())
            want = torch.fft.fftn(xc)
            got = torch.view_as_complex(out[0].double().contiguous())
            row.update({"ms": round(ms, 5), "gflops": round(5 * n * math.log2(n) * shape[0] / ms / 1e6, 1),
                        "hbm_frac": round(ab / ms / 1e6 / peak, 4), "launches": plan.launches,
                        "rel_l2": float((got - want).norm() / want.norm()),
                        "kernels": [l.split(" n=")[0] for l in plan.describe().strip().split("\n")]})
            plan.destroy()
            try:
                cf = CuFFT(shape)
                cms = time_gpu(lambda: cf.exec(x, out, st), 5, 20, torch)
                cf.destroy()
                row.update({"cufft_ms": round(cms, 5), "ours_over_cufft": round(ms / cms, 3)})
            except Exception as e:
                row["cufft_error"] = str(e)
            del x, out
            torch.cuda.empty_cache()
        except Exception as e:
            row["error"] = "%s: %s" % (type(e).__name__, e)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
