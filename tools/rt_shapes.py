#!/usr/bin/env python3
"""Lengths WITHOUT a registered variant (runtime-length tier, csrc/rt.cu) vs cuFFT: python tools/rt_shapes.py"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import numpy as np
import torch
import b200fft
from bench import CuFFT, time_gpu, measured_peak

SHAPES = [(200000, 100), (50000, 1000), (100000, 243), (20000, 2000), (4000, 10000), (200, 300, 300), (20, 1000, 1000),
          (50, 100, 100, 100), (4, 200, 200, 200), (100000, 210)]
if os.environ.get("RT_SHAPES"):   # e.g. RT_SHAPES="50000x1000;20x1000x1000"
    SHAPES = [tuple(int(v) for v in t.split("x")) for t in os.environ["RT_SHAPES"].split(";")]
peak, _ = measured_peak()
st = torch.cuda.current_stream().cuda_stream
for shape in SHAPES:
    row = {"shape": list(shape)}
    try:
        x = torch.randn(tuple(shape) + (2,), device="cuda"); out = torch.empty_like(x)
        plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape)
        ms = time_gpu(lambda: plan.exec(out, x, st), 3, 10, torch)
        xc = torch.view_as_complex(x[0].double().contiguous())
        got = torch.view_as_complex(out[0].double().contiguous()); want = torch.fft.fftn(xc)
        row.update({"ms": round(ms, 4), "hbm_frac": round(2 * x.numel() * 4 / ms / 1e6 / peak, 3),
                    "rel_l2": float((got - want).norm() / want.norm()),
                    "kernels": [l.split(" smem")[0] for l in plan.describe().strip().split("\n")]})
        plan.destroy()
        cf = CuFFT(shape); cms = time_gpu(lambda: cf.exec(x, out, st), 3, 10, torch); cf.destroy()
        row.update({"cufft_ms": round(cms, 4), "ours_over_cufft": round(ms / cms, 2)})
        del x, out; torch.cuda.empty_cache()
    except Exception as e:
        row["error"] = str(e)
    print(json.dumps(row), flush=True)
