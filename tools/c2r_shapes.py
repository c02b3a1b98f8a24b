#!/usr/bin/env python3
"""Half-spectrum INVERSE (C2R) of every BASELINE shape: ms, algorithmic GB/s, round-trip error against the R2C of the
same library and numpy irfftn on one batch item. One JSON line per shape."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import numpy as np
import torch

import b200fft
from bench import time_gpu, measured_peak

peak = measured_peak()[0]
st = torch.cuda.current_stream().cuda_stream
for shape in [(500000, 128), (100000, 1024), (500000, 93), (100, 640, 480), (100, 64, 64, 64), (10, 128, 128, 128),
              (1, 256, 256, 256), (50000, 1000)]:
    cshape = shape[:-1] + (shape[-1] // 2 + 1, 2)
    real = torch.randn(shape + (1,), device="cuda")
    spec = torch.empty(cshape, device="cuda")
    fwd = b200fft.plan_fft("float32", "float32", real.shape, cshape, real_mode=b200fft.REAL_HALF)
    inv = b200fft.plan_fft("float32", "float32", cshape, real.shape, real_mode=b200fft.REAL_HALF, inverse=True)
    fwd.exec(spec, real, st)
    back = torch.empty_like(real)
    keep = spec.clone()
    ms = time_gpu(lambda: inv.exec(back, spec, st), 3, 20, torch)
    torch.cuda.synchronize()
    ab = (spec.numel() + back.numel()) * 4
    want = np.fft.irfftn(keep[0].double().cpu().numpy().view(np.complex128)[..., 0], s=shape[1:])
    got = back[0, ..., 0].double().cpu().numpy()
    print(json.dumps({"shape": list(shape), "c2r_ms": round(ms, 4), "gbs": round(ab / ms / 1e6, 1), "hbm_frac": round(ab / ms / 1e6 / peak, 3),
                      "round_trip_rel_l2": float((back - real).norm() / real.norm()),
                      "rel_l2_vs_numpy_irfftn": float(np.linalg.norm(got - want) / np.linalg.norm(want)),
                      "input_preserved": bool(torch.equal(spec, keep)),
                      "plan": [l.split(" smem")[0] for l in inv.describe().strip().split("\n")]}), flush=True)
    fwd.destroy(); inv.destroy()
    del real, spec, back, keep
    torch.cuda.empty_cache()
