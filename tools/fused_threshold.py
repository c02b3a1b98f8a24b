#!/usr/bin/env python3
"""fused (v2) vs per-axis kernels over batch size, to set the default policy: python tools/fused_threshold.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import b200fft
from bench import time_gpu

def run(shape, env):
    for k in ("B200FFT_FUSED", "B200FFT_CHUNK_MB", "B200FFT_FUSED_PREFER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    x = torch.randn(tuple(shape) + (2,), device="cuda"); out = torch.empty_like(x)
    plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape)
    st = torch.cuda.current_stream().cuda_stream
    ms = time_gpu(lambda: plan.exec(out, x, st), 5, 30, torch)
    d = plan.describe().split(":")[0][:60]
    plan.destroy(); del x, out; torch.cuda.empty_cache()
    return ms, d

for dims, batches, chunks in (((64, 64, 64), (1, 2, 4, 8, 16, 32, 50, 100, 400), (8, 12)), ((128, 128, 128), (1, 2, 3, 5, 10, 20, 50), (8, 12)),
                              ((256, 256, 256), (1, 2, 4), (8, 12))):
    for b in batches:
        row = {"dims": dims, "batch": b}
        row["per_axis_ms"] = round(run((b,) + dims, {"B200FFT_FUSED": "0"})[0], 5)
        for c in chunks:
            ms, d = run((b,) + dims, {"B200FFT_FUSED": "1", "B200FFT_CHUNK_MB": str(c)})
            row["fused_c%d_ms" % c] = round(ms, 5); row["kernel"] = d
        print(json.dumps(row), flush=True)
