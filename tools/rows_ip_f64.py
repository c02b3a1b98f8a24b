#!/usr/bin/env python3
"""fp64 rows on the plan-time tier: one-buffer kernel (default from 16 KB per row = 1024 fp64 points) vs two buffers
(B200FFT_ROWS_INPLACE=0) vs cuFFT Z2Z (torch.fft on complex128)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import b200fft
from bench import time_gpu

st = torch.cuda.current_stream().cuda_stream
for batch, n in ((100000, 1024), (25000, 4096), (100, 4096), (30000, 1000), (10000, 10000)):
    x = torch.randn(batch, n, 2, device="cuda", dtype=torch.float64); out = torch.empty_like(x)
    row = {"shape": [batch, n]}
    for tag, env in (("inplace", "1"), ("before", "0")):
        os.environ["B200FFT_ROWS_INPLACE"] = env
        plan = b200fft.plan_fft("float64", "float64", x.shape, x.shape)
        ms = time_gpu(lambda: plan.exec(out, x, st), 3, 10, torch)
        want = torch.fft.fft(torch.view_as_complex(x[0]))
        got = torch.view_as_complex(out[0])
        row[tag] = {"ms": round(ms, 5), "rel": float((got - want).norm() / want.norm()),
                    "plan": plan.describe().strip().split(" user stages")[0].replace("axis 0: ", "")[:80]}
        plan.destroy()
    xc = torch.view_as_complex(x)
    row["torch_cufft_z2z_ms"] = round(time_gpu(lambda: torch.fft.fft(xc, out=torch.view_as_complex(out)), 3, 10, torch), 5)
    print(json.dumps(row), flush=True)
    del x, out, xc
