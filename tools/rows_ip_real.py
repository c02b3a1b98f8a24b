#!/usr/bin/env python3
"""Real input, full spectrum out, registered long rows: one-buffer kernel vs B200FFT_ROWS_INPLACE=0 (two buffers)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import b200fft
from bench import time_gpu

st = torch.cuda.current_stream().cuda_stream
for batch, n in ((25000, 2048), (12000, 2160), (25000, 4096), (6000, 4320), (3200, 8192), (1600, 16384)):
    x = torch.randn(batch, n, 1, device="cuda"); out = torch.empty(batch, n, 2, device="cuda")
    row = {"shape": [batch, n]}
    for tag, env in (("inplace", "1"), ("before", "0")):
        os.environ["B200FFT_ROWS_INPLACE"] = env
        plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape)
        ms = time_gpu(lambda: plan.exec(out, x, st), 3, 15, torch)
        want = torch.fft.fft(x[0, :, 0].double())
        got = torch.view_as_complex(out[0].double().contiguous())
        row[tag] = {"ms": round(ms, 5), "rel": float((got - want).norm() / want.norm()), "plan": plan.describe().strip().split(" user stages")[0].replace("axis 0: ", "")[:70]}
        plan.destroy()
    print(json.dumps(row), flush=True)
    del x, out
