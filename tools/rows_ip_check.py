#!/usr/bin/env python3
"""(100, 16384) and neighbours: one in-place launch (rows_ip_kernel) vs the two split passes (B200FFT_ROWS_INPLACE=0) vs cuFFT."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import b200fft
from bench import CuFFT, time_gpu

st = torch.cuda.current_stream().cuda_stream
for batch in (1, 100, 148, 296, 1000, 10000):
    x = torch.randn(batch, 16384, 2, device="cuda"); out = torch.empty_like(x)
    row = {"shape": [batch, 16384]}
    for tag, env in (("inplace", "1"), ("split", "0")):
        os.environ["B200FFT_ROWS_INPLACE"] = env
        plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
        row[tag + "_ms"] = round(time_gpu(lambda: plan.exec(out, x, st), 5, 30, torch), 5)
        want = torch.fft.fft(torch.view_as_complex(x[0].double().contiguous()))
        got = torch.view_as_complex(out[0].double().contiguous())
        row[tag + "_rel"] = float((got - want).norm() / want.norm())
        row[tag + "_desc"] = plan.describe().strip().split("\n")[0][:60]
        plan.destroy()
    cf = CuFFT((batch, 16384)); row["cufft_ms"] = round(time_gpu(lambda: cf.exec(x, out, st), 5, 30, torch), 5); cf.destroy()
    print(json.dumps(row), flush=True)
