#!/usr/bin/env python3
"""Long contiguous rows (4096 .. 20000 points), ~200 MB per array: the one-buffer in-place kernel (default) vs the plan with
B200FFT_ROWS_INPLACE=0 (two exchange buffers, or two split passes) vs cuFFT.  -> profiles/r2_long_rows.md"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import b200fft
from bench import CuFFT, time_gpu, measured_peak

peak, _ = measured_peak()
st = torch.cuda.current_stream().cuda_stream
for n in (4096, 5000, 6000, 8192, 8640, 10000, 12288, 15625, 16384, 20000):
    for batch in (100, (200 << 20) // (8 * n)):
        x = torch.randn(batch, n, 2, device="cuda"); out = torch.empty_like(x)
        row = {"shape": [batch, n]}
        for tag, env in (("inplace", "1"), ("before", "0")):
            os.environ["B200FFT_ROWS_INPLACE"] = env
            plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
            ms = time_gpu(lambda: plan.exec(out, x, st), 5, 30, torch)
            want = torch.fft.fft(torch.view_as_complex(x[0].double().contiguous()))
            got = torch.view_as_complex(out[0].double().contiguous())
            row[tag] = {"ms": round(ms, 5), "rel": float((got - want).norm() / want.norm()), "hbm_frac": round(2 * x.numel() * 4 / ms / 1e6 / peak, 3),
                        "plan": plan.describe().strip().split(" user stages")[0].replace("axis 0: ", "")[:90]}
            plan.destroy()
        cf = CuFFT((batch, n)); row["cufft_ms"] = round(time_gpu(lambda: cf.exec(x, out, st), 5, 30, torch), 5); cf.destroy()
        row["ours_over_cufft"] = round(row["inplace"]["ms"] / row["cufft_ms"], 3)
        print(json.dumps(row), flush=True)
        del x, out
