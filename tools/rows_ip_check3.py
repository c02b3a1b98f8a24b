#!/usr/bin/env python3
"""2048 / 2160 / 4320-point rows: the registered two-buffer kernels vs in-place variants (B200FFT_PREFER) vs cuFFT, and the
published 2-D shapes whose contiguous axis they serve."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import b200fft
from bench import CuFFT, time_gpu

st = torch.cuda.current_stream().cuda_stream
CASES = [((100, 2048), ("", "rowsIP2048_32x8x8", "rowsIP2048_16x16x8")), ((12800, 2048), ("", "rowsIP2048_32x8x8", "rowsIP2048_16x16x8")),
         ((12000, 2160), ("", "rowsIP2160_16x15x9_c1_t144", "rowsIP2160_16x15x9_c1_t288")), ((6000, 4320), ("", "rowsIP4320")),
         ((1, 3840, 2160), ("", "rowsIP2160_16x15x9_c1_t144", "rowsIP2160_16x15x9_c1_t288")), ((1, 7680, 4320), ("", "rowsIP4320"))]
for shape, prefs in CASES:
    x = torch.randn(*shape, 2, device="cuda"); out = torch.empty_like(x)
    row = {"shape": list(shape)}
    for pref in prefs:
        os.environ.pop("B200FFT_PREFER", None)
        if pref: os.environ["B200FFT_PREFER"] = pref
        plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
        ms = round(time_gpu(lambda: plan.exec(out, x, st), 5, 30, torch), 5)
        xc = torch.view_as_complex(x[0].double().contiguous())
        want = torch.fft.fftn(xc) if len(shape) == 3 else torch.fft.fft(xc)
        got = torch.view_as_complex(out[0].double().contiguous())
        row[plan.describe().strip().split(" n=")[0].split(": ")[1]] = [ms, float((got - want).norm() / want.norm())]
        plan.destroy()
    os.environ.pop("B200FFT_PREFER", None)
    cf = CuFFT(shape); row["cufft_ms"] = round(time_gpu(lambda: cf.exec(x, out, st), 5, 30, torch), 5); cf.destroy()
    print(json.dumps(row), flush=True)
    del x, out
