#!/usr/bin/env python3
"""A/B of the half-spectrum R2C row kernels in one process: register unpack (default) vs the shared-memory unpack
(B200FFT_R2C_SMEM=1), and alternative row variants (B200FFT_PREFER), on the BASELINE shapes. One JSON line per run
-> gpurun_out/r2c_ab.jsonl. Parity of every run: relative L2 vs torch.fft.rfftn (float64) on the first batch items."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch

import b200fft
import bench

RUNS = [
    # (shape, env)
    ((100000, 1024), {}),
    ((100000, 1024), {"B200FFT_R2C_SMEM": "1"}),
    ((100000, 1024), {"B200FFT_PREFER": "rows512_32x16_c16"}),
    ((100000, 1024), {"B200FFT_PREFER": "rows512_32x16_c16", "B200FFT_R2C_SMEM": "1"}),
    ((100000, 1024), {"B200FFT_PREFER": "rows512_8x8x8"}),
    ((500000, 128), {}),
    ((500000, 128), {"B200FFT_R2C_SMEM": "1"}),
    ((500000, 128), {"B200FFT_PREFER": "rows64_8x8"}),
    ((500000, 93), {}),
    ((100, 64, 64, 64), {}),
    ((100, 64, 64, 64), {"B200FFT_R2C_SMEM": "1"}),
    ((10, 128, 128, 128), {}),
    ((1, 256, 256, 256), {}),
    ((100, 640, 480), {}),
    ((100, 640, 480), {"B200FFT_R2C_SMEM": "1"}),
]


def main():
    out_path = os.path.join(ROOT, "gpurun_out", "r2c_ab.jsonl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    stream = torch.cuda.current_stream().cuda_stream
    with open(out_path, "w") as f:
        for shape, env in RUNS:
            for k in ("B200FFT_R2C_SMEM", "B200FFT_PREFER"):
                os.environ.pop(k, None)
            os.environ.update(env)
            g = torch.Generator(device="cuda").manual_seed(7)
            x = torch.randn(tuple(shape) + (1,), generator=g, device="cuda")
            oshape = tuple(shape[:-1]) + (shape[-1] // 2 + 1, 2)
            out = torch.full(oshape, float("nan"), device="cuda")
            row = {"shape": list(shape), "env": env}
            try:
                plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape, real_mode=b200fft.REAL_HALF)
                row["ms"] = bench.time_gpu(lambda: plan.exec(out, x, stream), 3, 20, torch)
                nb = min(shape[0], 4)
                want = torch.fft.rfftn(x[:nb, ..., 0].double(), dim=tuple(range(1, len(shape))))
                got = torch.view_as_complex(out[:nb].double().contiguous())
                row["rel_l2"] = float((got - want).norm() / want.norm())
                row["finite"] = bool(torch.isfinite(out).all())
                row["kernels"] = [k.split(" n=")[0] for k in plan.describe().strip().split("\n")]
                plan.destroy()
            except Exception as e:
                row["error"] = "%s: %s" % (type(e).__name__, e)
            line = json.dumps(row)
            print(line)
            f.write(line + "\n")
            del x, out
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
