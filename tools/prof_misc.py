#!/usr/bin/env python3
"""One workload per invocation for the kernels that tools/prof_one.py cannot reach by shape alone (profiled under ncu):
    python tools/prof_misc.py scatter|slab_fused|split|rt|c2r|jit|mgpu_slab [--steps 3]
scatter    : b200fft_exec_scatter, 512 x 512 planes, 64 local planes, 8 virtual peers on this GPU (cols_scatter_kernel)
slab_fused : b200fft_slab_exec, 512^3, one rank (slab_fused_kernel: X rows, Y + scatter, Z in one persistent kernel)
split      : (100, 16384) contiguous long axis (cols_split_a_kernel + rows_split_b_kernel)
rt         : 50000 x 1000 on the runtime-length tier (rt_axis_kernel; B200FFT_FLAG_FORCE_RT)
c2r        : half-spectrum inverse 100000 x 1024 (rows_c2r_kernel)
jit        : 50000 x 1000 on the plan-time specialised kernel (NVRTC build of rows_kernel<1000, Radices<40, 25>, ...>)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch

import b200fft


def timed(fn, steps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    what = sys.argv[1]
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 3
    st = torch.cuda.current_stream().cuda_stream
    info = {"what": what}
    if what == "scatter":
        n, G = 512, 8
        zl, yl = n // G, n // G
        x = torch.randn((zl, n, n, 2), device="cuda")
        work = torch.empty_like(x)
        recv = [torch.empty((n, yl, n, 2), device="cuda") for _ in range(G)]
        plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape, flags=b200fft.FLAG_NO_FUSED)
        info["ms"] = timed(lambda: plan.exec_scatter(recv, 3, x, work, st), steps)
        info["plan"] = plan.describe().strip().split("\n")
    elif what == "slab_fused":
        n = 512
        sp = b200fft.SlabPlan(n, 1, 0)
        nfloat = sp.recv_bytes // 4
        bufs = [torch.zeros(nfloat, device="cuda") for _ in range(2)]
        x = torch.randn((n, n, n, 2), device="cuda")
        work = torch.empty_like(x)
        k = [0]

        def call():
            sp.exec(x, work, [bufs[k[0] & 1]], k[0] & 1, st)
            k[0] += 1
        info["ms"] = timed(call, steps)
        info["plan"] = [sp.describe()]
    else:
        kw = {}
        if what == "split":
            shape, mode = (100, 16384), "c2c"
        elif what == "rt":
            shape, mode, kw = (50000, 1000), "c2c", {"flags": b200fft.FLAG_FORCE_RT}
        elif what == "jit":
            shape, mode = (50000, 1000), "c2c"
        elif what == "c2r":
            shape, mode = (100000, 1024), "c2r"
        else:
            raise SystemExit(__doc__)
        if mode == "c2r":
            x = torch.randn(shape[:-1] + (shape[-1] // 2 + 1, 2), device="cuda")
            out = torch.empty(shape + (1,), device="cuda")
            plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape, inverse=True, real_mode=b200fft.REAL_HALF)
        else:
            x = torch.randn(shape + (2,), device="cuda")
            out = torch.empty_like(x)
            plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape, **kw)
        info["ms"] = timed(lambda: plan.exec(out, x, st), steps)
        info["plan"] = plan.describe().strip().split("\n")
        info["algorithmic_gbs"] = (x.numel() + out.numel()) * 4 / info["ms"] / 1e6
    print(json.dumps(info))


if __name__ == "__main__":
    main()
