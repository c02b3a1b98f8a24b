#!/usr/bin/env python3
"""Experiment: does running the axis passes per L2-sized chunk of batch items (instead of
whole-array passes) cut HBM traffic enough to pay for the extra launches? Captures the
chunked launch sequence in a CUDA graph and times it against the whole-array plan."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import b200fft


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best


def run(shape, chunks):
    x = torch.randn(tuple(shape) + (2,), device="cuda")
    out = torch.empty_like(x)
    rank = len(shape) - 1
    whole = b200fft.plan_fft("float32", "float32", x.shape, out.shape)
    st = torch.cuda.current_stream().cuda_stream
    base = timeit(lambda: whole.exec(out, x, st))
    print("shape %s whole-array passes: %.4f ms" % (shape, base))
    ref = out.clone()
    for cb in chunks:
        if cb > shape[0]:
            continue
        sub = (cb,) + tuple(shape[1:]) + (2,)
        plans = []
        for a in range(rank - 1, -1, -1):
            plans.append(b200fft.plan_fft("float32", "float32", sub, sub, axis_mask=1 << a))
        per = x[0].numel() * 4
        s = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            def body():
                cs = torch.cuda.current_stream().cuda_stream
                for b0 in range(0, shape[0] - cb + 1, cb):
                    for i, p in enumerate(plans):
                        src = (x if i == 0 else out).data_ptr() + b0 * per
                        p.exec(out.data_ptr() + b0 * per, src, cs)
            body()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                body()
        t = timeit(lambda: g.replay())
        ok = torch.allclose(out[: (shape[0] // cb) * cb], ref[: (shape[0] // cb) * cb], rtol=1e-4, atol=1e-3)
        print("   chunk %3d items (%.1f MB): %.4f ms  speedup %.2fx  ok=%s" % (cb, cb * per / 1e6, t, base / t, ok))


if __name__ == "__main__":
    run((100, 640, 480), [2, 4, 5, 10, 20, 25])
    run((100, 64, 64, 64), [2, 4, 5, 10, 20, 25])
    run((10, 128, 128, 128), [1, 2, 5])
