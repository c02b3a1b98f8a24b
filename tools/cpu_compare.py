#!/usr/bin/env python3
"""Python CPU-comparison harness: the other CPU libraries next to the reference's CPU path and to this library.

The reference ships `benchmark-cpu-others/benchmark.py` (NumPy / SciPy-pocketfft / PyFFTW `fftn` over its benchmark
shapes, complex64, one thread and all cores). This is the equivalent for this repo, with three differences: PyFFTW is
not in the image (NumPy and SciPy only), every CPU leg runs on a BOUNDED batch sample and is extrapolated linearly to
the full batch (the legs are embarrassingly parallel over the batch; `sample_batch` is printed), and two more legs are
timed on the same data:

  * `oracle`  — the C++ restatement of the reference's own CPU radix-n path (oracle/ref_fft.cpp, workers = all cores),
  * `b200fft` — this library through its ctypes shim with HOST buffers (b200fft_exec_host: H2D + kernels + D2H), only
                when a GPU is present; there is no CPU path in the product, so without a GPU the leg reports null.

    python tools/cpu_compare.py [--budget-s 2.0] [--shapes 1d,2d,3d] [--out gpurun_out/cpu_compare.jsonl]

One JSON line per shape: milliseconds for the FULL batch per leg, threads used, sample size, and the relative L2 error
of each leg against numpy float64 on the first transform (the legs must agree before their times are compared).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import numpy as np
import scipy.fft
from threadpoolctl import threadpool_limits

SHAPES = {
    "1d": [(500000, 128), (100000, 1024), (500000, 93)],
    "2d": [(100, 640, 480)],
    "3d": [(100, 64, 64, 64), (10, 128, 128, 128), (1, 256, 256, 256)],
    "big": [(1, 512, 512, 512)],
}


def best_ms(fn, budget_s, min_reps=2, max_reps=20):
    """Best of several runs, stopping when the time budget is used up."""
    best, spent, reps = float("inf"), 0.0, 0
    while reps < min_reps or (spent < budget_s and reps < max_reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best, spent, reps = min(best, dt), spent + dt, reps + 1
    return best * 1e3


def sample_batch(shape, target_points=4_000_000):
    """Batch items to time on the CPU: about `target_points` complex points, at least one transform."""
    per = int(np.prod(shape[1:]))
    return int(max(1, min(shape[0], target_points // per)))


def rel_l2(got, want):
    return float(np.linalg.norm(got - want) / np.linalg.norm(want))


def run_shape(shape, threads, budget_s, gpu):
    rng = np.random.default_rng(1234)
    nb = sample_batch(shape)
    axes = tuple(range(1, len(shape)))
    x = rng.standard_normal((nb,) + tuple(shape[1:]) + (2,)).astype(np.float32)
    xc = x[..., 0] + 1j * x[..., 1]                      # complex64, like the reference's harness
    scale = shape[0] / nb
    want0 = np.fft.fftn(xc[0].astype(np.complex128))
    row = {"shape": list(shape), "sample_batch": nb, "threads": threads, "extrapolation": "linear in the batch",
           "ms_full_batch": {}, "rel_l2_vs_numpy_f64": {}}

    def leg(name, fn, first):
        try:
            row["ms_full_batch"][name] = best_ms(fn, budget_s) * scale
            row["rel_l2_vs_numpy_f64"][name] = rel_l2(first(), want0)
        except Exception as e:  # a missing leg is reported, never silently dropped
            row["ms_full_batch"][name] = None
            row["rel_l2_vs_numpy_f64"][name] = "%s: %s" % (type(e).__name__, e)

    with threadpool_limits(limits=1):
        leg("numpy_1thread", lambda: np.fft.fftn(xc, axes=axes), lambda: np.fft.fftn(xc[:1], axes=axes)[0])
    leg("scipy_1thread", lambda: scipy.fft.fftn(xc, axes=axes, workers=1), lambda: scipy.fft.fftn(xc[:1], axes=axes)[0])
    leg("scipy_all_threads", lambda: scipy.fft.fftn(xc, axes=axes, workers=threads),
        lambda: scipy.fft.fftn(xc[:1], axes=axes, workers=threads)[0])

    import oracle
    def oracle_first():
        o = oracle.ref_fft(x[:1])
        return o[0, ..., 0] + 1j * o[0, ..., 1]
    rplan = oracle.RefPlan(x.shape, x.dtype)  # plan outside the timed calls, like fft/bench.mojo:83-90
    rout = np.empty(rplan.out_shape, np.float32)
    leg("oracle_reference_cpu_path_all_threads", lambda: rplan.exec(rout, x, workers=threads), oracle_first)
    rplan.destroy()

    if gpu:
        import b200fft
        try:
            plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
            out = np.empty_like(x)
            leg("b200fft_host_buffers", lambda: plan.exec_host(out, x), lambda: (plan.exec_host(out, x), out[0, ..., 0] + 1j * out[0, ..., 1])[1])
            plan.destroy()
        except Exception as e:
            row["ms_full_batch"]["b200fft_host_buffers"] = None
            row["rel_l2_vs_numpy_f64"]["b200fft_host_buffers"] = "%s: %s" % (type(e).__name__, e)
    else:
        row["ms_full_batch"]["b200fft_host_buffers"] = None
        row["rel_l2_vs_numpy_f64"]["b200fft_host_buffers"] = "no GPU in this process: the product has no CPU path"
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget-s", type=float, default=2.0, help="time budget per leg and shape")
    ap.add_argument("--shapes", default="1d,2d,3d")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cpu_compare.jsonl"))
    a = ap.parse_args()
    threads = len(os.sched_getaffinity(0))
    try:
        import torch
        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for key in a.shapes.split(","):
            for shape in SHAPES[key]:
                line = json.dumps(run_shape(shape, threads, a.budget_s, gpu))
                print(line)
                f.write(line + "\n")


if __name__ == "__main__":
    main()
