#!/usr/bin/env python3
"""ONE process, all visible GPUs, through the C ABI's multi-device entry points (b200fft_mgpu_*):
    python tools/mgpu_check.py [--size 512] [--gpus N]
1. slab mode, device-resident: parity of every device's Y slab against torch float64 (256^3) and timing at --size^3
   (CUDA events on every slot's stream, max over slots; exchange included);
2. exec_host: the whole job from / to pinned host memory — slab volume and the batch-sharded primary workload
   (N x 100000 x 1024), wall clock around the call. One JSON line per measurement."""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import numpy as np
import torch

import b200fft


def slab_parity(n, devs):
    G = len(devs)
    zl, yl = n // G, n // G
    lay = (1, n, n, n, 2)
    plan = b200fft.MgpuPlan("float32", "float32", lay, lay, devices=devs, mode=b200fft.MGPU_SLAB)
    g = torch.Generator().manual_seed(7)
    vol = torch.randn((n, n, n, 2), generator=g)
    ins = [vol[i * zl:(i + 1) * zl].to("cuda:%d" % devs[i]) for i in range(G)]
    outs = [torch.full((n, yl, n, 2), float("nan"), device="cuda:%d" % devs[i]) for i in range(G)]
    for d in devs:
        torch.cuda.synchronize(d)
    plan.exec(outs, ins)
    plan.exec(outs, ins)
    plan.synchronize()
    want = torch.fft.fftn(torch.view_as_complex(vol.double().contiguous().to("cuda:%d" % devs[0])))
    errs = []
    for h in range(G):
        got = torch.view_as_complex(outs[h].double().contiguous().to("cuda:%d" % devs[0]))
        ref = want[:, h * yl:(h + 1) * yl, :]
        errs.append(float((got - ref).norm() / ref.norm()))
    plan.destroy()
    return errs


def slab_time(n, devs, steps=10, warmup=3):
    G = len(devs)
    zl, yl = n // G, n // G
    lay = (1, n, n, n, 2)
    plan = b200fft.MgpuPlan("float32", "float32", lay, lay, devices=devs, mode=b200fft.MGPU_SLAB)
    ins = [torch.randn((zl, n, n, 2), device="cuda:%d" % d) for d in devs]
    outs = [torch.empty((n, yl, n, 2), device="cuda:%d" % d) for d in devs]
    streams = [torch.cuda.ExternalStream(plan.stream(i), device="cuda:%d" % devs[i]) for i in range(G)]
    for _ in range(warmup):
        plan.exec(outs, ins)
    plan.synchronize()
    ev = []
    for i, d in enumerate(devs):
        with torch.cuda.device(d):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(streams[i])
            ev.append((e0, e1))
    for _ in range(steps):
        plan.exec(outs, ins)
    for i, d in enumerate(devs):
        with torch.cuda.device(d):
            ev[i][1].record(streams[i])
    plan.synchronize()
    ms = max(e0.elapsed_time(e1) for e0, e1 in ev) / steps
    text = plan.describe()
    plan.destroy()
    return ms, text


def host_time(plan, h_out, h_in, reps=3):
    plan.exec_host(h_out, h_in)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        plan.exec_host(h_out, h_in)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--batch-per-gpu", type=int, default=100000)
    ap.add_argument("--slab-only", action="store_true", help="parity at 256^3 and device-resident timing only")
    args = ap.parse_args()
    G = args.gpus or torch.cuda.device_count()
    devs = list(range(G))
    errs = slab_parity(256, devs)
    print(json.dumps({"what": "mgpu slab parity 256^3", "gpus": G, "rel_l2_vs_torch_f64_per_device": errs,
                      "ok": bool(max(errs) < 2e-6)}), flush=True)
    n = args.size
    ms, text = slab_time(n, devs)
    print(json.dumps({"what": "mgpu slab %d^3 device-resident" % n, "gpus": G, "ms": round(ms, 4),
                      "gflops": round(5.0 * n ** 3 * math.log2(n ** 3) / ms / 1e6, 1), "plan": text.strip().split("\n")[:1],
                      "env": {k: v for k, v in os.environ.items() if k.startswith("B200FFT_")}}), flush=True)
    if args.slab_only:
        return
    # host paths
    lay = (1, n, n, n, 2)
    plan = b200fft.MgpuPlan("float32", "float32", lay, lay, devices=devs, mode=b200fft.MGPU_SLAB)
    h_in = torch.randn((n, n, n, 2)).pin_memory()
    h_out = torch.empty((n, n, n, 2)).pin_memory()
    ms = host_time(plan, h_out, h_in)
    chk = torch.fft.fftn(torch.view_as_complex(h_in[:, :8, :8].double().contiguous()))  # cheap sanity only on a corner is meaningless
    print(json.dumps({"what": "mgpu slab %d^3 exec_host (natural order in and out)" % n, "gpus": G, "ms": round(ms, 3),
                      "bytes_each_way": int(h_in.numel() * 4)}), flush=True)
    plan.destroy()
    del h_in, h_out
    B = args.batch_per_gpu * G
    lay = (B, 1024, 2)
    plan = b200fft.MgpuPlan("float32", "float32", lay, lay, devices=devs, mode=b200fft.MGPU_BATCH_SHARD)
    h_in = torch.randn(lay).pin_memory()
    h_out = torch.empty(lay).pin_memory()
    ms = host_time(plan, h_out, h_in)
    k = 5
    want = torch.fft.fft(torch.view_as_complex(h_in[-k:].double().contiguous()), dim=1)
    got = torch.view_as_complex(h_out[-k:].double().contiguous())
    print(json.dumps({"what": "mgpu batch-shard exec_host %d x 1024" % B, "gpus": G, "ms": round(ms, 3),
                      "gflops_e2e": round(5.0 * 1024 * 10 * B / ms / 1e6, 1), "bytes_each_way": int(h_in.numel() * 4),
                      "rel_l2_last_rows": float((got - want).norm() / want.norm())}), flush=True)
    plan.destroy()


if __name__ == "__main__":
    main()
