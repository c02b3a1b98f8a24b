#!/usr/bin/env python3
"""torchrun entry: slab-decomposed 3-D FFT across the node's GPUs, both exchange modes.

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/slab_check.py [--size 512]

1. correctness at 128^3: every rank's Y slab vs torch.fft.fftn of the gathered volume (rank 0 prints)
2. timing at N^3 (default 512): K calls bracketed by barriers, CUDA events, max over ranks,
   for exchange = p2p (fused scattering store over CUDA IPC / NVLink) and nccl (pack + all_to_all_single).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import torch.distributed as dist

from b200fft.slab import SlabFFT3D


def check(n, mode, rank, world):
    zl, yl = n // world, n // world
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    x = torch.randn((zl, n, n, 2), generator=g, device="cuda")
    slab = SlabFFT3D((n, n, n), exchange=mode)
    out = slab.forward(x).clone()
    out2 = slab.forward(x).clone()      # second call uses the other receive slab
    full = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(full, x)
    vol = torch.cat(full, 0)
    want = torch.fft.fftn(torch.view_as_complex(vol.double().contiguous()))[:, rank * yl:(rank + 1) * yl, :]
    got = torch.view_as_complex(out.double().contiguous())
    err = float((got - want).norm() / want.norm())
    same = bool(torch.equal(out, out2))
    errs = [None] * world
    dist.all_gather_object(errs, (err, same))
    slab.close()
    return errs


def check_restore_and_pencil(n, rank, world):
    """Natural-order output (second exchange), forward -> inverse round trip, and the pencil decomposition on the
    most square grid the world size allows; relative L2 per rank, gathered."""
    from b200fft.pencil import PencilFFT3D
    zl = n // world
    g = torch.Generator(device="cuda").manual_seed(200 + rank)
    x = torch.randn((zl, n, n, 2), generator=g, device="cuda")
    full = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(full, x)
    want = torch.fft.fftn(torch.view_as_complex(torch.cat(full, 0).double().contiguous()))
    fwd, inv = SlabFFT3D((n, n, n), exchange="p2p"), SlabFFT3D((n, n, n), exchange="p2p", inverse=True)
    spec = fwd.forward(x, natural=True)
    e_nat = float((torch.view_as_complex(spec.double().contiguous()) - want[rank * zl:(rank + 1) * zl]).norm()
                  / want[rank * zl:(rank + 1) * zl].norm())
    back = inv.forward(spec, natural=True)
    e_rt = float((back - x).norm() / x.norm())
    fwd.close()
    inv.close()
    P0 = max(d for d in (1, 2, 4, 8) if world % d == 0 and d * d <= world)
    P1 = world // P0
    p0, p1 = divmod(rank, P1)
    vol = torch.cat(full, 0)
    zp, yp = n // P0, n // P1
    xl = vol[p0 * zp:(p0 + 1) * zp, p1 * yp:(p1 + 1) * yp].contiguous()
    pen = PencilFFT3D((n, n, n), (P0, P1))
    got = torch.view_as_complex(pen.forward(xl).double().contiguous())
    ref = want[:, p0 * (n // P0):(p0 + 1) * (n // P0), p1 * (n // P1):(p1 + 1) * (n // P1)]
    e_pen = float((got - ref).norm() / ref.norm())
    pen.close()
    errs = [None] * world
    dist.all_gather_object(errs, (e_nat, e_rt, e_pen))
    return {"natural_out_max_rel_l2": max(e[0] for e in errs), "round_trip_max_rel_l2": max(e[1] for e in errs),
            "pencil_grid": [P0, P1], "pencil_max_rel_l2": max(e[2] for e in errs)}


def bench(n, mode, rank, world, steps, warmup):
    zl = n // world
    x = torch.randn((zl, n, n, 2), device="cuda")
    slab = SlabFFT3D((n, n, n), exchange=mode)
    for _ in range(warmup):
        slab.forward(x)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        slab.forward(x)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    slab.close()
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--modes", default="nccl,p2p,fused")
    ap.add_argument("--delays", default="", help="comma list of B200FFT_SLAB_DELAY values to time the fused mode with")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = {"world": world, "n": a.n}
    for mode in a.modes.split(","):
        try:
            errs = check(128, mode, rank, world)
            res["check128_" + mode] = {"max_rel_l2": max(e for e, _ in errs), "repeat_identical": all(s for _, s in errs)}
            res["ms_%d_%s" % (a.n, mode)] = bench(a.n, mode, rank, world, a.steps, a.warmup)
        except Exception as e:  # report, keep going with the other modes
            res["error_" + mode] = "%s: %s" % (type(e).__name__, e)
    try:
        res["check128_restore_pencil"] = check_restore_and_pencil(128, rank, world)
    except Exception as e:
        res["error_restore_pencil"] = "%s: %s" % (type(e).__name__, e)
    for d in [v for v in a.delays.split(",") if v]:
        os.environ["B200FFT_SLAB_DELAY"] = d
        try:
            res["ms_%d_fused_delay%s" % (a.n, d)] = bench(a.n, "fused", rank, world, a.steps, a.warmup)
        except Exception as e:
            res["error_fused_delay" + d] = "%s: %s" % (type(e).__name__, e)
    if rank == 0:
        import math
        pts = a.n ** 3
        for k in [k for k in res if k.startswith("ms_")]:
            res["gflops" + k[2:]] = 5 * pts * math.log2(pts) / res[k] / 1e6
        line = json.dumps(res)
        print(line)
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "slab_check_w%d.jsonl" % world), "a") as f:
            f.write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
