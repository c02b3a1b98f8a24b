"""torchrun worker of tests/test_gpu_multirank.py: the slab decomposition on REAL ranks (one process per GPU, NCCL +
CUDA IPC over NVLink) — every exchange mode, forward and inverse, checked with b200fft.verify's distributed checks
(sampled bins on random data, every element on the impulse volume) and, at the smallest size, against torch.fft.fftn
of the gathered volume. Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch
import torch.distributed as dist

from b200fft import verify
from b200fft.slab import SlabFFT3D


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "128,256").split(",")]
    res = {"world": world, "cases": []}
    for n in sizes:
        dims = (n, n, n)
        zl, yl = n // world, n // world
        g = torch.Generator(device="cuda").manual_seed(500 + rank)
        x = torch.randn((zl, n, n, 2), generator=g, device="cuda")
        for mode in ("fused", "p2p", "nccl"):
            for inverse in (False, True):
                case = {"n": n, "mode": mode, "inverse": inverse}
                try:
                    sl = SlabFFT3D(dims, exchange=mode, inverse=inverse)
                    outs = [sl.forward(x).clone() for _ in range(3)]        # both receive slabs + reuse of the first
                    case["repeat_identical"] = bool(all(torch.equal(outs[0], o) for o in outs[1:]))
                    case["bins"] = verify.sampled_bins_check(x, outs[0], dims, rank, world, inverse=inverse)
                    d = sl.forward(verify.delta_input(dims, rank, world, x.device)).clone()
                    case["delta"] = verify.delta_volume_check(d, dims, rank, world, inverse=inverse)
                    if n == sizes[0]:
                        full = [torch.empty_like(x) for _ in range(world)]
                        dist.all_gather(full, x)
                        vc = torch.view_as_complex(torch.cat(full, 0).double().contiguous())
                        want = (torch.fft.ifftn(vc) if inverse else torch.fft.fftn(vc))[:, rank * yl:(rank + 1) * yl, :]
                        got = torch.view_as_complex(outs[0].double().contiguous())
                        e = torch.stack([(got - want).abs().pow(2).sum(), want.abs().pow(2).sum()])
                        dist.all_reduce(e)
                        case["rel_l2_vs_torch_f64"] = float((e[0] / e[1]).sqrt())
                    if mode == "fused":
                        case["peer_wait_timeouts"] = sl.timeouts()
                    sl.close()
                except Exception as ex:
                    case["error"] = "%s: %s" % (type(ex).__name__, ex)
                res["cases"].append(case)
    if rank == 0:
        print("MULTIRANK " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
