"""World-size-2 CPU test (gloo) of the slab decomposition's host logic: SlabFFT3D's block order,
all_to_all_single usage and result placement, with the three local device operations replaced by
a numpy engine defined HERE (test infrastructure; the product engine is CUDA-only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class NumpyEngine:
    """Restates, on the CPU, what CudaSlabEngine's kernels do (spec: include/b200fft.h, exec_scatter)."""

    def __init__(self, dims, world):
        self.Z, self.Y, self.X = dims
        self.zl, self.yl = self.Z // world, self.Y // world

    def alloc(self, shape, shared=False):
        return torch.zeros(shape, dtype=torch.float32)

    def fft2_scatter(self, x_local, targets, zrank):
        x = x_local.numpy().astype(np.float64)
        f = np.fft.fft2(x[..., 0] + 1j * x[..., 1], axes=(1, 2))
        for z in range(self.zl):
            for y in range(self.Y):
                t = targets[y // self.yl].view(-1, self.yl, self.X, 2)
                row = f[z, y]
                t[zrank * self.zl + z, y % self.yl, :, 0] = torch.from_numpy(row.real.astype(np.float32))
                t[zrank * self.zl + z, y % self.yl, :, 1] = torch.from_numpy(row.imag.astype(np.float32))

    def fftz(self, slab):
        s = slab.numpy().astype(np.float64)
        f = np.fft.fft(s[..., 0] + 1j * s[..., 1], axis=0)
        slab[..., 0] = torch.from_numpy(f.real.astype(np.float32))
        slab[..., 1] = torch.from_numpy(f.imag.astype(np.float32))
        return slab


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dims, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "hackathon-fft_b200", "python")]
    from b200fft.slab import SlabFFT3D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Z, Y, X = dims
    rng = np.random.default_rng(42)
    full = rng.standard_normal((Z, Y, X, 2)).astype(np.float32)
    zl = Z // world
    x_local = torch.from_numpy(full[rank * zl:(rank + 1) * zl].copy())
    slab = SlabFFT3D(dims, exchange="nccl", engine=NumpyEngine(dims, world))
    out = slab.forward(x_local)
    out2 = slab.forward(x_local)   # plan reuse
    assert torch.equal(out, out2)
    np.save(os.path.join(out_dir, "out_%d.npy" % rank), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dims", [(8, 12, 6), (4, 4, 10)])
def test_slab_decomposition_world2_gloo(tmp_path, dims):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dims, str(tmp_path)), nprocs=world, join=True)
    Z, Y, X = dims
    rng = np.random.default_rng(42)
    full = rng.standard_normal((Z, Y, X, 2)).astype(np.float32).astype(np.float64)
    want = np.fft.fftn(full[..., 0] + 1j * full[..., 1])
    yl = Y // world
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "out_%d.npy" % r)).astype(np.float64)
        got = got[..., 0] + 1j * got[..., 1]
        assert got.shape == (Z, yl, X)
        ref = want[:, r * yl:(r + 1) * yl, :]      # rank r holds the Y slab, all z, all x
        assert np.linalg.norm(got - ref) <= 2e-6 * np.linalg.norm(ref)


def _verify_worker(rank, world, port, dims, out_dir):
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "hackathon-fft_b200", "python")]
    from b200fft.slab import SlabFFT3D
    from b200fft import verify
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Z, Y, X = dims
    zl, yl = Z // world, Y // world
    x = torch.from_numpy(np.random.default_rng(7 + rank).standard_normal((zl, Y, X, 2)).astype(np.float32))
    slab = SlabFFT3D(dims, exchange="nccl", engine=NumpyEngine(dims, world))
    out = slab.forward(x).clone()
    res = {"good": verify.sampled_bins_check(x, out, dims, rank, world)}
    d = verify.delta_input(dims, rank, world, "cpu")
    dout = slab.forward(d).clone()
    res["good_delta"] = verify.delta_volume_check(dout, dims, rank, world)

    def energy(t):
        e = t.double().pow(2).sum().reshape(1)
        dist.all_reduce(e)
        return float(e)

    # (a) rank 0's block of z planes and rank 1's land in each other's slot of the receive slab (a wrong zbase)
    bad = torch.cat([out[zl:2 * zl], out[:zl], out[2 * zl:]], 0)
    dbad = torch.cat([dout[zl:2 * zl], dout[:zl], dout[2 * zl:]], 0)
    res["swapped_z_blocks"] = verify.sampled_bins_check(x, bad, dims, rank, world)
    res["swapped_z_blocks_delta"] = verify.delta_volume_check(dbad, dims, rank, world)
    res["swapped_parseval_ratio"] = energy(bad) / energy(out)
    # (b) ONE y row of ONE rank lands in the neighbouring row's place: only the full-coverage check must see it
    one = dout.clone()
    if rank == world - 1:
        one[:, [0, 1]] = one[:, [1, 0]]
    res["one_row_delta"] = verify.delta_volume_check(one, dims, rank, world)
    if rank == 0:
        with open(os.path.join(out_dir, "verify.json"), "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dims", [(2, (8, 12, 6)), (4, (16, 8, 10))])
def test_distributed_checks_see_misplaced_blocks_that_parseval_cannot(tmp_path, world, dims):
    """b200fft.verify (what bench.py's slab leg and the multi-GPU test report): ~1e-7 on a correct result, O(1) when
    blocks are permuted although the energy is unchanged."""
    import json
    mp.spawn(_verify_worker, args=(world, _free_port(), dims, str(tmp_path)), nprocs=world, join=True)
    res = json.load(open(os.path.join(str(tmp_path), "verify.json")))
    assert res["good"] < 2e-6 and res["good_delta"] < 2e-6, res
    assert abs(res["swapped_parseval_ratio"] - 1.0) < 1e-12          # Parseval is blind to it
    assert res["swapped_z_blocks"] > 0.1 and res["swapped_z_blocks_delta"] > 0.1, res
    assert res["one_row_delta"] > 0.1, res


def test_slab_rejects_indivisible_dims_and_cpu_product_path():
    import b200fft
    from b200fft.slab import SlabFFT3D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        if not torch.cuda.is_available():
            with pytest.raises(b200fft.B200FFTError):   # no CPU fallback in the product engine
                SlabFFT3D((8, 8, 8), exchange="nccl")
        with pytest.raises(ValueError):
            SlabFFT3D((8, 8, 8), exchange="mpi", engine=NumpyEngine((8, 8, 8), 1))
    finally:
        dist.destroy_process_group()
