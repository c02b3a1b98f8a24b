"""World-size-4 (2x2, 4x1, 1x4) and world-size-2 CPU tests (gloo) of the host logic around the multi-rank 3-D
transforms: PencilFFT3D's grid, sub-groups, block order and result placement, and SlabFFT3D.restore()'s second
exchange (natural-order output, forward -> inverse round trip). The local device operations are replaced by
numpy engines defined in the tests (test infrastructure; the product engines are CUDA-only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_slab_gloo import NumpyEngine as NumpySlabEngine


class NumpyPencilEngine:
    """Restates CudaPencilEngine on the CPU: in-place transform of one axis of a [a][b][c][2] array."""

    def __init__(self, inverse=False):
        self.inverse = inverse

    def alloc(self, shape):
        return torch.zeros(shape, dtype=torch.float32)

    def fft_axis(self, t, axis):
        s = t.numpy().astype(np.float64)
        c = s[..., 0] + 1j * s[..., 1]
        f = np.fft.ifft(c, axis=axis) if self.inverse else np.fft.fft(c, axis=axis)
        t[..., 0] = torch.from_numpy(f.real.astype(np.float32))
        t[..., 1] = torch.from_numpy(f.imag.astype(np.float32))
        return t


class NumpyInverseSlabEngine(NumpySlabEngine):
    def fft2_scatter(self, x_local, targets, zrank):
        x = x_local.numpy().astype(np.float64)
        f = np.fft.ifft2(x[..., 0] + 1j * x[..., 1], axes=(1, 2))
        for y in range(self.Y):
            t = targets[y // self.yl].view(-1, self.yl, self.X, 2)
            t[zrank * self.zl:(zrank + 1) * self.zl, y % self.yl, :, 0] = torch.from_numpy(f[:, y].real.astype(np.float32))
            t[zrank * self.zl:(zrank + 1) * self.zl, y % self.yl, :, 1] = torch.from_numpy(f[:, y].imag.astype(np.float32))

    def fftz(self, slab):
        s = slab.numpy().astype(np.float64)
        f = np.fft.ifft(s[..., 0] + 1j * s[..., 1], axis=0)
        slab[..., 0] = torch.from_numpy(f.real.astype(np.float32))
        slab[..., 1] = torch.from_numpy(f.imag.astype(np.float32))
        return slab


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "hackathon-fft_b200", "python"), os.path.join(root, "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _volume(dims):
    return np.random.default_rng(7).standard_normal(tuple(dims) + (2,)).astype(np.float32)


PENCIL_CASES = [((8, 12, 6), (2, 2), False), ((4, 8, 10), (2, 2), True), ((8, 4, 6), (4, 1), False),
                ((3, 8, 8), (1, 4), False)]


def _pencil_worker(rank, world, port, out_dir):
    _setup(rank, world, port)
    from b200fft.pencil import PencilFFT3D
    for k, (dims, grid, inverse) in enumerate(PENCIL_CASES):     # one rendezvous for every case (spawning is the slow part)
        Z, Y, X = dims
        P0, P1 = grid
        p0, p1 = divmod(rank, P1)
        zl, yl = Z // P0, Y // P1
        full = _volume(dims)
        x_local = torch.from_numpy(full[p0 * zl:(p0 + 1) * zl, p1 * yl:(p1 + 1) * yl].copy())
        keep = x_local.clone()
        pen = PencilFFT3D(dims, grid, inverse=inverse, engine=NumpyPencilEngine(inverse))
        out = pen.forward(x_local).clone()
        out2 = pen.forward(x_local)          # plan reuse; the caller's input is not modified
        assert torch.equal(out, out2) and torch.equal(x_local, keep)
        np.save(os.path.join(out_dir, "out_%d_%d.npy" % (k, rank)), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_pencil_decomposition_world4_gloo(tmp_path):
    world = 4
    mp.spawn(_pencil_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for k, (dims, grid, inverse) in enumerate(PENCIL_CASES):
        Z, Y, X = dims
        P0, P1 = grid
        full = _volume(dims).astype(np.float64)
        c = full[..., 0] + 1j * full[..., 1]
        want = np.fft.ifftn(c) if inverse else np.fft.fftn(c)
        yl, xl = Y // P0, X // P1
        for r in range(world):
            p0, p1 = divmod(r, P1)
            got = np.load(os.path.join(str(tmp_path), "out_%d_%d.npy" % (k, r))).astype(np.float64)
            got = got[..., 0] + 1j * got[..., 1]
            assert got.shape == (Z, yl, xl)
            ref = want[:, p0 * yl:(p0 + 1) * yl, p1 * xl:(p1 + 1) * xl]    # all z, Y slab p0, X slab p1
            assert np.linalg.norm(got - ref) <= 2e-6 * np.linalg.norm(ref), (dims, grid, r)


def _restore_worker(rank, world, port, dims, out_dir):
    _setup(rank, world, port)
    from b200fft.slab import SlabFFT3D
    Z, Y, X = dims
    zl = Z // world
    full = _volume(dims)
    x_local = torch.from_numpy(full[rank * zl:(rank + 1) * zl].copy())
    fwd = SlabFFT3D(dims, exchange="nccl", engine=NumpySlabEngine(dims, world))
    inv = SlabFFT3D(dims, exchange="nccl", inverse=True, engine=NumpyInverseSlabEngine(dims, world))
    spec = fwd.forward(x_local, natural=True)               # Z-slab distributed spectrum [Z/G][Y][X]
    assert tuple(spec.shape) == (zl, Y, X, 2)
    assert torch.equal(spec, fwd.restore(fwd.forward(x_local)))
    back = inv.forward(spec, natural=True)                   # and the round trip returns the input
    np.save(os.path.join(out_dir, "spec_%d.npy" % rank), spec.numpy())
    np.save(os.path.join(out_dir, "back_%d.npy" % rank), back.numpy())
    with pytest.raises(Exception):
        fwd.restore(spec)                                    # wrong distribution: restore() takes forward()'s result
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dims,world", [((8, 12, 6), 2), ((4, 4, 10), 4)])
def test_slab_restore_natural_order_and_round_trip_gloo(tmp_path, dims, world):
    mp.spawn(_restore_worker, args=(world, _free_port(), dims, str(tmp_path)), nprocs=world, join=True)
    Z, Y, X = dims
    full = _volume(dims).astype(np.float64)
    want = np.fft.fftn(full[..., 0] + 1j * full[..., 1])
    zl = Z // world
    for r in range(world):
        spec = np.load(os.path.join(str(tmp_path), "spec_%d.npy" % r)).astype(np.float64)
        spec = spec[..., 0] + 1j * spec[..., 1]
        ref = want[r * zl:(r + 1) * zl]
        assert np.linalg.norm(spec - ref) <= 2e-6 * np.linalg.norm(ref)
        back = np.load(os.path.join(str(tmp_path), "back_%d.npy" % r)).astype(np.float64)
        src = full[r * zl:(r + 1) * zl]
        assert np.linalg.norm(back - src) <= 2e-6 * np.linalg.norm(src)


def test_pencil_rejects_bad_grids_and_cpu_product_path():
    import b200fft
    from b200fft.pencil import PencilFFT3D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        if not torch.cuda.is_available():
            with pytest.raises(b200fft.B200FFTError):   # no CPU fallback in the product engine
                PencilFFT3D((8, 8, 8), (1, 1))
        with pytest.raises(b200fft.B200FFTError):       # grid does not match the world size
            PencilFFT3D((8, 8, 8), (2, 2), engine=NumpyPencilEngine())
        pen = PencilFFT3D((4, 6, 8), (1, 1), engine=NumpyPencilEngine())
        x = torch.from_numpy(_volume((4, 6, 8)))
        got = pen.forward(x).numpy().astype(np.float64)
        want = np.fft.fftn(x.numpy()[..., 0].astype(np.float64) + 1j * x.numpy()[..., 1])
        assert np.linalg.norm((got[..., 0] + 1j * got[..., 1]) - want) <= 2e-6 * np.linalg.norm(want)
    finally:
        dist.destroy_process_group()
