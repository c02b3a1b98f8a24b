"""Slab-decomposed single 3-D C2C transform over the ranks of a torch.distributed group.

One process per GPU. Rank g owns z planes [g*Z/G, (g+1)*Z/G) of a (Z, Y, X) complex64 volume:

  1. local 2-D transform over (Y, X) of its planes          (rows X, then strided axis Y)
  2. exchange: block (z in slab g, y in slab h, all x) goes g -> h
  3. transform along Z on the received [Z][Y/G][X] slab      (strided axis, inner = Y/G * X)

The result stays distributed by Y ("transposed out"): rank h returns out[z, y_local, x] for
y = h*Y/G + y_local. Two exchange modes share the same kernels:

  * "p2p"  — the Y pass of step 1 stores every output row directly into the owning peer's slab
             through CUDA-IPC-mapped pointers (b200fft_exec_scatter): the all-to-all is fused
             into the compute kernel's stores and overlaps it tile by tile; a one-element
             all-reduce is the only collective (barrier before step 3).
  * "nccl" — the same scattering store targets a local pack buffer [G][Z/G][Y/G][X], followed
             by one all_to_all_single (the baseline; also what the gloo CPU tests exercise).
  * "fused" — ONE persistent kernel per rank (b200fft_slab_exec, csrc/slab.cuh) runs all three steps:
             the Y tiles store into the peers' slabs and bump per-x-block arrival counters on every rank
             over NVLink, the Z tiles of an x-block start as soon as its counter is complete, so the
             exchange overlaps the butterflies on both sides and there is no barrier at all. Cubic
             volumes (64/128/256/512) only.

`restore(out)` (or `forward(x, natural=True)`) is the optional step AFTER the path: a second
all-to-all that brings the Y-slab result back to the input's Z-slab distribution [Z/G][Y][X]
(natural order, what a caller chaining forward -> inverse needs). It is not part of the timed
transform; it uses one all_to_all_single and one strided copy in every exchange mode.

SlabFFT3D owns the decomposition and exchange logic; an `engine` supplies the three local
operations (alloc, fft2_scatter, fftz). The product engine is CUDA-only (no CPU fallback);
tests/test_slab_gloo.py injects a numpy engine to run the same orchestration over gloo.
"""
import torch
import torch.distributed as dist

import b200fft


class _DevBuf:
    """A library-allocated device buffer (base pointer exportable over CUDA IPC) viewed as a tensor."""

    def __init__(self, shape, device_index):
        self.shape = tuple(shape)
        n = 1
        for v in self.shape:
            n *= v
        self.ptr = b200fft.device_malloc(n * 4)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False),
                                         "version": 2, "strides": None}
        self.tensor = torch.as_tensor(self, device=torch.device("cuda", device_index))

    def free(self):
        if self.ptr:
            self.tensor = None
            b200fft.device_free(self.ptr)
            self.ptr = 0


class CudaSlabEngine:
    """Local device operations of the slab transform, through the C ABI."""

    def __init__(self, dims, world, inverse=False):
        Z, Y, X = dims
        self.zl, self.yl = Z // world, Y // world
        self.plan2d = b200fft.plan_fft("float32", "float32", (self.zl, Y, X, 2), (self.zl, Y, X, 2), inverse=inverse,
                                       flags=b200fft.FLAG_NO_FUSED)  # exec_scatter drives the per-axis passes
        self.planz = b200fft.plan_fft("float32", "float32", (1, Z, self.yl, X, 2), (1, Z, self.yl, X, 2),
                                      axis_mask=1, inverse=inverse)
        self.work = torch.empty((self.zl, Y, X, 2), device="cuda", dtype=torch.float32)
        self.shared = []

    def alloc(self, shape, shared=False):
        if shared:
            buf = _DevBuf(shape, torch.cuda.current_device())
            self.shared.append(buf)
            return buf.tensor
        return torch.empty(shape, device="cuda", dtype=torch.float32)

    def fft2_scatter(self, x_local, targets, zrank):
        """2-D transform of the local planes; output row y of plane z lands in
        targets[y // yl][(zrank*zl + z), y % yl, :] (targets: tensors or raw device pointers)."""
        st = torch.cuda.current_stream().cuda_stream
        self.plan2d.exec_scatter(targets, zrank, x_local, self.work, st)

    def fftz(self, slab):
        v = slab.unsqueeze(0)  # the plan's layout carries the batch dimension
        b200fft.fft(v, v, plan=self.planz)
        return slab

    def close(self):
        for b in self.shared:
            b.free()
        self.shared = []
        self.plan2d.destroy()
        self.planz.destroy()


class SlabFFT3D:
    """forward(x_local[Z/G, Y, X, 2]) -> out_local[Z, Y/G, X, 2] (Y-slab distributed)."""

    def __init__(self, dims, group=None, exchange="p2p", inverse=False, engine=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dims = tuple(int(v) for v in dims)
        Z, Y, X = self.dims
        if exchange not in ("p2p", "nccl", "fused"):
            raise ValueError("exchange must be 'p2p', 'nccl' or 'fused'")
        if Z % self.world or Y % self.world:
            raise b200fft.B200FFTError(1, "slab decomposition needs Z and Y divisible by the number of ranks")
        self.exchange = exchange
        self.zl, self.yl = Z // self.world, Y // self.world
        if engine is None:
            if not torch.cuda.is_available():
                raise b200fft.B200FFTError(5, "SlabFFT3D needs CUDA devices: there is no CPU path")
            if exchange == "fused":
                self._init_fused(inverse)
                return
            engine = CudaSlabEngine(self.dims, self.world, inverse)
        self.engine = engine
        self.calls = 0
        self._opened = []
        if exchange == "p2p":
            # Two receive slabs, alternated per call: a peer may already be scattering call k+1 while
            # this rank still reads call k's slab; by the barrier of call k+1 every rank has finished
            # its Z pass of call k, so slab k%2 is free again at call k+2.
            self.recv = [engine.alloc((Z, self.yl, X, 2), shared=True) for _ in range(2)]
            mine = [b200fft.ipc_export(t.data_ptr()) for t in self.recv]
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            self.peer_targets = []
            for k in range(2):
                row = []
                for r in range(self.world):
                    if r == self.rank:
                        row.append(self.recv[k].data_ptr())
                    else:
                        p = b200fft.ipc_open(handles[r][k])
                        self._opened.append(p)
                        row.append(p)
                self.peer_targets.append(row)
            self.flag = engine.alloc((1,))
            self.flag.zero_()
        else:
            self.pack = engine.alloc((self.world, self.zl, self.yl, X, 2))
            self.recv = [engine.alloc((Z, self.yl, X, 2))]

    def _init_fused(self, inverse):
        Z, Y, X = self.dims
        if not (Z == Y == X):
            raise b200fft.B200FFTError(4, "the fused slab kernel covers cubic volumes only")
        self.engine = None
        self.calls = 0
        self._opened = []
        self.plan = b200fft.SlabPlan(Z, self.world, self.rank, inverse)
        nfloat = self.plan.recv_bytes // 4
        dev = torch.cuda.current_device()
        self._bufs = [_DevBuf((nfloat,), dev) for _ in range(2)]
        for b in self._bufs:
            b.tensor.zero_()                      # arrival counters start at 0 (they only ever grow)
        torch.cuda.synchronize()
        self.recv = [b.tensor[:Z * self.yl * X * 2].view(Z, self.yl, X, 2) for b in self._bufs]
        mine = [b200fft.ipc_export(b.ptr) for b in self._bufs]
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=self.group)   # also orders "zeroed" before any peer's first store
        self.peer_targets = []
        for k in range(2):
            row = []
            for r in range(self.world):
                if r == self.rank:
                    row.append(self._bufs[k].ptr)
                else:
                    p = b200fft.ipc_open(handles[r][k])
                    self._opened.append(p)
                    row.append(p)
            self.peer_targets.append(row)
        self.work = torch.empty((self.zl, Y, X, 2), device="cuda", dtype=torch.float32)
        dist.barrier(group=self.group)

    def timeouts(self):
        """fused mode: Z tiles that gave up waiting for a peer (0 unless a rank died or skipped a call). Synchronises."""
        if self.exchange != "fused":
            return 0
        torch.cuda.synchronize()
        w = self.plan.timeout_offset // 4
        return int(sum(int(b.tensor[w:w + 1].view(torch.int32).item()) for b in self._bufs))

    def restore(self, out_local):
        """Second exchange: the Y-slab result [Z][Y/G][X] of forward() back to the Z-slab distribution
        [Z/G][Y][X] of the input. Block g of out_local (z in slab g) goes to rank g; rank g
        receives (its z planes, y rows of slab h) from every h and interleaves them along y."""
        Z, Y, X = self.dims
        if tuple(out_local.shape) != (Z, self.yl, X, 2) or not out_local.is_contiguous():
            raise b200fft.B200FFTError(2, "restore() takes forward()'s dense [Z][Y/G][X][2] result")
        got = torch.empty((self.world, self.zl, self.yl, X, 2), dtype=out_local.dtype, device=out_local.device)
        dist.all_to_all_single(got.view(-1), out_local.view(-1), group=self.group)
        nat = torch.empty((self.zl, Y, X, 2), dtype=out_local.dtype, device=out_local.device)
        nat.view(self.zl, self.world, self.yl, X, 2).copy_(got.permute(1, 0, 2, 3, 4))
        return nat

    def forward(self, x_local, natural=False):
        out = self._forward(x_local)
        return self.restore(out) if natural else out

    def _forward(self, x_local):
        if self.exchange == "fused":
            k = self.calls & 1
            self.calls += 1
            self.plan.exec(x_local, self.work, self.peer_targets[k], k, torch.cuda.current_stream().cuda_stream)
            return self.recv[k]
        if self.exchange == "p2p":
            k = self.calls & 1
            self.calls += 1
            self.engine.fft2_scatter(x_local, self.peer_targets[k], self.rank)
            dist.all_reduce(self.flag, group=self.group)  # barrier: all stores have landed
            slab = self.recv[k]
        else:
            # block h of the pack buffer = (my z planes, y rows owned by rank h); after the
            # all-to-all, block g of the receive slab = rank g's z planes: [Z][Y/G][X] in z order
            self.engine.fft2_scatter(x_local, [self.pack[h] for h in range(self.world)], 0)
            slab = self.recv[0]
            dist.all_to_all_single(slab.view(-1), self.pack.view(-1), group=self.group)
        return self.engine.fftz(slab)

    def close(self):
        for p in self._opened:
            b200fft.ipc_close(p)
        self._opened = []
        if self.exchange == "fused" and getattr(self, "plan", None) is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)        # no peer may still be storing into buffers we free
            self.recv = []
            for b in self._bufs:
                b.free()
            self.plan.destroy()
            self.plan = None
            return
        if hasattr(self.engine, "close"):
            self.engine.close()
