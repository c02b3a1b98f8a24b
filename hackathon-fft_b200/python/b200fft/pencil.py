"""Pencil-decomposed single 3-D C2C transform over a P0 x P1 process grid (the
decomposition for more ranks than one axis has planes, i.e. beyond one NVSwitch node of 8 GPUs).

One process per GPU, rank = p0 * P1 + p1. A (Z, Y, X) complex64 volume is split over two axes at a time:

  input    rank (p0, p1) owns x[z in slab p0 of P0][y in slab p1 of P1][all X]          "X pencils"
  1. transform along X on the local rows                                              (contiguous axis)
  2. exchange inside the ROW group (fixed p0, the P1 ranks): y <-> x                   -> [Z/P0][Y][X/P1]
  3. transform along Y                                                                (strided axis, inner X/P1)
  4. exchange inside the COLUMN group (fixed p1, the P0 ranks): z <-> y                -> [Z][Y/P0][X/P1]
  5. transform along Z                                                                (strided axis)
  output   rank (p0, p1) owns out[all Z][y in slab p0 of P0][x in slab p1 of P1]        "Z pencils"

Each exchange is one all_to_all_single among P1 (resp. P0) ranks, so a rank talks to P0 + P1 - 2 peers
instead of P0 * P1 - 1: that is the point of the decomposition when the job spans more than one node.
With P1 == 1 steps 1-3 collapse to the slab decomposition's local 2-D transform (b200fft.slab.SlabFFT3D
is the faster path on one node: it fuses the exchange into the Y pass's stores).

PencilFFT3D owns the grid, the sub-groups and the block order; an `engine` supplies the local
operations (alloc, fft_axis). The product engine is CUDA-only (plans over the C ABI, no CPU fallback);
tests/test_pencil_gloo.py injects a numpy engine to run the same orchestration over gloo with 4 ranks.
The pack / unpack steps around the exchanges are strided tensor copies (they are the "transposes" the
single-GPU path does not need; fusing them into the passes' stores as exec_scatter does for slabs is the
next step for this row).
"""
import torch
import torch.distributed as dist

import b200fft


class CudaPencilEngine:
    """fft_axis(t, axis): in-place transform of one axis of a dense local [a][b][c][2] fp32 array."""

    def __init__(self, inverse=False):
        self.inverse = inverse
        self.plans = {}

    def alloc(self, shape):
        return torch.empty(shape, device="cuda", dtype=torch.float32)

    def fft_axis(self, t, axis):
        a, b, c, _ = t.shape
        key = (a, b, c, axis)
        if key not in self.plans:
            # the plan layout carries the batch dimension; axes other than `axis` are batch / inner strides
            if axis == 2:
                layout, mask = (a * b, c, 2), 0
            elif axis == 1:
                layout, mask = (a, b, c, 2), 1
            else:
                layout, mask = (1, a, b * c, 2), 1
            plan = b200fft.plan_fft("float32", "float32", layout, layout, inverse=self.inverse, axis_mask=mask)
            self.plans[key] = (plan, layout)
        plan, layout = self.plans[key]
        v = t.view(layout)
        b200fft.fft(v, v, plan=plan)
        return t

    def close(self):
        for plan, _ in self.plans.values():
            plan.destroy()
        self.plans = {}


class PencilFFT3D:
    """forward(x_local[Z/P0, Y/P1, X, 2]) -> out_local[Z, Y/P0, X/P1, 2]."""

    def __init__(self, dims, grid, group=None, inverse=False, engine=None):
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dims = tuple(int(v) for v in dims)
        self.P0, self.P1 = int(grid[0]), int(grid[1])
        Z, Y, X = self.dims
        if self.P0 < 1 or self.P1 < 1 or self.P0 * self.P1 != self.world:
            raise b200fft.B200FFTError(1, "grid %dx%d does not match %d ranks" % (self.P0, self.P1, self.world))
        if Z % self.P0 or Y % self.P1 or Y % self.P0 or X % self.P1:
            raise b200fft.B200FFTError(1, "pencil decomposition needs Z and Y divisible by P0, Y and X divisible by P1")
        if engine is None:
            if not torch.cuda.is_available():
                raise b200fft.B200FFTError(5, "PencilFFT3D needs CUDA devices: there is no CPU path")
            engine = CudaPencilEngine(inverse)
        self.engine = engine
        self.p0, self.p1 = divmod(self.rank, self.P1)
        self.zl, self.yl1, self.yl0, self.xl = Z // self.P0, Y // self.P1, Y // self.P0, X // self.P1
        # every rank must create every sub-group, in the same order (torch.distributed rule)
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(self.world))
        self.row_group = self.col_group = None
        for a in range(self.P0):
            g = dist.new_group([ranks[a * self.P1 + b] for b in range(self.P1)])
            if a == self.p0:
                self.row_group = g
        for b in range(self.P1):
            g = dist.new_group([ranks[a * self.P1 + b] for a in range(self.P0)])
            if b == self.p1:
                self.col_group = g
        self.buf_x = engine.alloc((self.zl, self.yl1, X, 2))                 # step 1 (keeps the caller's input intact)
        self.pack1 = engine.alloc((self.P1, self.zl, self.yl1, self.xl, 2))
        self.recv1 = engine.alloc((self.P1, self.zl, self.yl1, self.xl, 2))
        self.buf_y = engine.alloc((self.zl, Y, self.xl, 2))                  # step 3
        self.pack2 = engine.alloc((self.P0, self.zl, self.yl0, self.xl, 2))
        self.buf_z = engine.alloc((Z, self.yl0, self.xl, 2))                 # step 5 = the result

    def forward(self, x_local):
        Z, Y, X = self.dims
        if tuple(x_local.shape) != (self.zl, self.yl1, X, 2) or not x_local.is_contiguous():
            raise b200fft.B200FFTError(2, "forward() takes a dense [Z/P0][Y/P1][X][2] block")
        e = self.engine
        # 1. X rows
        self.buf_x.copy_(x_local)
        e.fft_axis(self.buf_x, 2)
        # 2. row-group exchange: block q = my (z, y) rows, x columns of slab q; received block h = rank (p0, h)'s
        #    y rows of my x columns -> y = h * yl1 + y_local
        self.pack1.copy_(self.buf_x.view(self.zl, self.yl1, self.P1, self.xl, 2).permute(2, 0, 1, 3, 4))
        dist.all_to_all_single(self.recv1.view(-1), self.pack1.view(-1), group=self.row_group)
        self.buf_y.view(self.zl, self.P1, self.yl1, self.xl, 2).copy_(self.recv1.permute(1, 0, 2, 3, 4))
        # 3. Y (strided by X/P1)
        e.fft_axis(self.buf_y, 1)
        # 4. column-group exchange: block q = my z planes, y rows of slab q (of P0); received block h = rank
        #    (h, p1)'s z planes -> z = h * zl + z_local, which is the receive buffer's own order
        self.pack2.copy_(self.buf_y.view(self.zl, self.P0, self.yl0, self.xl, 2).permute(1, 0, 2, 3, 4))
        dist.all_to_all_single(self.buf_z.view(-1), self.pack2.view(-1), group=self.col_group)
        # 5. Z (strided by Y/P0 * X/P1)
        return e.fft_axis(self.buf_z, 0)

    def close(self):
        if hasattr(self.engine, "close"):
            self.engine.close()
