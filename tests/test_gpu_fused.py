"""Fused N-d kernel (one persistent kernel for all axes, intermediate kept in L2) against torch float64
FFTs and against the per-axis passes of the same library, through the C ABI."""
import os

import numpy as np
import pytest

import b200fft

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fused_before_plane(monkeypatch):
    """Since round 2 the two-pass plane plan (csrc/plane.cuh, tests/test_gpu_plane.py) is tried before the fused persistent
    kernel where both cover a problem (they tie on 100 x 64^3). These tests are about the fused kernels: switch the plane
    passes off, which is also what a user who wants the fused kernel everywhere does."""
    monkeypatch.setenv("B200FFT_PLANE", "0")


@pytest.fixture(autouse=True)
def _enable_fused(monkeypatch):
    monkeypatch.setenv("B200FFT_FUSED", "1")   # opt-in while the per-axis kernels are faster

REL_L2 = 2e-6          # ours vs float64, N(0,1) data (x sqrt(ndims) margin included)
MAX_ABS = 1e-5


def _ref(x, inverse=False, half=False):
    import torch
    if x.shape[-1] == 1:
        xd = x[..., 0].double()
        axes = tuple(range(1, xd.dim()))
        return torch.fft.rfftn(xd, dim=axes) if half else torch.fft.fftn(xd, dim=axes)
    xc = torch.view_as_complex(x.double().contiguous())
    axes = tuple(range(1, xc.dim()))
    return torch.fft.ifftn(xc, dim=axes) if inverse else torch.fft.fftn(xc, dim=axes)


def _check(out, want):
    import torch
    got = torch.view_as_complex(out.double().contiguous())
    rel = float((got - want).norm() / want.norm())
    mx = float((got - want).abs().max() / want.abs().max())
    assert rel < REL_L2 and mx < MAX_ABS, (rel, mx)


CASES = [
    # (shape incl. batch, inverse, mode, variant-name substring: "ndA" = v2 (producer warp + bulk-async
    #  staging, fused2.cuh), "nd<dims>" = v1 (fused.cuh))
    ((5, 64, 64, 64), False, "c2c", "t256_plane"),
    ((5, 64, 64, 64), True, "c2c", "t256_plane"),
    ((23, 64, 64, 64), False, "c2c", "t256_rows"),
    ((3, 128, 128, 128), False, "c2c", "t512_plane"),
    ((2, 128, 128, 128), True, "c2c", "nd128x128x128_inv_t256_rows"),
    ((1, 256, 256, 256), False, "c2c", "nd256"),
    ((2, 256, 256, 256), True, "c2c", "nd256"),
    ((1, 512, 512, 512), False, "c2c", "nd512"),
    ((7, 640, 480), False, "c2c", "nd640"),
    ((3, 640, 480), True, "c2c", "nd640"),
    ((5, 640, 480), False, "real", "nd640"),
    ((5, 640, 480), False, "half", "nd640"),
    ((1, 640, 480), False, "c2c", "nd640"),
    ((5, 64, 64, 64), False, "c2c", "+32_plane"),
    ((37, 64, 64, 64), True, "c2c", "+32_plane"),
    ((23, 64, 64, 64), False, "c2c", "+32_rows"),
    ((3, 128, 128, 128), False, "c2c", "ndA128"),
    ((2, 128, 128, 128), True, "c2c", "ndA128"),
    ((1, 256, 256, 256), False, "c2c", "ndA256"),
    ((2, 256, 256, 256), True, "c2c", "ndA256"),
    ((1, 512, 512, 512), False, "c2c", "ndA512"),
    ((7, 640, 480), False, "c2c", "ndA640"),
    ((3, 640, 480), True, "c2c", "ndA640"),
    ((5, 640, 480), False, "real", "ndA640"),
    ((1, 640, 480), False, "c2c", "ndA640"),
    # half-spectrum 3-D: (y, x) planes with unpack + y pass in shared memory (AR2CPlane), then the strided z phase
    ((5, 64, 64, 64), False, "half", "t288+32_r2cplane64x64(8x8;2;8x4)_z32:"),
    ((37, 64, 64, 64), False, "half", "t288+32_r2cplane64x64(8x8;2;8x4)_z32:"),
    ((3, 64, 64, 64), False, "half", "r2cplane64x64(8x8;2;8x4)_z64"),
    ((7, 64, 64, 64), False, "half", "z32_t160_zslot"),
    ((41, 64, 64, 64), False, "half", "z32_t128_zslot"),
    ((5, 128, 128, 128), False, "half", "r2cplane128x128(16x8;2;8x8)_z32_t384"),
    ((9, 64, 64, 64), False, "c2c", "z32_t128_r1"),
    ((9, 64, 64, 64), True, "c2c", "z64_t256_r1"),
    ((30, 64, 64, 64), False, "c2c", "z32_t192_r1"),
]


@pytest.mark.parametrize("shape,inverse,mode,prefer", CASES)
def test_fused_matches_float64_and_per_axis_passes(shape, inverse, mode, prefer, monkeypatch):
    import torch
    monkeypatch.setenv("B200FFT_FUSED_PREFER", prefer)
    comps = 2 if mode == "c2c" else 1
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(tuple(shape) + (comps,), generator=g, device="cuda")
    oshape = tuple(shape[:-1]) + (shape[-1] // 2 + 1, 2) if mode == "half" else tuple(shape) + (2,)
    rm = b200fft.REAL_HALF if mode == "half" else b200fft.REAL_FULL
    plan = b200fft.plan_fft("float32", "float32", x.shape, oshape, inverse=inverse, real_mode=rm)
    assert plan.describe().startswith("fused"), plan.describe()
    assert prefer in plan.describe(), plan.describe()
    assert plan.launches == 1
    out = torch.full(oshape, float("nan"), device="cuda")
    keep = x.clone()
    b200fft.fft(out, x, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(x, keep)
    want = _ref(x, inverse, mode == "half")
    _check(out, want)
    # a second and third launch reuse the self-clearing counters
    out2 = torch.full(oshape, float("nan"), device="cuda")
    b200fft.fft(out2, x, plan=plan)
    b200fft.fft(out2, x, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    # the per-axis passes of the same library give the same result up to rounding order
    plain = b200fft.plan_fft("float32", "float32", x.shape, oshape, inverse=inverse, real_mode=rm,
                             flags=b200fft.FLAG_NO_FUSED)
    assert not plain.describe().startswith("fused")
    out3 = torch.empty(oshape, device="cuda")
    b200fft.fft(out3, x, plan=plain)
    torch.cuda.synchronize()
    assert float((out3 - out).norm() / out.norm()) < 1e-6
    plan.destroy()
    plain.destroy()


def test_fused_in_place():
    import torch
    x = torch.randn((6, 64, 64, 64, 2), device="cuda")
    want = _ref(x)
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    assert plan.describe().startswith("fused")
    b200fft.fft(x, x, plan=plan)
    torch.cuda.synchronize()
    _check(x, want)


def test_fused_respects_user_bases_and_opt_out(monkeypatch):
    import torch
    x = torch.randn((2, 64, 64, 64, 2), device="cuda")
    # [4] x3 can be grouped into the variant's (8)(8)? 4*4*4 = 64 -> groups of product 8 do not exist
    p = b200fft.plan_fft("float32", "float32", x.shape, x.shape, bases=[[4], [2], [2]])
    assert not p.describe().startswith("fused")
    out = torch.empty_like(x)
    b200fft.fft(out, x, plan=p)
    torch.cuda.synchronize()
    _check(out, _ref(x))
    monkeypatch.setenv("B200FFT_FUSED", "0")
    p2 = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    assert not p2.describe().startswith("fused")


def test_fused_through_exec_host():
    """exec_host runs the fused kernel on batch chunks (a schedule per chunk size)."""
    import torch
    x = torch.randn((20, 64, 64, 64, 2))
    h_in = x.pin_memory()
    h_out = torch.empty_like(x).pin_memory()
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    plan.exec_host(h_out.numpy(), h_in.numpy())
    want = _ref(x.cuda())
    _check(h_out.cuda(), want)


def test_fused_exec_is_graph_capturable():
    """b200fft_exec of a fused plan neither allocates nor synchronises (the schedule is built at plan creation, the
    launch is cooperative, no event chain): it can be captured into a CUDA graph and replayed."""
    import torch
    x = torch.randn((8, 64, 64, 64, 2), device="cuda")
    out = torch.full_like(x, float("nan"))
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    assert plan.describe().startswith("fused ndA")
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            plan.exec(out, x, st.cuda_stream)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    _check(out, _ref(x))


def test_refused_cooperative_launch_falls_back_to_per_axis_passes(monkeypatch):
    """a statically scheduled grid that cannot be co-resident must not be launched. The launch is cooperative;
    when the driver refuses it (simulated here) the plan's non-persistent passes run instead and the result is the same."""
    import torch
    x = torch.randn((4, 64, 64, 64, 2), device="cuda")
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    assert plan.describe().startswith("fused ndA")
    good = torch.empty_like(x)
    b200fft.fft(good, x, plan=plan)
    torch.cuda.synchronize()
    monkeypatch.setenv("B200FFT_TEST_REFUSE_COOP", "1")
    before = b200fft.launch_count()
    out = torch.full_like(x, float("nan"))
    b200fft.fft(out, x, plan=plan)
    torch.cuda.synchronize()
    assert b200fft.launch_count() - before == 3          # three per-axis kernels (plane passes are off here), not one fused launch
    _check(out, _ref(x))
    assert float((out - good).norm() / good.norm()) < 1e-6
    monkeypatch.delenv("B200FFT_TEST_REFUSE_COOP")
    b200fft.fft(out, x, plan=plan)                       # the plan stays on the fallback once refused
    torch.cuda.synchronize()
    _check(out, _ref(x))


def test_two_fused_plans_on_two_streams():
    """Statically scheduled persistent kernels are launched cooperatively: each launch waits until its whole grid can be
    resident, so two plans launched back to back on different streams cannot starve each other of SMs."""
    import torch
    xs = [torch.randn((40, 64, 64, 64, 2), device="cuda") for _ in range(2)]
    outs = [torch.empty_like(x) for x in xs]
    plans = [b200fft.plan_fft("float32", "float32", x.shape, x.shape) for x in xs]
    assert all(p.describe().startswith("fused ndA") for p in plans)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for _ in range(20):
        for p, x, o, st in zip(plans, xs, outs, streams):
            p.exec(o, x, st.cuda_stream)
    torch.cuda.synchronize()
    for x, o in zip(xs, outs):
        _check(o, _ref(x))
