"""Plane kernels (csrc/plane.cuh): the two innermost axes of a transform in one tile per (y, x) plane — complex forward /
inverse, real input, half-spectrum R2C — followed by strided passes over the outer axes. Default wherever the innermost
plane is 64 x 64 and the fused persistent kernel is not (2-D batches, real input, rank 4+); B200FFT_PLANE=1 prefers it over
the fused kernel too. Against numpy float64 at the stated fp32 tolerance, and against the per-axis plan."""
import numpy as np
import pytest

import b200fft

pytestmark = pytest.mark.gpu

CASES = [
    # shape, mode, inverse, set B200FFT_PLANE=1 explicitly
    ((5, 64, 64, 64), "c2c", False, False),
    ((5, 64, 64, 64), "c2c", True, True),
    ((7, 64, 64, 64), "half", False, False),
    ((3, 64, 64, 64), "real", False, False),
    ((9, 64, 64), "c2c", False, False),
    ((9, 64, 64), "c2c", True, False),
    ((4, 64, 64), "half", False, False),
    ((6, 64, 64), "real", False, False),
    ((2, 3, 64, 64, 64), "c2c", False, False),
    ((2, 10, 64, 64), "half", False, False),
    ((3, 100, 64, 64), "c2c", True, False),          # outer axis on a run-time specialised strided kernel
    # 128 x 128 planes: ONE shared buffer (139 KB, one CTA per SM), stages exchanged in place through registers
    ((3, 128, 128, 128), "c2c", False, False),
    ((2, 128, 128, 128), "c2c", True, False),
    ((2, 128, 128, 128), "real", False, False),
    ((5, 128, 128), "c2c", False, False),
    ((3, 128, 128, 128), "half", False, False),      # in-place R2C plane: the 65 bins of a row fit the 72-element pitch
    ((6, 128, 128), "half", False, False),
]


@pytest.mark.parametrize("shape,mode,inverse,force", CASES)
def test_plane_plans(shape, mode, inverse, force, monkeypatch):
    import torch
    if force:
        monkeypatch.setenv("B200FFT_PLANE", "1")
    rng = np.random.default_rng(19)
    comps = 2 if mode == "c2c" else 1
    x = rng.standard_normal(shape + (comps,)).astype(np.float32)
    oshape = shape[:-1] + (shape[-1] // 2 + 1, 2) if mode == "half" else shape + (2,)
    rm = b200fft.REAL_HALF if mode == "half" else b200fft.REAL_FULL
    plan = b200fft.plan_fft("float32", "float32", x.shape, oshape, inverse=inverse, real_mode=rm)
    desc = plan.describe()
    tag = "plane%dx%d" % (shape[-2], shape[-1])
    assert (("r2c" + tag) if mode == "half" else tag) in desc.split("\n")[0] and not desc.startswith("fused"), desc
    assert plan.launches == len(shape) - 2
    xt = torch.from_numpy(x).cuda()
    keep = xt.clone()
    out = torch.full(oshape, float("nan"), device="cuda")
    b200fft.fft(out, xt, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(xt, keep)
    got = out.cpu().numpy().astype(np.float64)
    got = got[..., 0] + 1j * got[..., 1]
    xd = x.astype(np.float64)
    xc = xd[..., 0] + (1j * xd[..., 1] if comps == 2 else 0)
    axes = tuple(range(1, len(shape)))
    want = np.fft.rfftn(xd[..., 0], axes=axes) if mode == "half" else (np.fft.ifftn(xc, axes=axes) if inverse else np.fft.fftn(xc, axes=axes))
    assert np.linalg.norm(got - want) <= 2e-6 * np.sqrt(len(axes)) * np.linalg.norm(want), desc
    plain = b200fft.plan_fft("float32", "float32", x.shape, oshape, inverse=inverse, real_mode=rm, flags=b200fft.FLAG_NO_FUSED)
    assert tag not in plain.describe()
    out2 = torch.empty_like(out)
    b200fft.fft(out2, xt, plan=plain)
    torch.cuda.synchronize()
    assert float((out - out2).norm() / out2.norm()) < 1e-6
    plan.destroy()
    plain.destroy()


def test_plane_respects_user_bases_and_knob(monkeypatch):
    p = b200fft.plan_fft("float32", "float32", (4, 64, 64, 2), (4, 64, 64, 2), bases=[[4], [4]])   # [4,4,4] cannot form 8 x 8:
    assert "plane64x64(8x8" not in p.describe() and "jitplane64x64(16x4;16x4)" in p.describe(), p.describe()   # specialised as 16 x 4
    p.destroy()
    monkeypatch.setenv("B200FFT_PLANE", "0")
    p = b200fft.plan_fft("float32", "float32", (4, 64, 64, 2), (4, 64, 64, 2))
    assert "plane64x64" not in p.describe()
    p.destroy()


JIT_PLANES = [
    # shape, mode, inverse: plane sizes without a registered variant -> in-place plane kernel specialised at plan time
    ((6, 100, 100, 100), "c2c", False),
    ((3, 100, 100, 100), "c2c", True),
    ((5, 48, 48, 48), "c2c", False),
    ((2, 96, 96, 96), "c2c", False),
    ((7, 50, 60), "c2c", False),
    ((4, 100, 100, 100), "half", False),
    ((9, 48, 96), "half", False),
    ((3, 32, 32, 32), "real", False),
    ((2, 3, 60, 100), "c2c", False),
]


@pytest.mark.parametrize("shape,mode,inverse", JIT_PLANES)
def test_plane_kernels_specialised_at_plan_time(shape, mode, inverse):
    import torch
    rng = np.random.default_rng(29)
    comps = 2 if mode == "c2c" else 1
    x = rng.standard_normal(shape + (comps,)).astype(np.float32)
    oshape = shape[:-1] + (shape[-1] // 2 + 1, 2) if mode == "half" else shape + (2,)
    rm = b200fft.REAL_HALF if mode == "half" else b200fft.REAL_FULL
    plan = b200fft.plan_fft("float32", "float32", x.shape, oshape, inverse=inverse, real_mode=rm)
    desc = plan.describe()
    assert ("jitr2cplane" if mode == "half" else "jitplane") in desc.split("\n")[0] and "NVRTC" in desc, desc
    assert plan.launches == len(shape) - 2
    xt = torch.from_numpy(x).cuda()
    out = torch.full(oshape, float("nan"), device="cuda")
    b200fft.fft(out, xt, plan=plan)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    got = got[..., 0] + 1j * got[..., 1]
    xd = x.astype(np.float64)
    xc = xd[..., 0] + (1j * xd[..., 1] if comps == 2 else 0)
    axes = tuple(range(1, len(shape)))
    want = np.fft.rfftn(xd[..., 0], axes=axes) if mode == "half" else (np.fft.ifftn(xc, axes=axes) if inverse else np.fft.fftn(xc, axes=axes))
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - want) <= 2e-6 * np.sqrt(len(axes)) * np.linalg.norm(want), desc
    plan.destroy()
