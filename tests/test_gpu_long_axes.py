"""Axes too long for one tile run as two passes ("four-step": N = N1*N2, twiddle fused into pass A's store,
natural-order store in pass B; csrc/split_registry.cu), plus the row variants for 1080 / 2160 / 4320: the
remaining shapes the reference publishes (fft/bench.mojo:107-122), against torch float64."""
import numpy as np
import pytest

import b200fft

pytestmark = pytest.mark.gpu


def _run(shape, inverse=False, bases=None):
    import torch
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(tuple(shape) + (2,), generator=g, device="cuda")
    out = torch.full_like(x, float("nan"))
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape, inverse=inverse, bases=bases)
    keep = x.clone()
    b200fft.fft(out, x, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(x, keep)
    xc = torch.view_as_complex(x.double().contiguous())
    axes = tuple(range(1, xc.dim()))
    want = torch.fft.ifftn(xc, dim=axes) if inverse else torch.fft.fftn(xc, dim=axes)
    got = torch.view_as_complex(out.double().contiguous())
    rel = float((got - want).norm() / want.norm())
    mx = float((got - want).abs().max() / want.abs().max())
    desc = plan.describe()
    plan.destroy()
    return rel, mx, desc


@pytest.mark.parametrize("shape,inverse", [((3, 16384), False), ((2, 16384), True), ((100, 16384), False)])
def test_contiguous_long_axis(shape, inverse, monkeypatch):
    monkeypatch.setenv("B200FFT_ROWS_INPLACE", "0")
    rel, mx, desc = _run(shape, inverse)
    assert "split n=16384 = 128 x 128" in desc and "splitB_rows128" in desc, desc
    assert rel < 2e-6 and mx < 1e-5, (rel, mx)


@pytest.mark.parametrize("shape,inverse", [((3, 16384), False), ((2, 16384), True), ((100, 16384), False), ((149, 16384), True)])
def test_long_row_in_one_launch(shape, inverse):
    """16384 points do not fit two exchange buffers; rows_ip_kernel (csrc/fast.cuh) keeps ONE and exchanges the middle stage
    in place through registers, so the published (100, 16384) shape is one launch instead of two split passes."""
    rel, mx, desc = _run(shape, inverse)
    assert desc.strip().count("\n") == 0 and "rowsIP16384_32x32x16_c1_t512" in desc, desc
    assert rel < 2e-6 and mx < 1e-5, (rel, mx)
    rel2, _, desc2 = _run(shape, inverse, bases=[[4]])   # user bases that regroup into (32, 32, 16)? 4^7: no -> other tiers
    assert rel2 < 2e-6, desc2


@pytest.mark.parametrize("shape,inverse", [((2, 1920, 1080), False), ((1, 1920, 1080), True),
                                           ((1, 3840, 2160), False), ((1, 7680, 4320), False)])
def test_published_2d_shapes(shape, inverse):
    rel, mx, desc = _run(shape, inverse)
    lines = desc.strip().split("\n")
    assert lines[0].startswith(("axis 1: rows%d_" % shape[2], "axis 1: rowsIP%d_" % shape[2])), desc  # contiguous axis: one row kernel
    assert "split n=%d = " % shape[1] in lines[1], desc                     # strided axis: two passes
    assert "generic" not in desc
    assert rel < 2e-6 and mx < 1e-5, (rel, mx)


def test_published_4d_5d_shapes():
    rel, mx, desc = _run((1, 64, 64, 64, 64))
    assert "generic" not in desc and rel < 2e-6 and mx < 1e-5, (rel, mx, desc)
    rel, mx, desc = _run((1, 25, 160, 160, 48))
    assert "generic" not in desc and rel < 2e-6 and mx < 1e-5, (rel, mx, desc)


def test_split_respects_user_bases():
    """Bases that cannot be grouped into any (N1, N2) kernel pair fall back to the runtime-length / generic kernels."""
    rel, mx, desc = _run((2, 1920, 64), bases=[[64, 30], [64]])      # 64*30: no such pair registered
    assert "split" not in desc.split("\n")[1] and rel < 2e-6, desc
    rel, mx, desc = _run((2, 1920, 64), bases=[[16, 8, 15], [8, 8]])  # groupable into (16,8) x (15)
    assert "split n=1920 = 128 x 15" in desc and rel < 2e-6, desc


@pytest.mark.parametrize("n", [2048, 2160, 4096, 4320, 8192, 16384])
@pytest.mark.parametrize("inverse", [False, True])
def test_registered_one_buffer_rows(n, inverse):
    """The registered one-buffer variants (fast_reg_rows_*.cu: reg_rows_inplace) serve complex and real input (full
    spectrum out) in both directions; R2C / C2R of these lengths stay on the two-buffer variants."""
    rel, mx, desc = _run((301, n), inverse)
    assert "rowsIP%d_" % n in desc and desc.strip().count("\n") == 0, desc
    assert rel < 2e-6 and mx < 1e-5, (rel, mx)
    import torch
    x = torch.randn(5, n, 1, device="cuda")
    out = torch.empty(5, n, 2, device="cuda")
    plan = b200fft.plan_fft("float32", "float32", x.shape, out.shape)
    assert "rowsIP%d_" % n in plan.describe() and "real-in" in plan.describe(), plan.describe()
    b200fft.fft(out, x, plan=plan)
    torch.cuda.synchronize()
    plan.destroy()
    want = torch.fft.fft(x[..., 0].double())
    got = torch.view_as_complex(out.double().contiguous())
    assert float((got - want).norm() / want.norm()) < 2e-6


@pytest.mark.parametrize("shape,inverse,dtype", [((3, 20000), False, "float32"), ((5, 10000), True, "float32"), ((150, 5000), False, "float32"),
                                                 ((3, 8640), True, "float32"), ((2, 15625), False, "float32"), ((3, 10000), False, "float64"),
                                                 ((4, 4096), True, "float64"), ((2, 12288), False, "float32")])
def test_long_rows_specialised_in_place(shape, inverse, dtype):
    """Contiguous rows of 4096 .. ~25000 points without a registered variant: rows_ip_kernel specialised at plan time
    (csrc/jit.cu, JIT_ROWS_IP) — any stage count from 3 to 5, the stage order chosen so the in-place stages fit the
    register file. One launch; B200FFT_ROWS_INPLACE=0 gives the two-buffer / two-pass plan back."""
    import torch
    rng = np.random.default_rng(17)
    x = rng.standard_normal(shape + (2,)).astype(dtype)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.full_like(d_in, float("nan"))
    plan = b200fft.plan_fft(dtype, dtype, d_in.shape, d_in.shape, inverse=inverse)
    desc = plan.describe()
    assert "jitrowsIP%d_" % shape[1] in desc and desc.strip().count("\n") == 0 and "local=" not in desc, desc
    b200fft.fft(d_out, d_in, plan=plan)
    torch.cuda.synchronize()
    plan.destroy()
    xc = x[..., 0].astype(np.float64) + 1j * x[..., 1]
    want = np.fft.ifft(xc, axis=1) if inverse else np.fft.fft(xc, axis=1)
    got = d_out.cpu().numpy().astype(np.float64)
    rel = np.linalg.norm((got[..., 0] + 1j * got[..., 1]) - want) / np.linalg.norm(want)
    assert rel < (1e-14 if dtype == "float64" else 2.5e-6), (rel, desc)


@pytest.mark.parametrize("n,dtype", [(3000, "float32"), (6000, "uint8"), (10000, "float64")])
def test_long_rows_real_input_full_spectrum(n, dtype):
    """Real input, full spectrum out (the reference's default mode): the one-buffer kernel's REAL instantiation reads the
    scalar array and casts on load."""
    import torch
    rng = np.random.default_rng(23)
    x = rng.integers(0, 256, size=(301, n, 1)).astype(dtype) if dtype == "uint8" else rng.standard_normal((301, n, 1)).astype(dtype)
    odt = "float64" if dtype == "float64" else "float32"
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.full((301, n, 2), float("nan"), dtype=getattr(torch, odt), device="cuda")
    plan = b200fft.plan_fft(dtype, odt, d_in.shape, d_out.shape)
    desc = plan.describe()
    assert "jitrowsIP%d_" % n in desc and "real-in" in desc, desc
    b200fft.fft(d_out, d_in, plan=plan)
    torch.cuda.synchronize()
    plan.destroy()
    want = np.fft.fft(x[..., 0].astype(np.float64), axis=1)
    got = d_out.cpu().numpy().astype(np.float64)
    rel = np.linalg.norm((got[..., 0] + 1j * got[..., 1]) - want) / np.linalg.norm(want)
    assert rel < (1e-14 if odt == "float64" else 2.5e-6), (rel, desc)


def test_long_rows_cast_on_load():
    """uint8 complex input straight into the in-place row kernel (the reference's own input type, fft/tests/fft.mojo:96-104)."""
    import torch
    rng = np.random.default_rng(19)
    x = rng.integers(0, 256, size=(6, 6000, 2), dtype=np.uint8)
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty((6, 6000, 2), dtype=torch.float32, device="cuda")
    plan = b200fft.plan_fft("uint8", "float32", d_in.shape, d_out.shape)
    assert "jitrowsIP6000_" in plan.describe() and "_inu8" in plan.describe(), plan.describe()
    b200fft.fft(d_out, d_in, plan=plan)
    torch.cuda.synchronize()
    plan.destroy()
    want = np.fft.fft(x[..., 0].astype(np.float64) + 1j * x[..., 1], axis=1)
    got = d_out.cpu().numpy().astype(np.float64)
    assert np.linalg.norm((got[..., 0] + 1j * got[..., 1]) - want) / np.linalg.norm(want) < 2e-6


@pytest.mark.parametrize("shape,inverse", [((3, 30000), False), ((2, 100000), True), ((1, 1 << 18), False), ((1, 1 << 20), False),
                                           ((2, 20000, 6), False), ((2, 3, 30000), True), ((1, 5 ** 7), False)])
def test_long_axes_without_a_registered_split(shape, inverse):
    """Any long axis whose stage list splits into two tile-sized halves: both passes specialised at plan time
    (csrc/jit.cu, make_jit_split_pass). These lengths used to return B200FFT_ERR_UNSUPPORTED."""
    rel, mx, desc = _run(shape, inverse)
    n = max(shape[1:])
    assert "split n=%d = " % n in desc and "jitsplitA" in desc and "generic" not in desc, desc
    assert rel < 2.5e-6 and mx < 1e-5, (rel, mx, desc)


@pytest.mark.parametrize("n", [10000, 40000])
def test_long_axis_fp64(n):
    import torch
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, n, 2))
    plan = b200fft.plan_fft("float64", "float64", x.shape, x.shape)
    desc = plan.describe()
    # 10000 points: one in-place tile (160 KB); 40000: 640 KB per row, two passes
    assert ("jitrowsIP10000" if n == 10000 else "split n=40000") in desc and "_f64" in desc, desc
    out = torch.full(x.shape, float("nan"), device="cuda", dtype=torch.float64)
    b200fft.fft(out, torch.from_numpy(x).cuda(), plan=plan)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    want = np.fft.fft(x[..., 0] + 1j * x[..., 1], axis=1)
    assert np.linalg.norm((got[..., 0] + 1j * got[..., 1]) - want) <= 1e-13 * np.linalg.norm(want), desc
    plan.destroy()
