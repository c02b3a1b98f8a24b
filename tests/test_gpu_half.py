"""Half-spectrum R2C / C2R (B200FFT_REAL_HALF, north-star piece 3) against numpy rfftn / irfftn.

The reference itself has no half-spectrum mode (its real-input path writes the full
spectrum, _fft.mojo:254-255), so numpy float64 is the oracle here and the
reference-compatible full-spectrum mode is cross-checked against the half one.
Tolerances: the fp32 bounds of test_gpu_parity (rel-L2 2e-6, max-abs 1e-5*max|X|)."""
import numpy as np
import pytest

import b200fft

pytestmark = pytest.mark.gpu


def c2(a):
    a = np.asarray(a, dtype=np.float64)
    return a[..., 0] + 1j * a[..., 1]


def r2c(x, generic=False, dtype=np.float32, bases=None):
    import torch
    shape = x.shape[:-1]
    out_shape = shape[:-1] + (shape[-1] // 2 + 1, 2)
    xt = torch.from_numpy(x).cuda()
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    out = torch.full(out_shape, float("nan"), dtype=tdt, device="cuda")
    plan = b200fft.plan_fft(str(x.dtype), np.dtype(dtype).name, x.shape, out_shape, real_mode=b200fft.REAL_HALF,
                            bases=bases, _test=("generic" if generic else None))
    b200fft.fft(out, xt, plan=plan)
    torch.cuda.synchronize()
    desc = plan.describe()
    plan.destroy()
    return out.cpu().numpy(), desc


def c2r(X, n_last, generic=False, dtype=np.float32):
    import torch
    out_shape = X.shape[:-2] + (n_last, 1)
    Xt = torch.from_numpy(X).cuda()
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    out = torch.full(out_shape, float("nan"), dtype=tdt, device="cuda")
    plan = b200fft.plan_fft(str(X.dtype), np.dtype(dtype).name, X.shape, out_shape, real_mode=b200fft.REAL_HALF,
                            inverse=True, _test=("generic" if generic else None))
    b200fft.fft(out, Xt, plan=plan)
    torch.cuda.synchronize()
    ws = plan.workspace_bytes
    plan.destroy()
    return out.cpu().numpy(), ws


SHAPES = [(8, 128), (5, 64), (3, 16), (7, 480), (4, 1024), (3, 640, 480), (2, 64, 64, 64), (2, 20, 12), (3, 6, 10, 8),
          (2, 256), (3, 2), (2, 640), (2, 512), (1, 128, 128, 128),
          (5, 93), (2, 7), (3, 15, 9), (2, 6, 4, 21)]  # odd last axis: generic Hermitian kernel


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_r2c_vs_numpy(shape, generic):
    rng = np.random.default_rng(abs(hash(shape)) % 2**32)
    x = rng.standard_normal(shape + (1,)).astype(np.float32)
    want = np.fft.rfftn(x[..., 0].astype(np.float64), axes=tuple(range(1, len(shape))))
    got, desc = r2c(x, generic)
    got = c2(got)
    assert got.shape == want.shape
    assert np.isfinite(got).all(), desc
    assert np.linalg.norm(got - want) <= 2e-6 * np.linalg.norm(want), desc
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), desc


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_c2r_vs_numpy(shape, generic):
    rng = np.random.default_rng(abs(hash(shape)) % 2**32 + 1)
    x = rng.standard_normal(shape).astype(np.float32)
    axes = tuple(range(1, len(shape)))
    X = np.fft.rfftn(x.astype(np.float64), axes=axes)
    Xin = np.stack([X.real, X.imag], axis=-1).astype(np.float32)
    got, ws = c2r(Xin, shape[-1], generic)
    want = np.fft.irfftn(c2(Xin), s=shape[1:], axes=axes)
    assert np.isfinite(got).all()
    assert np.linalg.norm(got[..., 0] - want) <= 2e-6 * np.linalg.norm(want)
    assert np.abs(got[..., 0] - want).max() <= 1e-5 * np.abs(want).max()
    assert (ws == 0) == (len(shape) == 2)   # only N-d needs the half-spectrum workspace


def test_fast_path_is_used_for_the_benchmark_shape():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 640, 480, 1)).astype(np.float32)
    _, desc = r2c(x)
    assert "r2c" in desc and "[rows240" in desc and "cols640" in desc, desc


def test_half_matches_reference_full_mode():
    """The reference-compatible full-spectrum real transform restricted to bins 0..n/2."""
    import torch
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 48, 128, 1)).astype(np.float32)
    half, _ = r2c(x)
    xt = torch.from_numpy(x).cuda()
    full = torch.empty((4, 48, 128, 2), device="cuda")
    plan = b200fft.plan_fft("float32", "float32", x.shape, full.shape)
    b200fft.fft(full, xt, plan=plan)
    torch.cuda.synchronize()
    f = full.cpu().numpy()[:, :, :65]
    assert np.linalg.norm(f - half) <= 2e-6 * np.linalg.norm(f)


def test_r2c_c2r_round_trip_full_size():
    """BASELINE config 2-D R2C at full size: irfft(rfft(x)) == x."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((100, 640, 480, 1), generator=g, device="cuda")
    X = torch.full((100, 640, 241, 2), float("nan"), device="cuda")
    y = torch.full_like(x, float("nan"))
    f = b200fft.plan_fft("float32", "float32", x.shape, X.shape, real_mode=b200fft.REAL_HALF)
    i = b200fft.plan_fft("float32", "float32", X.shape, y.shape, real_mode=b200fft.REAL_HALF, inverse=True)
    b200fft.fft(X, x, plan=f)
    b200fft.fft(y, X, plan=i)
    torch.cuda.synchronize()
    assert float((y - x).double().norm() / x.double().norm()) < 2e-6
    want = torch.fft.rfftn(x[0, ..., 0].double(), dim=(0, 1))
    got = torch.view_as_complex(X[0].double().contiguous())
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5


def test_f64_and_errors():
    rng = np.random.default_rng(9)
    x = rng.standard_normal((3, 30, 1))
    got, _ = r2c(x, dtype=np.float64)
    want = np.fft.rfft(x[..., 0], axis=1)
    assert np.linalg.norm(c2(got) - want) <= 1e-13 * np.linalg.norm(want)
    with pytest.raises(b200fft.B200FFTError) as e:     # uint8 input is not accepted in half-spectrum mode
        b200fft.plan_fft("uint8", "float32", (2, 32, 1), (2, 17, 2), real_mode=b200fft.REAL_HALF)
    assert e.value.status == 4
    with pytest.raises(b200fft.B200FFTError) as e:     # wrong number of bins for the real length
        b200fft.plan_fft("float32", "float32", (2, 93, 1), (2, 46, 2), real_mode=b200fft.REAL_HALF)
    assert e.value.status == 2


def test_unregistered_lengths_use_the_rt_tier(monkeypatch):
    """Half-spectrum rows without a registered R2C/C2R kernel: with the plan-time specialisation tier switched off
    (B200FFT_JIT=0; with it on they get an NVRTC-built R2C kernel, tests/test_gpu_jit.py) they run on the runtime-length
    tier (n-point transform, bins 0..n/2 stored / Hermitian-extended load), not on the generic kernel; an odd length with
    a registered row variant (93 = 31 x 3) runs its R2C on that compile-time kernel with a half-bins store (C2R of odd n
    stays on the rt tier)."""
    monkeypatch.setenv("B200FFT_JIT", "0")
    rng = np.random.default_rng(4)
    for shape, kernel in (((6, 93), "r2c-odd[rows93"), ((3, 1000), "rt_rows"), ((2, 12, 30), "rt_rows"),
                          ((2, 20, 93), "r2c-odd[rows93")):
        x = rng.standard_normal(shape + (1,)).astype(np.float32)
        got, desc = r2c(x)
        assert kernel in desc and "generic" not in desc, desc
        want = np.fft.rfftn(x[..., 0].astype(np.float64), axes=tuple(range(1, len(shape))))
        assert np.linalg.norm(c2(got) - want) <= 2e-6 * np.sqrt(len(shape) - 1) * np.linalg.norm(want)
        back, _ = c2r(got, shape[-1])
        assert np.linalg.norm(back[..., 0] - x[..., 0]) <= 4e-6 * np.linalg.norm(x)


def test_odd_c2r_on_the_compile_time_kernel():
    """C2R of an odd length with a registered row variant (93 = 31 x 3): the n-point inverse on the Hermitian-extended
    row (rows_c2r_odd_kernel), no longer the runtime-length tier."""
    import torch
    rng = np.random.default_rng(8)
    for shape in ((64, 93), (5, 12, 93)):
        real = rng.standard_normal(shape)
        axes = tuple(range(1, len(shape)))
        spec = np.fft.rfftn(real, axes=axes)
        x = np.stack([spec.real, spec.imag], axis=-1).astype(np.float32)
        plan = b200fft.plan_fft("float32", "float32", x.shape, shape + (1,), real_mode=b200fft.REAL_HALF, inverse=True)
        desc = plan.describe()
        assert "c2r-odd[rows93" in desc and "rt_" not in desc and "generic" not in desc, desc
        out = torch.full(shape + (1,), float("nan"), device="cuda")
        b200fft.fft(out, torch.from_numpy(x).cuda(), plan=plan)
        torch.cuda.synchronize()
        got = out.cpu().numpy()[..., 0].astype(np.float64)
        assert np.linalg.norm(got - real) <= 2e-6 * np.sqrt(len(axes)) * np.linalg.norm(real), desc
        plan.destroy()


@pytest.mark.parametrize("shape", [(5, 64, 64, 64), (3, 128, 128, 128), (9, 64, 64), (2, 3, 5, 64, 64), (1, 6, 128, 128)])
def test_plane_c2r_kernel(shape, monkeypatch):
    """Half-spectrum inverse: the two innermost axes in ONE tile per (y, x) plane (csrc/plane.cuh) instead of a strided pass
    over the ragged n/2+1 extent + a row pass. Against numpy irfftn, and against the per-axis plan of the same library."""
    import torch
    if shape[-1] == 128:
        monkeypatch.setenv("B200FFT_PLANE_C2R", "1")   # the 128 x 128 variant is opt-in (one CTA per SM: slower than per axis)
    rng = np.random.default_rng(12)
    axes = tuple(range(1, len(shape)))
    real = rng.standard_normal(shape)
    spec = np.fft.rfftn(real, axes=axes)
    x = torch.from_numpy(np.stack([spec.real, spec.imag], axis=-1).astype(np.float32)).cuda()
    keep = x.clone()
    plan = b200fft.plan_fft("float32", "float32", tuple(x.shape), shape + (1,), real_mode=b200fft.REAL_HALF, inverse=True)
    desc = plan.describe()
    assert "c2rplane%dx%d" % (shape[-2], shape[-1]) in desc, desc
    assert plan.launches == len(shape) - 2          # one pass per outer axis + the plane pass
    out = torch.full(shape + (1,), float("nan"), device="cuda")
    b200fft.fft(out, x, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(x, keep)                      # the half spectrum is not overwritten
    got = out.cpu().numpy()[..., 0].astype(np.float64)
    assert np.linalg.norm(got - real) <= 2e-6 * np.sqrt(len(axes)) * np.linalg.norm(real), desc
    monkeypatch.setenv("B200FFT_PLANE_C2R", "0")
    plain = b200fft.plan_fft("float32", "float32", tuple(x.shape), shape + (1,), real_mode=b200fft.REAL_HALF, inverse=True)
    assert "c2rplane" not in plain.describe()
    out2 = torch.empty_like(out)
    b200fft.fft(out2, x, plan=plain)
    torch.cuda.synchronize()
    assert float((out - out2).norm() / out2.norm()) < 1e-6
    plan.destroy()
    plain.destroy()
