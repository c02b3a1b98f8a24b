"""Plan-time specialisation tier (csrc/jit.cu): axis lengths without a hand-registered variant get the compile-time
kernels of fast.cuh instantiated through NVRTC when the plan is created. Parity against numpy float64 and the oracle
at the stated fp32 tolerance, that the tier is the one that runs, and that it beats the runtime-length tier."""
import numpy as np
import pytest

import b200fft

pytestmark = pytest.mark.gpu


def c2(a):
    a = np.asarray(a, dtype=np.float64)
    return a[..., 0] + 1j * a[..., 1]


CASES = [
    # shape, bases, inverse, comps
    ((50, 1000), None, False, 2),
    ((50, 1000), [[10, 10, 10]], True, 2),
    ((33, 100), [[5, 2]], False, 2),
    ((9, 343), None, False, 2),                # 7^3
    ((6, 31 * 29), None, False, 2),            # two large primes, one stage each
    ((4, 2187), [[3]], True, 2),               # 3^7 -> (27)(27)(3)
    ((3, 30, 21), None, False, 2),
    ((2, 12, 100, 18), None, True, 2),
    ((5, 600), None, False, 1),                # real input, full spectrum
    ((2, 4000), None, False, 2),
    ((1, 10000), None, False, 2),              # four stages, one row per CTA
    ((7, 17, 1000), None, False, 2),           # long strided axis next to a long contiguous one
    ((11, 74), None, False, 2),                # 37 x 2: a prime radix above 32 (the reference's prime list, fft.mojo:83-104)
    ((5, 61 * 3), [[61, 3]], True, 2),
    ((3, 6, 10), None, False, 2),              # tiles wider than the strided axis's inner extent
    ((2, 90, 5), None, False, 2),              # inner extent 5 < 8 columns
    ((300, 210), None, False, 2),
    # primes above 64: looped codelets over a constant table (dft.cuh: looped_prime) — the reference's prime list goes to 97
    ((11, 194), [[97, 2]], False, 2),
    ((5, 67), None, True, 2),
    ((4, 7, 89 * 3), None, False, 2),
    ((3, 127), [[127]], False, 2),
    ((6, 83, 4), None, False, 2),              # a large prime on a strided axis
]


@pytest.mark.parametrize("shape,bases,inverse,comps", CASES)
def test_jit_tier(oracle, shape, bases, inverse, comps):
    import torch
    rng = np.random.default_rng(23)
    full = shape + (comps,)
    x = rng.standard_normal(full).astype(np.float32)
    plan = b200fft.plan_fft("float32", "float32", full, shape + (2,), bases=bases, inverse=inverse)
    desc = plan.describe()
    assert "jit" in desc and "NVRTC" in desc and "rt_" not in desc and "generic" not in desc, desc
    out = torch.full(shape + (2,), float("nan"), device="cuda")
    before = b200fft.launch_count()
    b200fft.fft(out, torch.from_numpy(x).cuda(), plan=plan)
    torch.cuda.synchronize()
    assert b200fft.launch_count() - before == plan.launches
    got = c2(out.cpu().numpy())
    xd = x.astype(np.float64)
    xc = xd[..., 0] if comps == 1 else c2(xd)
    axes = tuple(range(1, len(shape)))
    want = np.fft.ifftn(xc, axes=axes) if inverse else np.fft.fftn(xc, axes=axes)
    assert np.isfinite(got).all()
    tol = 2e-6 * max(1.0, np.sqrt(len(axes)))
    assert np.linalg.norm(got - want) <= tol * np.linalg.norm(want), desc
    ref = c2(oracle.ref_fft(x, bases=bases, inverse=inverse))
    assert np.linalg.norm(got - ref) <= 5e-6 * max(1.0, np.sqrt(len(axes))) * np.linalg.norm(want)
    # a second plan of the same shape reuses the compiled module (no second compile: same text, same kernel)
    again = b200fft.plan_fft("float32", "float32", full, shape + (2,), bases=bases, inverse=inverse)
    assert again.describe() == desc
    out2 = torch.empty_like(out)
    b200fft.fft(out2, torch.from_numpy(x).cuda(), plan=again)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    again.destroy()
    plan.destroy()


def test_jit_can_be_disabled(monkeypatch):
    monkeypatch.setenv("B200FFT_JIT", "0")
    p = b200fft.plan_fft("float32", "float32", (4, 1000, 2), (4, 1000, 2))
    assert "rt_rows" in p.describe()
    p.destroy()


def test_jit_beats_the_runtime_length_tier():
    import torch
    x = torch.randn((50000, 1000, 2), device="cuda")
    out = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    times = {}
    for name, kw in (("jit", {}), ("rt", {"flags": b200fft.FLAG_FORCE_RT})):
        plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape, **kw)
        assert name in plan.describe()
        for _ in range(3):
            plan.exec(out, x, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            plan.exec(out, x, st)
        e1.record()
        torch.cuda.synchronize()
        times[name] = e0.elapsed_time(e1) / 10
        plan.destroy()
    print("50000 x 1000: jit %.3f ms, rt %.3f ms" % (times["jit"], times["rt"]))
    assert times["jit"] < times["rt"]


HALF_CASES = [
    # real-side shape, inverse (C2R)
    ((40, 1000), False), ((40, 1000), True),
    ((64, 200), False), ((64, 200), True),
    ((30, 243), False),                       # odd length: the n-point row kernel on real rows
    ((30, 243), True),                        # ... and its inverse: Hermitian-extended load, real rows stored
    ((4, 10, 75), True),
    ((7, 50, 600), False), ((7, 50, 600), True),
    ((3, 12, 10, 100), False), ((3, 12, 10, 100), True),
    ((9, 6), False), ((9, 6), True),
    ((5, 2000), False), ((5, 2000), True),
    ((16, 2 * 37 * 4), False),               # H = 148 = 37 x 4
]


@pytest.mark.parametrize("shape,inverse", HALF_CASES)
def test_jit_half_spectrum(shape, inverse):
    """Half-spectrum R2C / C2R of unregistered lengths on the specialised kernels, against numpy rfftn / irfftn."""
    import torch
    rng = np.random.default_rng(29)
    axes = tuple(range(1, len(shape)))
    cshape = shape[:-1] + (shape[-1] // 2 + 1,)
    real = rng.standard_normal(shape).astype(np.float32)
    if not inverse:
        plan = b200fft.plan_fft("float32", "float32", shape + (1,), cshape + (2,), real_mode=b200fft.REAL_HALF)
        desc = plan.describe()
        assert "jitr2c" in desc and "generic" not in desc and "rt_" not in desc, desc
        out = torch.full(cshape + (2,), float("nan"), device="cuda")
        b200fft.fft(out, torch.from_numpy(real).cuda().unsqueeze(-1), plan=plan)
        torch.cuda.synchronize()
        got = c2(out.cpu().numpy())
        want = np.fft.rfftn(real.astype(np.float64), axes=axes)
    else:
        spec = np.fft.rfftn(real.astype(np.float64), axes=axes)
        x = np.stack([spec.real, spec.imag], axis=-1).astype(np.float32)
        plan = b200fft.plan_fft("float32", "float32", cshape + (2,), shape + (1,), real_mode=b200fft.REAL_HALF, inverse=True)
        desc = plan.describe()
        assert "jitc2r" in desc and "generic" not in desc and "rt_" not in desc, desc
        out = torch.full(shape + (1,), float("nan"), device="cuda")
        b200fft.fft(out, torch.from_numpy(x).cuda(), plan=plan)
        torch.cuda.synchronize()
        got = out.cpu().numpy()[..., 0].astype(np.float64)
        want = np.fft.irfftn(c2(x), s=shape[1:], axes=axes)
    assert np.isfinite(got).all()
    tol = 2e-6 * max(1.0, np.sqrt(len(axes)))
    assert np.linalg.norm(got - want) <= tol * np.linalg.norm(want), desc
    plan.destroy()


def test_jit_scattering_store_in_the_slab_decomposition(monkeypatch):
    """b200fft_exec_scatter on a split axis without a registered variant: the scattering kernel is specialised on first
    use (virtual device slots on one GPU when fewer than two are visible)."""
    import torch
    ngpu = 2
    have = torch.cuda.device_count()
    if have < ngpu:
        monkeypatch.setenv("B200FFT_MGPU_ALLOW_SAME_DEVICE", "1")
    devs = [g % have for g in range(ngpu)]
    Z, Y, X = 20, 120, 100
    plan = b200fft.MgpuPlan("float32", "float32", (1, Z, Y, X, 2), (1, Z, Y, X, 2), devices=devs, mode=b200fft.MGPU_SLAB)
    assert "jitcols120" in plan.describe(), plan.describe()
    rng = np.random.default_rng(31)
    x = rng.standard_normal((Z, Y, X, 2)).astype(np.float32)
    h_out = np.full((Z, Y, X, 2), np.nan, dtype=np.float32)
    plan.exec_host(h_out, x)
    want = np.fft.fftn(c2(x))
    assert np.linalg.norm(c2(h_out) - want) <= 2e-6 * np.sqrt(3) * np.linalg.norm(want)
    plan.destroy()


F64_CASES = [
    # shape, in dtype, comps, inverse, bases
    ((20, 1024), "float64", 2, False, None),
    ((20, 1024), "float64", 2, True, None),
    ((6, 64, 64, 64), "float64", 2, False, None),           # registered fp32 lengths: fp64 goes through NVRTC too
    ((9, 93), "float64", 2, False, [[31, 3]]),
    ((4, 640, 480), "float64", 2, True, None),
    ((5, 100), "uint8", 1, False, None),                     # the reference's tests: uint8 in, float64 out (tests.mojo:394)
    ((3, 6, 4, 8), "uint8", 1, False, None),
    ((7, 1000), "float32", 2, False, None),                  # fp32 in, fp64 out
    ((3, 74), "float64", 2, False, None),                    # 37 x 2
    ((3, 97), "float64", 2, False, None),                    # looped radix-97 codelet in fp64
    ((2, 30, 21), "float64", 1, True, None),
]


@pytest.mark.parametrize("shape,in_dtype,comps,inverse,bases", F64_CASES)
def test_jit_fp64_output(oracle, shape, in_dtype, comps, inverse, bases):
    """fp64 output no longer lands on the generic kernel: every axis runs the fp64 build of the compile-time kernels.
    Tolerance: relative L2 <= 1e-13 x sqrt(axes) against numpy float64 (fp64 round-off of a log-depth transform)."""
    import torch
    rng = np.random.default_rng(37)
    full = shape + (comps,)
    x = rng.integers(0, 256, size=full).astype(np.uint8) if in_dtype == "uint8" else rng.standard_normal(full).astype(in_dtype)
    plan = b200fft.plan_fft(in_dtype, "float64", full, shape + (2,), bases=bases, inverse=inverse)
    desc = plan.describe()
    assert "_f64" in desc and "generic" not in desc and "rt_" not in desc, desc
    out = torch.full(shape + (2,), float("nan"), device="cuda", dtype=torch.float64)
    b200fft.fft(out, torch.from_numpy(x).cuda(), plan=plan)
    torch.cuda.synchronize()
    got = c2(out.cpu().numpy())
    xd = x.astype(np.float64)
    xc = xd[..., 0] if comps == 1 else c2(xd)
    axes = tuple(range(1, len(shape)))
    want = np.fft.ifftn(xc, axes=axes) if inverse else np.fft.fftn(xc, axes=axes)
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - want) <= 1e-13 * np.sqrt(len(axes)) * np.linalg.norm(want), desc
    ref = c2(oracle.ref_fft(x, bases=bases, inverse=inverse, out_dtype=np.float64))
    assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(want)
    plan.destroy()


@pytest.mark.parametrize("shape,inverse", [((12, 1000), False), ((12, 1000), True), ((3, 40, 128), False), ((3, 40, 128), True),
                                           ((8, 243), False)])
def test_jit_fp64_half_spectrum(shape, inverse):
    import torch
    rng = np.random.default_rng(41)
    axes = tuple(range(1, len(shape)))
    cshape = shape[:-1] + (shape[-1] // 2 + 1,)
    real = rng.standard_normal(shape)
    if not inverse:
        plan = b200fft.plan_fft("float64", "float64", shape + (1,), cshape + (2,), real_mode=b200fft.REAL_HALF)
        desc = plan.describe()
        assert "jitr2c" in desc and "_f64" in desc and "generic" not in desc, desc
        out = torch.full(cshape + (2,), float("nan"), device="cuda", dtype=torch.float64)
        b200fft.fft(out, torch.from_numpy(real).cuda().unsqueeze(-1), plan=plan)
        torch.cuda.synchronize()
        got = c2(out.cpu().numpy())
        want = np.fft.rfftn(real, axes=axes)
    else:
        spec = np.fft.rfftn(real, axes=axes)
        x = np.stack([spec.real, spec.imag], axis=-1)
        plan = b200fft.plan_fft("float64", "float64", cshape + (2,), shape + (1,), real_mode=b200fft.REAL_HALF, inverse=True)
        desc = plan.describe()
        assert "jitc2r" in desc and "_f64" in desc and "generic" not in desc, desc
        out = torch.full(shape + (1,), float("nan"), device="cuda", dtype=torch.float64)
        b200fft.fft(out, torch.from_numpy(x).cuda(), plan=plan)
        torch.cuda.synchronize()
        got = out.cpu().numpy()[..., 0]
        want = real
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - want) <= 1e-13 * np.sqrt(len(axes)) * np.linalg.norm(want), desc
    plan.destroy()


def test_jit_fp32_output_from_other_input_types(oracle):
    """uint8 / fp64 arrays into an fp32 transform: the specialised kernels cast on load (in_scalar, rtc_prelude.cuh)."""
    import torch
    rng = np.random.default_rng(43)
    for shape, in_dtype, comps in (((5, 600), "uint8", 1), ((3, 50, 36), "uint8", 2), ((4, 225), "float64", 2)):
        full = shape + (comps,)
        x = rng.integers(0, 256, size=full).astype(np.uint8) if in_dtype == "uint8" else rng.standard_normal(full)
        plan = b200fft.plan_fft(in_dtype, "float32", full, shape + (2,))
        desc = plan.describe()
        assert "jit" in desc and ("_inu8" in desc or "_inf64" in desc) and "generic" not in desc, desc
        out = torch.full(shape + (2,), float("nan"), device="cuda")
        b200fft.fft(out, torch.from_numpy(x).cuda(), plan=plan)
        torch.cuda.synchronize()
        xd = x.astype(np.float64)
        xc = xd[..., 0] if comps == 1 else c2(xd)
        want = np.fft.fftn(xc, axes=tuple(range(1, len(shape))))
        got = c2(out.cpu().numpy())
        assert np.linalg.norm(got - want) <= 2e-6 * np.sqrt(len(shape) - 1) * np.linalg.norm(want), desc
        plan.destroy()
