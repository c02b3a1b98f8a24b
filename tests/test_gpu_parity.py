"""GPU parity: the CUDA path, called through the C ABI (b200fft_plan_create /
b200fft_exec via the ctypes shim), against the oracle (oracle/ref_fft.cpp), the
reference's golden vectors and numpy float64.

Stated fp32 tolerance, per transform on N(0,1) data:
  vs numpy float64:      relative L2 <= 2e-6, max-abs <= 1e-5 * max|X|
  vs the fp32 oracle:    relative L2 <= 5e-6, max-abs <= 2e-5 * max|X|  (covers the
                         reference's fp32-theta twiddle error, not ours)
  golden vectors:        the reference's own atol=1e-2, rtol=1e-5 (tests.mojo:40-41)
fp64: relative L2 <= 1e-13.
"""
import json
import os

import numpy as np
import pytest

import b200fft

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
    _G = json.load(f)
CASES = [(c["length"], tuple(c["bases"])) for c in _G["cases_1d"]]

RTOL_L2_NP, RTOL_MAX_NP = 2e-6, 1e-5
RTOL_L2_ORACLE, RTOL_MAX_ORACLE = 5e-6, 2e-5


def c2(a):
    a = np.asarray(a, dtype=np.float64)
    return a[..., 0] + 1j * a[..., 1]


def run_gpu(x, *, bases=None, inverse=False, out_dtype=np.float32, generic=False, nan_fill=True, tier=None, describe=None):
    """plan_fft + fft on device buffers; output prefilled with NaN like the reference's
    tests (tests.mojo:219-222) to catch unwritten elements."""
    import torch
    xt = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out_layout = tuple(x.shape[:-1]) + (2,)
    tdt = torch.float32 if np.dtype(out_dtype) == np.float32 else torch.float64
    out = torch.full(out_layout, float("nan"), dtype=tdt, device="cuda")
    plan = b200fft.plan_fft(str(x.dtype), np.dtype(out_dtype).name, x.shape, out_layout, bases=bases,
                            inverse=inverse, _test=(tier if tier is not None else "generic" if generic else None))
    if describe is not None:
        describe.append(plan.describe())
    b200fft.fft(out, xt, plan=plan)
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    plan.destroy()
    return res


def check_vs(got, want, l2, mx):
    got, want = c2(got), np.asarray(want)
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - want) <= l2 * np.linalg.norm(want)
    assert np.abs(got - want).max() <= mx * np.abs(want).max()


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_1d_golden_forward_and_inverse(golden, dtype, generic):
    """fft/tests.mojo `_test_fft` (:274-371): 56 (length, bases) x forward / inverse."""
    for length, bases in CASES:
        vs = golden["vectors_1d"][str(length)]
        x = np.array([v["x"] for v in vs], dtype=dtype)[:, :, None]
        spec = np.array([v["X"] for v in vs], dtype=np.float64)
        got = run_gpu(x, bases=[list(bases)], out_dtype=dtype, generic=generic)
        np.testing.assert_allclose(got, spec, atol=golden["atol"], rtol=golden["rtol"], err_msg=str((length, bases)))
        back = run_gpu(spec.astype(dtype), bases=[list(bases)], inverse=True, out_dtype=dtype, generic=generic)
        np.testing.assert_allclose(back[..., 0], x[..., 0], atol=golden["atol"], rtol=golden["rtol"])
        np.testing.assert_allclose(back[..., 1], 0, atol=golden["atol"])


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("key", ["2d", "3d"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_nd_golden_uint8(golden, key, dtype, generic):
    """test_2d_gpu / test_3d_gpu (tests.mojo:521-610, 973-1071): uint8 in, float out."""
    d = golden["nd"][key]
    x = np.array(d["x"], dtype=np.uint8).reshape([1] + d["dims"] + [1])
    want = np.array(d["X"], dtype=np.float64).reshape([1] + d["dims"] + [2])
    got = run_gpu(x, out_dtype=dtype, generic=generic)
    np.testing.assert_allclose(got, want, atol=golden["atol"], rtol=golden["rtol"])


@pytest.mark.parametrize("tier", [b200fft.GPUTest.BLOCK, b200fft.GPUTest.WARP, b200fft.GPUTest.DEVICE_WIDE,
                                  b200fft.GPUTest.CLUSTER])
def test_gpu_test_tiers_on_the_golden_vectors(golden, tier):
    """fft/tests.mojo:398-417 runs its vectors once per `_GPUTest` value; here each value forces one kernel tier of
    this library (b200fft.GPUTest) and the plan's description shows the tier really ran."""
    names = {b200fft.GPUTest.BLOCK: ("rows", "cols"), b200fft.GPUTest.WARP: ("rt_", "generic"),
             b200fft.GPUTest.DEVICE_WIDE: ("fused ", "rows", "cols"), b200fft.GPUTest.CLUSTER: ("generic",)}[tier]
    for length, bases in CASES[::3]:
        vs = golden["vectors_1d"][str(length)]
        x = np.array([v["x"] for v in vs], dtype=np.float32)[:, :, None]
        spec = np.array([v["X"] for v in vs], dtype=np.float64)
        text = []
        got = run_gpu(x, bases=[list(bases)], tier=tier, describe=text)
        np.testing.assert_allclose(got, spec, atol=golden["atol"], rtol=golden["rtol"], err_msg=str((length, bases)))
        if tier in (b200fft.GPUTest.WARP, b200fft.GPUTest.CLUSTER):
            assert any(n in text[0] for n in names), text[0]
    # a shape every tier covers, N-d: the fused kernel really is picked by DEVICE_WIDE and never by BLOCK
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 64, 64, 64, 2)).astype(np.float32)
    text = []
    got = run_gpu(x, tier=tier, describe=text)
    check_vs(got, np.fft.fftn(c2(x), axes=(1, 2, 3)), RTOL_L2_NP, RTOL_MAX_NP)
    assert text[0].startswith("fused ") == (tier == b200fft.GPUTest.DEVICE_WIDE), text[0]
    assert any(n in text[0] for n in names), text[0]


RANDOM_SHAPES = [
    ((64, 128), None), ((7, 128), [[16, 8]]), ((5, 128), [[2]]), ((16, 1024), None), ((3, 1024), [[32]]),
    ((33, 93), None), ((9, 93), [[3, 31]]), ((4, 97), [[97]]), ((2, 194), [[97, 2]]), ((3, 2048), None),
    ((2, 640, 480), None), ((3, 48, 40), [[6, 8], [5, 2]]), ((3, 64, 64, 64), None), ((1, 128, 128, 128), None),
    ((2, 6, 10, 12, 14), None), ((1, 16, 16, 16, 16), None), ((5, 2), None), ((1, 3, 5), None),
    ((2, 512), None), ((3, 256), None), ((2, 480), None), ((2, 640), None), ((1, 4096), None),
]


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("shape,bases", RANDOM_SHAPES)
@pytest.mark.parametrize("inverse", [False, True])
def test_random_c2c_vs_oracle_and_numpy(oracle, shape, bases, inverse, generic):
    rng = np.random.default_rng(abs(hash((shape, inverse))) % 2**32)
    x = rng.standard_normal(shape + (2,)).astype(np.float32)
    axes = tuple(range(1, len(shape)))
    xc = c2(x)
    want = np.fft.ifftn(xc, axes=axes) if inverse else np.fft.fftn(xc, axes=axes)
    got = run_gpu(x, bases=bases, inverse=inverse, generic=generic)
    check_vs(got, want, RTOL_L2_NP, RTOL_MAX_NP)
    ref = oracle.ref_fft(x, bases=bases, inverse=inverse)
    check_vs(got, c2(ref), RTOL_L2_ORACLE, RTOL_MAX_ORACLE)


@pytest.mark.parametrize("shape", [(4, 128), (2, 96, 80), (2, 8, 12, 10)])
def test_fp64_vs_numpy(shape):
    rng = np.random.default_rng(5)
    x = rng.standard_normal(shape + (2,))
    want = np.fft.fftn(c2(x), axes=tuple(range(1, len(shape))))
    got = c2(run_gpu(x, out_dtype=np.float64))
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)


@pytest.mark.parametrize("shape", [(5, 128), (3, 93), (2, 640, 480), (2, 20, 12), (2, 8, 6, 10)])
def test_real_input_full_spectrum(oracle, shape):
    """in_layout (..., 1): reference 'rfft' semantics = FULL spectrum (_fft.mojo:254-255)."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape + (1,)).astype(np.float32)
    want = np.fft.fftn(x[..., 0].astype(np.float64), axes=tuple(range(1, len(shape))))
    got = run_gpu(x)
    check_vs(got, want, RTOL_L2_NP, RTOL_MAX_NP)
    check_vs(got, c2(oracle.ref_fft(x)), RTOL_L2_ORACLE, RTOL_MAX_ORACLE)


def test_mixed_dtypes(oracle):
    rng = np.random.default_rng(2)
    x8 = rng.integers(0, 256, size=(3, 24, 20, 2), dtype=np.uint8)
    want = np.fft.fftn(c2(x8), axes=(1, 2))
    check_vs(run_gpu(x8), want, RTOL_L2_NP, RTOL_MAX_NP)
    x64 = rng.standard_normal((3, 60, 2))
    check_vs(run_gpu(x64, out_dtype=np.float32), np.fft.fft(c2(x64), axis=1), RTOL_L2_NP, RTOL_MAX_NP)
    x32 = x64.astype(np.float32)
    got = c2(run_gpu(x32, out_dtype=np.float64))
    want = np.fft.fft(c2(x32), axis=1)
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)


def test_in_place_and_plan_reuse():
    import torch
    rng = np.random.default_rng(4)
    x = rng.standard_normal((6, 32, 24, 2)).astype(np.float32)
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    buf = torch.from_numpy(x).cuda()
    want = np.fft.fftn(c2(x), axes=(1, 2))
    for _ in range(2):                      # same plan, twice
        buf.copy_(torch.from_numpy(x))
        b200fft.fft(buf, buf, plan=plan)    # in place
        torch.cuda.synchronize()
        check_vs(buf.cpu().numpy(), want, RTOL_L2_NP, RTOL_MAX_NP)
    assert plan.workspace_bytes == 0
    plan.destroy()


def test_exec_host_matches_device_path():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((37, 128, 2)).astype(np.float32)   # ragged chunking: 37 batches
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    out = np.full(x.shape, np.nan, dtype=np.float32)
    plan.exec_host(out, x)
    check_vs(out, np.fft.fft(c2(x), axis=1), RTOL_L2_NP, RTOL_MAX_NP)
    plan.destroy()


def test_stream_argument():
    import torch
    rng = np.random.default_rng(9)
    x = rng.standard_normal((8, 1024, 2)).astype(np.float32)
    s = torch.cuda.Stream()
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    xt = torch.from_numpy(x).cuda()
    out = torch.empty_like(xt)
    torch.cuda.synchronize()
    b200fft.fft(out, xt, s, plan=plan)
    s.synchronize()
    check_vs(out.cpu().numpy(), np.fft.fft(c2(x), axis=1), RTOL_L2_NP, RTOL_MAX_NP)


def test_errors_are_reported_not_thrown():
    with pytest.raises(b200fft.B200FFTError) as e:
        b200fft.plan_fft("float32", "float32", (2, 60, 2), (2, 60, 2), bases=[[7, 2]])
    assert e.value.status == 3
    import torch
    plan = b200fft.plan_fft("float32", "float32", (2, 8, 2), (2, 8, 2))
    with pytest.raises(b200fft.B200FFTError):
        b200fft.fft(torch.zeros(2, 8, 2), torch.zeros(2, 8, 2), plan=plan)   # CPU tensors: no CPU path
    with pytest.raises(b200fft.B200FFTError):
        b200fft.fft(torch.zeros(2, 9, 2, device="cuda"), torch.zeros(2, 8, 2, device="cuda"), plan=plan)


# ---- BASELINE.json full sizes: size-independent properties + sampled rows vs oracle ----
FULL = [
    ("1d_500000x128", (500000, 128)), ("1d_100000x1024", (100000, 1024)), ("1d_500000x93", (500000, 93)),
    ("2d_100x640x480", (100, 640, 480)), ("3d_100x64^3", (100, 64, 64, 64)), ("3d_10x128^3", (10, 128, 128, 128)),
    ("3d_1x512^3", (1, 512, 512, 512)),
]


@pytest.mark.parametrize("name,shape", FULL)
def test_full_size_properties(oracle, name, shape):
    """Forward -> inverse round trip, Parseval, linearity, and sampled batch items against
    the oracle / numpy, at the exact BASELINE.json shapes (data generated on the device)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(shape + (2,), generator=g, device="cuda", dtype=torch.float32)
    out = torch.full_like(x, float("nan"))
    fwd = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
    inv = b200fft.plan_fft("float32", "float32", x.shape, x.shape, inverse=True)
    b200fft.fft(out, x, plan=fwd)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    n = int(np.prod(shape[1:]))
    # Parseval per batch item (first 64 items): sum|X|^2 = N * sum|x|^2
    k = min(64, shape[0])
    ex = (x[:k].double() ** 2).flatten(1).sum(1)
    eX = (out[:k].double() ** 2).flatten(1).sum(1)
    assert torch.allclose(eX, ex * n, rtol=1e-5)
    # sampled batch items vs numpy f64 and the fp32 oracle
    for b in sorted({0, shape[0] // 2, shape[0] - 1}):
        xb = x[b:b + 1].cpu().numpy()
        want = np.fft.fftn(c2(xb), axes=tuple(range(1, len(shape))))
        check_vs(out[b:b + 1].cpu().numpy(), want, RTOL_L2_NP, RTOL_MAX_NP)
        if n <= 2 ** 21:   # the oracle's radix-2 CPU path on 512^3 alone would take a minute
            check_vs(out[b:b + 1].cpu().numpy(), c2(oracle.ref_fft(xb)), RTOL_L2_ORACLE, RTOL_MAX_ORACLE)
    # round trip
    back = torch.full_like(x, float("nan"))
    b200fft.fft(back, out, plan=inv)
    torch.cuda.synchronize()
    err = (back - x).double().norm() / x.double().norm()
    assert err < 2e-6, float(err)
    # linearity: F(2x) == 2 F(x) exactly in fp32 (scaling by 2 is exact)
    x2 = x * 2
    b200fft.fft(back, x2, plan=fwd)
    torch.cuda.synchronize()
    assert torch.equal(back, out * 2)
    fwd.destroy(); inv.destroy()


@pytest.mark.parametrize("shape", [(1000, 128), (37, 1024), (65, 64), (33, 512), (77, 256)])
@pytest.mark.parametrize("inverse", [False, True])
def test_vectorised_rows_variants(shape, inverse, monkeypatch):
    """"rowsV" variants: 128-bit global loads/stores with the two-lane shuffle exchange (fast.cuh GlobalSrcV4)."""
    import torch
    monkeypatch.setenv("B200FFT_PREFER", "rowsV")
    rng = np.random.default_rng(5)
    x = rng.standard_normal(shape + (2,)).astype(np.float32)
    plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape, inverse=inverse)
    assert "rowsV" in plan.describe(), plan.describe()
    xt = torch.from_numpy(x).cuda()
    out = torch.full_like(xt, float("nan"))
    b200fft.fft(out, xt, plan=plan)
    torch.cuda.synchronize()
    want = (np.fft.ifft if inverse else np.fft.fft)(c2(x.astype(np.float64)), axis=1)
    check_vs(out.cpu().numpy(), want, RTOL_L2_NP, RTOL_MAX_NP)
    # in place, ragged last tile included (shape[0] is not a multiple of the rows per CTA)
    b200fft.fft(xt, xt, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(xt, out)
