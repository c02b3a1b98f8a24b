"""Pin the oracle (oracle/ref_fft.cpp) to the reference's own golden vectors.

Mirrors fft/tests.mojo: `_test_fft` (:274-371) x forward/inverse, `test_2d_cpu`
(:461-518), `test_3d_cpu` (:908-970), at the reference's atol=1e-2 / rtol=1e-5
(:40-41). The goldens are printed to 3 decimals, so the bound here is 2e-3.
"""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
    _G = json.load(f)
CASES = [(c["length"], tuple(c["bases"])) for c in _G["cases_1d"]]


def c2(a):
    return a[..., 0] + 1j * a[..., 1]


def test_golden_fixture_is_complete(golden):
    assert sum(len(v) for v in golden["vectors_1d"].values()) == 65
    assert len(golden["cases_1d"]) == 56
    assert golden["nd"]["2d"]["dims"] == [6, 4] and golden["nd"]["3d"]["dims"] == [6, 4, 8]


@pytest.mark.parametrize("length,bases", CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_1d_forward_golden(oracle, golden, length, bases, dtype):
    vs = golden["vectors_1d"][str(length)]
    x = np.array([v["x"] for v in vs], dtype=dtype)[:, :, None]  # (B, N, 1): real input
    want = np.array([v["X"] for v in vs], dtype=np.float64)
    got = oracle.ref_fft(x, bases=[list(bases)], out_dtype=dtype)
    np.testing.assert_allclose(got, want, atol=golden["atol"], rtol=golden["rtol"])
    assert np.abs(got - want).max() < 2e-3


@pytest.mark.parametrize("length,bases", CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_1d_inverse_golden(oracle, golden, length, bases, dtype):
    vs = golden["vectors_1d"][str(length)]
    spec = np.array([v["X"] for v in vs], dtype=dtype)  # (B, N, 2)
    want = np.array([v["x"] for v in vs], dtype=np.float64)
    got = oracle.ref_fft(spec, bases=[list(bases)], inverse=True, out_dtype=dtype)
    np.testing.assert_allclose(got[..., 0], want, atol=golden["atol"], rtol=golden["rtol"])
    np.testing.assert_allclose(got[..., 1], 0, atol=golden["atol"])


@pytest.mark.parametrize("key", ["2d", "3d"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_nd_golden_uint8_input(oracle, golden, key, dtype):
    d = golden["nd"][key]
    x = np.array(d["x"], dtype=np.uint8).reshape([1] + d["dims"] + [1])
    want = np.array(d["X"], dtype=np.float64).reshape([1] + d["dims"] + [2])
    got = oracle.ref_fft(x, out_dtype=dtype)
    np.testing.assert_allclose(got, want, atol=golden["atol"], rtol=golden["rtol"])


def test_ordered_bases_rules(oracle):
    # _utils.mojo:163-221 examples
    assert oracle.ordered_bases(8, [2]) == [2, 2, 2]
    assert oracle.ordered_bases(16, [2, 4]) == [4, 4]
    assert oracle.ordered_bases(20, [5, 2]) == [5, 2, 2]
    assert oracle.ordered_bases(48, [3, 2]) == [3, 2, 2, 2, 2]
    assert oracle.ordered_bases(60, [5, 3, 2]) == [5, 3, 2, 2]
    assert oracle.ordered_bases(60, [6, 5, 2]) == [6, 5, 2]
    assert oracle.ordered_bases(60, [3, 4, 5]) == [5, 4, 3]
    assert oracle.ordered_bases(93, [3, 31]) == [31, 3]
    assert oracle.ordered_bases(60, [7, 2]) is None     # product never reaches 60
    assert oracle.ordered_bases(8, [1, 2]) is None      # base 1 rejected


def test_default_bases(oracle):
    # fft.mojo:49-104
    assert oracle.default_bases(128, "gpu") == [2] * 7
    assert oracle.default_bases(93, "gpu") == [31, 3]
    assert oracle.default_bases(480, "gpu") == [5, 3, 2, 2, 2, 2, 2]
    assert oracle.default_bases(640, "gpu") == [5] + [2] * 7
    assert sorted(oracle.default_bases(93, "cpu")) == [3, 31]
    assert oracle.default_bases(97 * 2, "cpu") == [2, 97]


SHAPES = [(3, 1024), (4, 93), (5, 128), (2, 48, 40), (2, 16, 16, 16), (1, 6, 10, 12, 14), (2, 97), (1, 194)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("inverse", [False, True])
def test_random_vs_numpy(oracle, shape, inverse):
    """numpy float64 fftn is the independent second oracle."""
    rng = np.random.default_rng(hash(shape) % 2**32)
    x = rng.standard_normal(shape + (2,)).astype(np.float32)
    axes = tuple(range(1, len(shape)))
    xc = c2(x.astype(np.float64))
    want = np.fft.ifftn(xc, axes=axes) if inverse else np.fft.fftn(xc, axes=axes)
    got32 = c2(oracle.ref_fft(x, inverse=inverse).astype(np.float64))
    got64 = c2(oracle.ref_fft(x, inverse=inverse, out_dtype=np.float64))
    assert np.linalg.norm(got64 - want) / np.linalg.norm(want) < 1e-13
    # stated fp32 tolerance of the reference arithmetic (fp32 theta twiddles)
    assert np.linalg.norm(got32 - want) / np.linalg.norm(want) < 5e-6
    assert np.abs(got32 - want).max() <= 2e-5 * np.abs(want).max()


def test_real_input_full_spectrum(oracle):
    """in_layout (..., 1): stage 0 reads reals, output is the FULL spectrum (_fft.mojo:254-255)."""
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 20, 12, 1)).astype(np.float32)
    want = np.fft.fftn(x[..., 0].astype(np.float64), axes=(1, 2))
    got = c2(oracle.ref_fft(x, out_dtype=np.float64))
    assert got.shape == want.shape
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-13


def test_workers_do_not_change_result(oracle):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((6, 24, 20, 2)).astype(np.float32)
    a = oracle.ref_fft(x, workers=1)
    b = oracle.ref_fft(x, workers=4)
    assert np.array_equal(a, b)


def test_plan_handle_is_reusable_and_writes_every_element(oracle):
    """RefPlan = plan_fft once + fft(out, x, plan=plan) many times (what bench.py times): every call writes the whole
    caller-provided output (pre-filled with NaN like tests.mojo:172-176) and gives the bits of the one-shot call,
    for any worker count, also when calls with different inputs are interleaved on the same plan (calc_buf reuse)."""
    rng = np.random.default_rng(21)
    for shape, kw in [((37, 1024, 2), {}), ((5, 24, 20, 2), {}), ((3, 6, 4, 8, 1), {}), ((11, 93, 2), {"bases": [[31, 3]]}),
                      ((4, 128, 2), {"inverse": True})]:
        x1 = rng.standard_normal(shape).astype(np.float32)
        x2 = rng.standard_normal(shape).astype(np.float32)
        plan = oracle.RefPlan(shape, np.float32, **kw)
        want1, want2 = oracle.ref_fft(x1, **kw), oracle.ref_fft(x2, **kw)
        for workers in (1, 3, 8):
            for x, want in ((x1, want1), (x2, want2), (x1, want1)):
                out = np.full(plan.out_shape, np.nan, np.float32)
                plan.exec(out, x, workers=workers)
                assert np.array_equal(out, want), (shape, workers)
        plan.destroy()
    with pytest.raises(ValueError):
        oracle.RefPlan((4, 100, 2), bases=[[7]])


def test_plan_exec_from_many_threads(oracle):
    """The worker pool is process-wide: concurrent execs of different plans serialise on it and stay correct."""
    import threading
    rng = np.random.default_rng(5)
    xs = [rng.standard_normal((16, 256, 2)).astype(np.float32) for _ in range(4)]
    wants = [oracle.ref_fft(x) for x in xs]
    errs = []

    def work(i):
        plan = oracle.RefPlan(xs[i].shape)
        out = np.empty(plan.out_shape, np.float32)
        for _ in range(5):
            plan.exec(out, xs[i], workers=4)
            if not np.array_equal(out, wants[i]):
                errs.append(i)
        plan.destroy()

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def test_simd_path_is_bit_identical_to_the_scalar_restatement(oracle, tmp_path):
    """The oracle's AVX2 stage loops exist only to make the CPU baseline run at a realistic speed; they must give the
    same bits as the scalar restatement of the reference (same FMA sequence per output point). Build the file once
    more without AVX2 and compare."""
    import ctypes
    import subprocess
    import importlib.util
    src = os.path.join(os.path.dirname(oracle.__file__), "ref_fft.cpp")
    so = str(tmp_path / "libref_scalar.so")
    subprocess.check_call(["g++", "-O2", "-march=x86-64-v3", "-mno-avx2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread",
                           "-shared", "-o", so, src])
    spec = importlib.util.spec_from_file_location("oracle_scalar", oracle.__file__)
    scalar = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(scalar)
    scalar._LIB_PATH, scalar._lib = so, None
    rng = np.random.default_rng(12)
    for shape, bases, inverse in [((9, 1024), None, False), ((9, 1024), None, True), ((3, 640, 480), None, False),
                                  ((5, 1000), [[10, 10, 10]], True), ((4, 1536), [[2, 3]], False), ((2, 4096), [[16]], False),
                                  ((6, 186), [[31, 3, 2]], False), ((7, 128), None, False)]:
        x = rng.standard_normal(shape + (2,)).astype(np.float32)
        assert np.array_equal(oracle.ref_fft(x, bases=bases, inverse=inverse), scalar.ref_fft(x, bases=bases, inverse=inverse)), shape
