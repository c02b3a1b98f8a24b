"""Randomised parity sweep (fixed seed): random ranks, axis lengths (primes, prime powers, composites), user
base lists drawn from the factorisation (incl. composite bases, any order, under-specified lists), direction,
real / complex input and dtypes — the GPU path through the C ABI against numpy float64 and, for fp32, the oracle."""
import numpy as np
import pytest

import b200fft

pytestmark = pytest.mark.gpu

LENGTHS = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 17, 20, 21, 24, 25, 27, 30, 31, 32, 35, 36, 48, 49, 60, 64, 93, 96,
           100, 121, 125, 128, 160, 169, 240, 243, 256, 343, 480, 512, 640, 1000, 1024]


def _factor(n):
    f, p = [], 2
    while n > 1:
        while n % p == 0:
            f.append(p)
            n //= p
        p += 1
    return f


def _random_bases(rng, n):
    """A base list the reference accepts: prime factors, randomly merged into composites, shuffled; sometimes only
    the distinct primes (under-specified: the planner repeats them, _utils.mojo:163-183), sometimes None."""
    mode = rng.integers(0, 4)
    if mode == 0:
        return None
    f = _factor(n)
    if mode == 1:
        return sorted(set(f))
    rng.shuffle(f)
    out = []
    while f:
        k = int(rng.integers(1, 3))
        prod = 1
        for _ in range(min(k, len(f))):
            prod *= f.pop()
        out.append(prod)
    return out


def _cases(count, seed):
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < count:
        rank = int(rng.integers(1, 4))
        cap = {1: 1024, 2: 256, 3: 64}[rank]
        dims = [int(rng.choice([v for v in LENGTHS if v <= cap])) for _ in range(rank)]
        batch = int(rng.integers(1, 6))
        bases = [_random_bases(rng, d) for d in dims]
        if any(b is None for b in bases):
            bases = None
        inverse = bool(rng.integers(0, 2))
        real = bool(rng.integers(0, 3) == 0) and not inverse
        in_dtype = str(rng.choice(["float32", "float32", "float64", "uint8"])) if not inverse else "float32"
        out_dtype = "float64" if in_dtype == "float64" or rng.integers(0, 5) == 0 else "float32"
        cases.append((batch, tuple(dims), bases, inverse, real, in_dtype, out_dtype))
    return cases


@pytest.mark.parametrize("case", _cases(80, 2026), ids=lambda c: "b%d_%s_%s%s%s_%s" % (
    c[0], "x".join(map(str, c[1])), "inv" if c[3] else "fwd", "_real" if c[4] else "", "_ub" if c[2] else "", c[5][0] + c[6][-2:]))
def test_random_configuration(case, oracle):
    import torch
    batch, dims, bases, inverse, real, in_dtype, out_dtype = case
    rng = np.random.default_rng(abs(hash(case[:2])) % (2 ** 31))
    comps = 1 if real else 2
    shape = (batch,) + dims + (comps,)
    if in_dtype == "uint8":
        x = rng.integers(0, 256, size=shape).astype(np.uint8)
    else:
        x = rng.standard_normal(shape).astype(in_dtype)
    plan = b200fft.plan_fft(in_dtype, out_dtype, shape, (batch,) + dims + (2,), bases=bases, inverse=inverse)
    for a, d in enumerate(dims):                      # the plan's stage list multiplies to the axis length
        assert int(np.prod(plan.bases(a))) == d
    xt = torch.from_numpy(x).cuda()
    out = torch.full((batch,) + dims + (2,), float("nan"), device="cuda",
                     dtype=torch.float64 if out_dtype == "float64" else torch.float32)
    b200fft.fft(out, xt, plan=plan)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    got = got[..., 0] + 1j * got[..., 1]
    xd = x.astype(np.float64)
    xc = xd[..., 0] if real else xd[..., 0] + 1j * xd[..., 1]
    axes = tuple(range(1, 1 + len(dims)))
    want = np.fft.ifftn(xc, axes=axes) if inverse else np.fft.fftn(xc, axes=axes)
    scale = np.linalg.norm(want)
    tol = 1e-12 if out_dtype == "float64" else 2e-6 * max(1.0, np.sqrt(len(dims)))
    assert np.isfinite(got).all()
    assert np.linalg.norm(got - want) <= tol * scale, (case, plan.describe())
    if out_dtype == "float32" and in_dtype != "float64":
        ref = oracle.ref_fft(x, bases=bases, inverse=inverse).astype(np.float64)
        ref = ref[..., 0] + 1j * ref[..., 1]
        assert np.linalg.norm(got - ref) <= 5e-6 * max(1.0, np.sqrt(len(dims))) * scale
    plan.destroy()
