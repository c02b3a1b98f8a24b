"""L2-resident pass groups (csrc/api.cu: plan_pass_group): the innermost axes' passes run chunk by chunk so the
intermediate stays in L2. Same kernels, same arithmetic: the result must be BIT-identical to the ungrouped plan, for
complex, real-full and half-spectrum inputs, forward and inverse, in place, and through exec_host's batch chunks."""
import os

import numpy as np
import pytest

import b200fft

pytestmark = pytest.mark.gpu


def _plan(shape_in, shape_out, chunk_mb, **kw):
    old = os.environ.get("B200FFT_PASS_CHUNK_MB")
    os.environ["B200FFT_PASS_CHUNK_MB"] = str(chunk_mb)
    try:
        return b200fft.plan_fft("float32", "float32", shape_in, shape_out, flags=b200fft.FLAG_NO_FUSED, **kw)
    finally:
        if old is None:
            del os.environ["B200FFT_PASS_CHUNK_MB"]
        else:
            os.environ["B200FFT_PASS_CHUNK_MB"] = old


CASES = [
    # (batch, dims), comps, half, inverse
    (((40, 96, 80)), 2, False, False),      # 2-D: group = both axes, chunked over the batch
    (((40, 96, 80)), 2, False, True),
    (((37, 96, 80)), 1, False, False),      # real input, full spectrum; batch not a multiple of the chunk
    (((80, 96, 80)), 1, True, False),       # half spectrum R2C: rows pass + strided pass over 41 bins
    (((2, 64, 64, 64)), 2, False, False),   # 3-D: a 2 MB volume does not fit a 1 MB chunk -> group = (y, x) per z plane
    (((6, 32, 48, 40)), 2, False, True),
    (((2, 64, 64, 64)), 1, True, False),
    (((2, 6, 10, 16, 18, 20)), 2, False, False),   # 5-D: group = innermost axes only
    (((12, 100, 243)), 2, False, False),    # runtime-length tier passes
]


@pytest.mark.parametrize("shape,comps,half,inverse", CASES)
def test_grouped_plan_is_bit_identical(shape, comps, half, inverse):
    import torch
    g = torch.Generator(device="cuda").manual_seed(17)
    in_shape = tuple(shape) + (comps,)
    out_shape = tuple(shape[:-1]) + (shape[-1] // 2 + 1, 2) if half else tuple(shape) + (2,)
    x = torch.randn(in_shape, generator=g, device="cuda")
    kw = dict(inverse=inverse, real_mode=b200fft.REAL_HALF if half else b200fft.REAL_FULL)
    plain = _plan(in_shape, out_shape, 0, **kw)
    grouped = _plan(in_shape, out_shape, 1, **kw)
    assert "L2-resident group" not in plain.describe()
    assert "L2-resident group" in grouped.describe(), grouped.describe()
    assert grouped.launches > plain.launches
    a = torch.full(out_shape, float("nan"), device="cuda")
    b = torch.full(out_shape, float("nan"), device="cuda")
    keep = x.clone()
    b200fft.fft(a, x, plan=plain)
    b200fft.fft(b, x, plan=grouped)
    torch.cuda.synchronize()
    assert torch.equal(x, keep)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    if comps == 2 and not half:                         # in place
        c = x.clone()
        b200fft.fft(c, c, plan=grouped)
        torch.cuda.synchronize()
        assert torch.equal(c, a)
    # host buffers: exec_host feeds the plan chunks of the batch; each chunk is grouped on its own
    h_out = np.empty(out_shape, np.float32)
    grouped.exec_host(h_out, x.cpu().numpy())
    assert np.array_equal(h_out, a.cpu().numpy())
    # and against float64
    xd = x.double().cpu().numpy()
    xc = xd[..., 0] + (1j * xd[..., 1] if comps == 2 else 0)
    axes = tuple(range(1, len(shape)))
    want = np.fft.rfftn(xc.real, axes=axes) if half else np.fft.ifftn(xc, axes=axes) if inverse else np.fft.fftn(xc, axes=axes)
    got = a.double().cpu().numpy()
    got = got[..., 0] + 1j * got[..., 1]
    assert np.linalg.norm(got - want) <= 2e-6 * np.linalg.norm(want)
    plain.destroy()
    grouped.destroy()


def test_grouping_is_opt_in():
    """Measured on B200 (profiles/r2_chunk_sweep.jsonl): every chunk size LOSES to whole-array passes (2-D 100 x 640 x 480:
    0.174 ms ungrouped, 0.194 ms at 64 MB chunks, 0.267 ms at 24 MB; 512^3: 1.03 vs 1.18 ms) because each extra launch costs
    ~4-5 us of ramp / tail and an L2-resident pass runs barely faster than an HBM-resident one (~8 vs ~6.5 TB/s). The
    mechanism therefore stays off unless B200FFT_PASS_CHUNK_MB is set."""
    assert "B200FFT_PASS_CHUNK_MB" not in os.environ
    d = b200fft.plan_fft("float32", "float32", (100, 640, 480, 2), (100, 640, 480, 2)).describe()
    assert "L2-resident group" not in d, d
    d = _plan((100, 640, 480, 2), (100, 640, 480, 2), 24).describe()
    assert "L2-resident group: the first 2 passes" in d, d
    d = _plan((1, 512, 512, 512, 2), (1, 512, 512, 512, 2), 24).describe()
    assert "L2-resident group: the first 2 passes" in d, d
