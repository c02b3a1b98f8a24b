"""ORACLE — test infrastructure only.

ctypes loader for oracle/libref_fft.so, the CPU restatement of the reference's
radix-n Stockham FFT (see ref_fft.cpp for the reference file:line map). Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libref_fft.so")
_lib = None

_IN_DTYPES = {np.dtype(np.uint8): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}


def build(force=False):
    """Compile libref_fft.so with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "ref_fft.cpp"))):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []) + ["all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        L.ref_ordered_bases.argtypes = [ctypes.c_uint64, u32p, ctypes.c_int, u32p, ctypes.c_int]
        L.ref_ordered_bases.restype = ctypes.c_int
        L.ref_default_bases.argtypes = [ctypes.c_uint64, ctypes.c_int, u32p, ctypes.c_int]
        L.ref_default_bases.restype = ctypes.c_int
        for name in ("ref_fft_exec_f32", "ref_fft_exec_f64"):
            f = getattr(L, name)
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                          ctypes.c_int, ctypes.POINTER(ctypes.c_int64), u32p,
                          ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int]
            f.restype = ctypes.c_int
        L.ref_hardware_threads.restype = ctypes.c_int
        L.ref_plan_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), u32p,
                                      ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
        L.ref_plan_create.restype = ctypes.c_int
        L.ref_plan_exec.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ref_plan_exec.restype = ctypes.c_int
        L.ref_plan_destroy.argtypes = [ctypes.c_void_p]
        L.ref_plan_destroy.restype = None
        _lib = L
    return _lib


def ordered_bases(length, bases):
    """Stage list the reference derives from user bases (_utils.mojo:163-221); None if rejected."""
    arr = (ctypes.c_uint32 * len(bases))(*bases)
    out = (ctypes.c_uint32 * 64)()
    n = lib().ref_ordered_bases(length, arr, len(bases), out, 64)
    return None if n < 0 else [int(out[i]) for i in range(n)]


def default_bases(length, target="cpu"):
    """Default user bases of plan_fft (fft.mojo:49-104) for target 'cpu' or 'gpu'."""
    out = (ctypes.c_uint32 * 64)()
    n = lib().ref_default_bases(length, 1 if target == "gpu" else 0, out, 64)
    return [int(out[i]) for i in range(n)]


def hardware_threads():
    return int(lib().ref_hardware_threads())


def _marshal_bases(bases, ndims):
    if bases is None:
        return None, None
    if len(bases) != ndims:
        raise ValueError("one bases list per axis")
    flat = [b for bl in bases for b in bl]
    return (ctypes.c_uint32 * len(flat))(*flat), (ctypes.c_int32 * len(bases))(*[len(bl) for bl in bases])


class RefPlan:
    """`plan_fft[...](cpu_workers=...)` of the reference's CPU path (fft.mojo:122-157 -> _CPUPlan,
    _ndim_fft_cpu.mojo:28-60): stage lists, twiddle tables and the calc_buf scratch are built once here;
    `exec(out, x)` is `fft(out, x, plan=plan)` (fft.mojo:213-259) and writes into the caller's buffer.
    This is the split the reference's own CPU bench times (fft/bench.mojo:83-90: plan outside the loop)."""

    def __init__(self, in_shape, in_dtype=np.float32, bases=None, inverse=False, out_dtype=np.float32):
        in_shape = tuple(int(v) for v in in_shape)
        if len(in_shape) < 3 or in_shape[-1] not in (1, 2):
            raise ValueError("layout must be (batches, dims..., 1|2)")
        in_dtype, out_dtype = np.dtype(in_dtype), np.dtype(out_dtype)
        if in_dtype not in _IN_DTYPES:
            raise ValueError("unsupported input dtype %s" % in_dtype)
        if out_dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("unsupported output dtype %s" % out_dtype)
        self.in_shape, self.in_dtype, self.out_dtype = in_shape, in_dtype, out_dtype
        self.out_shape = in_shape[:-1] + (2,)
        dims = in_shape[1:-1]
        cdims = (ctypes.c_int64 * len(dims))(*dims)
        flat_p, cnt_p = _marshal_bases(bases, len(dims))
        self._h = ctypes.c_void_p()
        rc = lib().ref_plan_create(ctypes.byref(self._h), int(out_dtype == np.float64), _IN_DTYPES[in_dtype],
                                   in_shape[-1], in_shape[0], len(dims), cdims, flat_p, cnt_p, 1 if inverse else 0)
        if rc != 0:
            self._h = None
            raise ValueError("reference rejects this plan (code %d)" % rc)

    def exec(self, out, x, workers=0):
        if self._h is None:
            raise ValueError("plan destroyed")
        if tuple(x.shape) != self.in_shape or x.dtype != self.in_dtype or not x.flags.c_contiguous:
            raise ValueError("x must be C-contiguous %s %s" % (self.in_shape, self.in_dtype))
        if tuple(out.shape) != self.out_shape or out.dtype != self.out_dtype or not out.flags.c_contiguous:
            raise ValueError("out must be C-contiguous %s %s" % (self.out_shape, self.out_dtype))
        rc = lib().ref_plan_exec(self._h, x.ctypes.data, out.ctypes.data, int(workers))
        if rc != 0:
            raise RuntimeError("ref_plan_exec failed (%d)" % rc)
        return out

    def destroy(self):
        if getattr(self, "_h", None):
            lib().ref_plan_destroy(self._h)
            self._h = None

    __del__ = destroy


def ref_fft(x, bases=None, inverse=False, out_dtype=np.float32, workers=0):
    """Reference-semantics transform of x with layout (batches, d0[, d1...], 1|2).

    Returns an array (batches, d0..., 2) of out_dtype. `bases` is one list per axis
    (None -> the reference's CPU default). Mirrors `fft(output, x, plan=plan_fft[...]())`
    on the CPU path (fft.mojo:213-259). Convenience for the parity tests: builds a plan, pre-fills
    the output with NaN like the reference's tests do (tests.mojo:172-176) and runs it once;
    timing code uses RefPlan so that only exec is inside the timed region.
    """
    x = np.ascontiguousarray(x)
    if x.ndim < 3 or x.shape[-1] not in (1, 2):
        raise ValueError("layout must be (batches, dims..., 1|2)")
    if x.dtype not in _IN_DTYPES:
        raise ValueError("unsupported input dtype %s" % x.dtype)
    plan = RefPlan(x.shape, x.dtype, bases=bases, inverse=inverse, out_dtype=out_dtype)
    out = np.full(plan.out_shape, np.nan, dtype=plan.out_dtype)
    try:
        plan.exec(out, x, workers=workers)
    finally:
        plan.destroy()
    return out
