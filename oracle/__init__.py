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


def ref_fft(x, bases=None, inverse=False, out_dtype=np.float32, workers=0):
    """Reference-semantics transform of x with layout (batches, d0[, d1...], 1|2).

    Returns an array (batches, d0..., 2) of out_dtype. `bases` is one list per axis
    (None -> the reference's CPU default). Mirrors `fft(output, x, plan=plan_fft[...]())`
    on the CPU path (fft.mojo:213-259).
    """
    x = np.ascontiguousarray(x)
    if x.ndim < 3 or x.shape[-1] not in (1, 2):
        raise ValueError("layout must be (batches, dims..., 1|2)")
    if x.dtype not in _IN_DTYPES:
        raise ValueError("unsupported input dtype %s" % x.dtype)
    dims = x.shape[1:-1]
    out_dtype = np.dtype(out_dtype)
    out = np.full(x.shape[:-1] + (2,), np.nan, dtype=out_dtype)
    cdims = (ctypes.c_int64 * len(dims))(*dims)
    if bases is None:
        flat_p, cnt_p = None, None
    else:
        if len(bases) != len(dims):
            raise ValueError("one bases list per axis")
        flat = [b for bl in bases for b in bl]
        flat_p = (ctypes.c_uint32 * len(flat))(*flat)
        cnt_p = (ctypes.c_int32 * len(bases))(*[len(bl) for bl in bases])
    fn = lib().ref_fft_exec_f32 if out_dtype == np.float32 else lib().ref_fft_exec_f64
    rc = fn(x.ctypes.data, _IN_DTYPES[x.dtype], x.shape[-1], out.ctypes.data, x.shape[0], len(dims),
            cdims, flat_p, cnt_p, 1 if inverse else 0, int(workers))
    if rc != 0:
        raise ValueError("reference rejects this plan (code %d)" % rc)
    return out
