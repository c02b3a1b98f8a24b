"""The Mojo host package cannot be compiled in this image (no Mojo toolchain), so guard it statically: every C symbol it
binds with external_call must be exported by the built library, its descriptor struct must list the fields of
`struct b200fft_desc` in the header's order with matching widths, and the wrappers must keep the reference's entry-point
names and parameter names (fft/fft/fft.mojo:160-210, 262-323)."""
import ctypes
import os
import re

import b200fft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOP = os.path.join(ROOT, "hackathon-fft_b200", "mojo", "fft")   # package `fft` (the reference's top-level fft/ directory)
MOJO = os.path.join(TOP, "fft")                                  # sub-package `fft.fft`


def _read(name):
    with open(os.path.join(MOJO, name)) as f:
        return f.read()


def test_external_calls_bind_exported_symbols():
    lib = ctypes.CDLL(b200fft.lib_path())
    syms = set(re.findall(r'external_call\["(\w+)"', _read("_ffi.mojo")))
    assert {"b200fft_plan_create", "b200fft_exec", "b200fft_plan_destroy", "b200fft_last_error"} <= syms
    for s in syms:
        assert hasattr(lib, s), "the Mojo binding calls %s, which the library does not export" % s


def test_descriptor_struct_matches_the_header():
    header = open(os.path.join(ROOT, "include", "b200fft.h")).read()
    body = header[header.index("typedef struct b200fft_desc {"):header.index("} b200fft_desc;")]
    c_fields = re.findall(r"^\s*(?:const\s+)?(\w+)\*?\s+(\w+)(?:\[\w+\])?;", body, flags=re.M)
    c_names = [n for _, n in c_fields]
    mojo = _read("_ffi.mojo")
    sbody = mojo[mojo.index("struct B200Desc"):]
    sbody = sbody[:sbody.index("\n\n\n")]
    m_fields = re.findall(r"^\s+var (\w+): ([\w\[\], ]+)$", sbody, flags=re.M)
    m_names = [n for n, _ in m_fields if not n.startswith("_pad")]
    assert m_names == c_names, (m_names, c_names)
    width = {"Int32": 4, "UInt32": 4, "Int64": 8}
    sizes = []
    for n, t in m_fields:
        if t.startswith("InlineArray[Int64"):
            sizes.append(8 * 8)
        elif t.startswith("UnsafePointer"):
            sizes.append(8)
        else:
            sizes.append(width[t])
    assert sum(sizes) == ctypes.sizeof(b200fft._Desc) == 128      # explicit _pad0 after `rank`, no hidden padding
    # same offsets as the ctypes struct the tests exercise
    off, offsets = 0, {}
    for (n, _), sz in zip(m_fields, sizes):
        offsets[n] = off
        off += sz
    for n in c_names:
        assert offsets[n] == getattr(b200fft._Desc, n).offset, n


def _defs(name):
    """top-level names a module defines: def / struct / comptime"""
    return set(re.findall(r"^(?:def|struct|comptime) (\w+)", _read(name), flags=re.M))


def test_reference_drivers_import_paths_resolve():
    """The import lines of the reference's own drivers must resolve against this package unchanged:
         fft/bench.mojo:16   from fft.fft.fft import fft, plan_fft
         fft/tests.mojo:12   from fft.fft.fft import fft, plan_fft, _estimate_best_bases_nd
         fft/tests.mojo:13   from fft.fft._ndim_fft_gpu import _run_gpu_nd_fft, _GPUTest
         fft/fft/__init__.mojo:1   from .fft import fft
    `fft` = the directory with __init__.mojo that holds the drivers, `fft.fft` its sub-package, `fft.fft.fft` the module."""
    assert os.path.isfile(os.path.join(TOP, "__init__.mojo")) and os.path.isfile(os.path.join(MOJO, "__init__.mojo"))
    for driver in ("bench.mojo", "tests.mojo", "profile.mojo"):
        assert os.path.isfile(os.path.join(TOP, driver)), driver
    assert {"fft", "plan_fft", "_estimate_best_bases_nd", "_estimate_best_bases", "_check_layout_conditions_nd"} <= _defs("fft.mojo")
    assert {"_run_gpu_nd_fft", "_GPUTest", "_GPUPlan"} <= _defs("_ndim_fft_gpu.mojo")
    assert {"_run_cpu_nd_fft", "_CPUPlan"} <= _defs("_ndim_fft_cpu.mojo")
    assert re.search(r"^from \.fft import fft$", _read("__init__.mojo"), flags=re.M)
    # the four members and the `v` field of the reference's _GPUTest (_ndim_fft_gpu.mojo:453-459)
    gt = _read("_ndim_fft_gpu.mojo")
    body = gt[gt.index("struct _GPUTest"):gt.index("def _test_flags")]
    for k, name in enumerate(("BLOCK", "WARP", "DEVICE_WIDE", "CLUSTER")):
        assert "comptime %s = Self(%d)" % (name, k) in body
    assert "var v: UInt" in body
    # every module an in-package import names exists, and every imported name is defined there
    for mod in ("fft.mojo", "_ndim_fft_gpu.mojo", "_ndim_fft_cpu.mojo", "_utils.mojo"):
        src = _read(mod)
        for target, names in re.findall(r"^from \.(\w+) import \(?([^)]*?)\)?$", src, flags=re.M | re.S):
            have = _defs(target + ".mojo")
            for n in [v.strip() for v in names.replace("\n", " ").split(",") if v.strip()]:
                assert n in have, "%s imports %s from .%s, which does not define it" % (mod, n, target)
    # our own drivers use the reference's import lines
    for driver in ("bench.mojo", "tests.mojo", "profile.mojo"):
        src = open(os.path.join(TOP, driver)).read()
        assert "from fft.fft.fft import fft, plan_fft" in src, driver
        for modpath, names in re.findall(r"^from fft\.fft\.(\w+) import (.+)$", src, flags=re.M):
            for n in [v.strip() for v in names.split(",")]:
                assert n in _defs(modpath + ".mojo"), (driver, modpath, n)


def test_wrappers_keep_the_reference_entry_points():
    """plan_fft x2 and fft x2 with the reference's parameter names (fft/fft/fft.mojo:122-132, 160-176, 213-233, 262-296),
    `_test: Optional[_GPUTest]` (:176), host-tensor overloads present (`cpu_workers`), and the GPU launch goes through
    the context's stream or is bracketed by synchronisation (never a bare NULL-stream launch)."""
    src = _read("fft.mojo")
    plans = [m.start() for m in re.finditer(r"^def plan_fft\[", src, flags=re.M)]
    ffts = [m.start() for m in re.finditer(r"^def fft\[", src, flags=re.M)]
    assert len(plans) == 2 and len(ffts) == 2
    cpu_plan, gpu_plan = src[plans[0]:plans[1]], src[plans[1]:ffts[0]]
    for param in ("in_dtype: DType", "out_dtype: DType", "in_layout: Layout", "out_layout: Layout", "bases: List[List[UInt]]",
                  "inverse: Bool"):
        assert param in cpu_plan and param in gpu_plan, param
    assert "cpu_workers: Optional[UInt] = None" in cpu_plan and "_CPUPlan[" in cpu_plan
    assert '_estimate_best_bases_nd[\n        in_layout, out_layout, "cpu"\n    ]()' in cpu_plan
    for param in ("runtime_twfs: Bool = True", "max_cluster_size: UInt = 8", "_test: Optional[_GPUTest] = None", "ctx: DeviceContext",
                  "_GPUPlan["):
        assert param in gpu_plan, param
    assert '_estimate_best_bases_nd[\n        in_layout, out_layout, "gpu"\n    ]()' in gpu_plan
    cpu_fft, gpu_fft = src[ffts[0]:ffts[1]], src[ffts[1]:src.index("# ---- half-spectrum")]
    for param in ("output: LayoutTensor[out_dtype, out_layout, out_origin, ...]", "x: LayoutTensor[in_dtype, in_layout, in_origin, ...]",
                  "plan: _CPUPlan[out_dtype, out_layout, inverse, bases]", "cpu_workers: Optional[UInt] = None"):
        assert param in cpu_fft, param
    for param in ("output: LayoutTensor[out_dtype, out_layout, out_origin]", "x: LayoutTensor[in_dtype, in_layout, in_origin]",
                  "ctx: DeviceContext", "plan: _GPUPlan["):
        assert param in gpu_fft, param
    for f in (cpu_fft, gpu_fft):
        assert "_check_layout_conditions_nd[in_layout, out_layout]()" in f and "raises" in f
    gpu = _read("_ndim_fft_gpu.mojo")
    launch = gpu[gpu.index("def _launch("):gpu.index("def _run_gpu_nd_fft[")]
    assert "CUDA(ctx.stream())" in launch                      # asynchronous on the context's own stream, or ...
    assert launch.index("ctx.synchronize()") < launch.index("UnsafePointer[NoneType]())") < launch.index("stream_synchronize(")
    assert "exec_host(" in _read("_ndim_fft_cpu.mojo")


def test_mojo_default_bases_rule_matches_the_library():
    """`_estimate_best_bases` is re-stated in Mojo for compile time (fft.mojo) and in C++ for run time (planner.cpp).
    Neither can be executed here from Mojo, so transliterate the Mojo text's rule (candidate order, greedy take,
    reversal, GPU window, prime list) and compare with the library over many lengths."""
    src = _read("fft.mojo")
    primes = [int(v) for v in re.search(r"var primes: List\[UInt\] = \[([\d,\s]+)\]", src).group(1).replace("\n", " ").split(",")]
    assert primes == [97, 89, 83, 79, 73, 71, 67, 61, 59, 53, 47, 43, 41, 37, 31, 29, 23, 19, 17, 13, 11, 7, 5, 3, 2]
    assert "comptime max_radix = 32" in src and "comptime block = 1024" in src
    assert "range(max((length + block - 1) // block, 2), max_radix + 1)" in src and "picked.reverse()" in src
    utils = _read("_utils.mojo")
    assert "while rest >= base and rest % base == 0:" in utils and "return twos // lg" in utils

    def times(length, base):          # _utils.mojo: _times_divisible_by
        if base & (base - 1) == 0:
            twos = 0
            while length > 0 and length % 2 == 0:
                length //= 2
                twos += 1
            return twos // (base.bit_length() - 1)
        t = 0
        while length >= base and length % base == 0:
            length //= base
            t += 1
        return t

    def greedy(length, cands):
        picked, processed = [], 1
        for c in cands:
            for _ in range(times(length // processed, c)):
                picked.append(c)
                processed *= c
            if processed == length:
                break
        return picked[::-1]

    def estimate(length, gpu):
        if gpu and length // 32 <= 1024:
            got = greedy(length, list(range(max((length + 1023) // 1024, 2), 33)))
            prod = 1
            for b in got:
                prod *= b
            if prod == length:
                return got
        return greedy(length, primes)

    for length in list(range(2, 700)) + [93, 128, 480, 640, 1000, 1024, 1080, 2160, 4096, 4320, 16384, 32768, 65536, 97 * 89, 3 * 31 * 37]:
        for gpu in (True, False):
            assert estimate(length, gpu) == b200fft.default_bases(length, "gpu" if gpu else "cpu"), (length, gpu)


def test_mojo_test_driver_lists_the_reference_cases_and_uses_defined_names():
    """hackathon-fft_b200/mojo/tests.mojo (the Mojo-side counterpart of fft/tests.mojo): its 1-D list is exactly the
    reference's 56 (length, bases) pairs in order (the same list the Python GPU suite reads from the golden fixture),
    and everything it imports from the package is defined there."""
    import json
    src = open(os.path.join(TOP, "tests.mojo")).read()
    body = src[src.index("def test_fft_1d_gpu()"):src.index("# beyond the reference's list")]
    listed = [(int(n), [int(v) for v in b.split(",")]) for n, b in re.findall(r"_test_1d\[(\d+), \[([\d, ]+)\]\]\(\)", body)]
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        cases = [(c["length"], c["bases"]) for c in json.load(f)["cases_1d"]]
    assert listed == cases
    assert "_GPUTest.BLOCK, _GPUTest.WARP, _GPUTest.DEVICE_WIDE, _GPUTest.CLUSTER" in src      # every tier, tests.mojo:398-417
    for entry in ("test_fft_host_tensors", "test_fft_1d_gpu", "test_fft_2d_gpu", "test_fft_3d_gpu", "test_rfft_half_gpu"):
        assert re.search(r"^    %s\(\)$" % entry, src[src.index("def main()"):], flags=re.M), entry
