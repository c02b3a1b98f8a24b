#!/usr/bin/env python3
"""Extract the SASS of the kernels that matter from libb200fft.so into profiles/sass/ (works without a GPU).

    python tools/dump_sass.py            # writes profiles/sass/<name>.sass and profiles/sass/README.md

For each kernel: the full `cuobjdump -sass` listing of that function and an opcode histogram; README.md
holds the histograms side by side plus the mnemonics that evidence the sm_100a features used (UTMALDG /
UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk, SYNCS = mbarrier, FADD2 = packed fp32x2 adds)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hackathon-fft_b200", "lib", "libb200fft.so")
OUT = os.path.join(ROOT, "profiles", "sass")

# (file name, regex on the demangled kernel name)
KERNELS = [
    ("rows1024_32x32_fwd", r"rows_kernel<1024, b200fft::Radices<32, 32>, 8, 256, false, false, false>"),
    ("rows128_16x8_fwd", r"rows_kernel<128, b200fft::Radices<16, 8>, 32, 256, false, false, false>"),
    ("rowsV128_16x8_fwd", r"rows_kernel<128, b200fft::Radices<16, 8>, 32, 256, false, false, true>"),
    ("rows93_31x3_fwd", r"rows_kernel<93, b200fft::Radices<31, 3>, 32, 96, false, false, false>"),
    ("rows480_24x20_fwd", r"rows_kernel<480, b200fft::Radices<24, 20>, 8, 192, false, false, false>"),
    ("cols640_32x20_fwd", r"cols_kernel<640, b200fft::Radices<32, 20>, 16, 320, false, false>"),
    ("cols512_32x16_fwd", r"cols_kernel<512, b200fft::Radices<32, 16>, 16, 256, false, false>"),
    ("cols_tma640_32x20_fwd", r"cols_tma_kernel<640, b200fft::Radices<32, 20>, 16, 320, false>"),
    ("cols_scatter512_32x16_fwd", r"cols_scatter_kernel<512, b200fft::Radices<32, 16>, 16, 256, false>"),
    ("r2c240_16x15", r"rows_r2c_kernel<240, b200fft::Radices<16, 15>, 16, 256>"),
    ("r2c_reg512_32x16", r"rows_r2c_reg_kernel<512, b200fft::Radices<32, 16>, 8, 256>"),
    ("r2c_reg64_8x8", r"rows_r2c_reg_kernel<64, b200fft::Radices<8, 8>, 32, 256>"),
    ("r2c_odd93_31x3", r"rows_r2c_odd_kernel<93, b200fft::Radices<31, 3>, 32, 96>"),
    ("nd_async_64x64x64_plane_fwd", r"nd_async_kernel<256, 2, b200fft::APlane<64, 64, .*?, false>, b200fft::ACols<64, .*?, 64, false>, b200fft::ANone, 2>"),
    # round 2: the kernels VERDICT r1 found without a listing
    ("nd_async_64x64x64_r2cplane_zslot", r"nd_async_kernel<128, 4, b200fft::AR2CPlane<64, 32, .*?, true>, b200fft::ACols<64, .*?, 32, false>, b200fft::ANone, 2>"),
    ("nd_async_128x128x128_rows_fwd", r"nd_async_kernel<256, 2, b200fft::ARows<128, .*?, 32, false, false>, b200fft::ACols<128, .*?, 32, false>, b200fft::ACols<128, .*?, 32, false>, 2>"),
    ("slab_fused512_fwd", r"slab_fused_kernel<512, 512, 512, .*?, false>"),
    ("cols_split_a128_16x8_fwd", r"cols_split_a_kernel<128, b200fft::Radices<16, 8>, 16, 128, false>"),
    ("rows_split_b128_16x8_fwd", r"rows_split_b_kernel<128, b200fft::Radices<16, 8>, 32, 256, false>"),
    ("cols_split_b15_fwd", r"cols_split_b_kernel<15, b200fft::Radices<15>, 128, 128, false>"),
    ("rt_axis_fwd_r32", r"rt_axis_kernel<false, 32>"),
    ("rt_axis_fwd_r16", r"rt_axis_kernel<false, 16>"),
    ("c2r512_32x16", r"rows_c2r_kernel<512, b200fft::Radices<32, 16>, 8, 256>"),
    ("gen_fft_f32", r"gen_fft_kernel<float>"),
    # plane kernels (csrc/plane.cuh)
    ("c2c_plane64x64_fwd", r"c2c_plane_kernel<64, 64, b200fft::Radices<8, 8>, b200fft::Radices<8, 8>, 256, false, false>"),
    ("r2c_plane64x64", r"r2c_plane_kernel<64, 32, b200fft::Radices<8, 8>, b200fft::Radices<8, 4>, 256>"),
    ("c2r_plane64x64", r"c2r_plane_kernel<64, 32, b200fft::Radices<8, 8>, b200fft::Radices<8, 4>, 128>"),
    ("c2c_plane128x128_inplace_fwd", r"c2c_plane_ip_kernel<128, 128, b200fft::Radices<16, 8>, b200fft::Radices<16, 8>, 512, false, false>"),
    ("r2c_plane128x128_inplace", r"r2c_plane_ip_kernel<128, 64, b200fft::Radices<8, 16>, b200fft::Radices<8, 8>, 512>"),
    # one-buffer row kernel (rows_ip_kernel, csrc/fast.cuh)
    ("rows_ip16384_32x32x16_fwd", r"rows_ip_kernel<16384, b200fft::Radices<32, 32, 16>, 1, 512, false, false>"),
    ("rows_ip4096_16x16x16_fwd", r"rows_ip_kernel<4096, b200fft::Radices<16, 16, 16>, 1, 256, false, false>"),
]
MAX_LINES = 6000  # longer listings (the runtime-length kernel unrolls 31 codelets: 25 MB) are cut here; the histogram counts all of it

# plan-time specialised kernels (csrc/jit.cu): compiled here with NVRTC through b200fft_jit_probe (no GPU needed), the
# cubins kept by B200FFT_JIT_DUMP_DIR. (file name, probe arguments)
JIT_KERNELS = [
    ("jit_rows_ip1000_10x10x10_fwd", dict(n=1000)),
    ("jit_rows1296_36x36_fwd", dict(n=1296)),
    ("jit_rows_ip20000_20x40x25_fwd", dict(n=20000)),
    ("jit_cols1000_40x25_fwd", dict(n=1000, inner=1000)),
    ("jit_rows100_10x10_fwd", dict(n=100)),
    ("jit_r2c500_25x20", dict(n=1000, half=1)),
    ("jit_rows1024_16x8x8_f64_fwd", dict(n=1024, in_dtype="float64", out_dtype="float64")),
    ("jit_rows74_37x2_fwd", dict(n=74)),
]


def jit_listings():
    """-> [(file name, demangled kernel name, sass lines)] of the NVRTC-built kernels"""
    import sys
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "hackathon-fft_b200", "python"))
    import b200fft
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        os.environ["B200FFT_JIT_DUMP_DIR"] = tmp
        os.environ["B200FFT_JIT_CACHE"] = "0"   # compile, do not load an earlier process's cubin
        for fname, kw in JIT_KERNELS:
            before = set(os.listdir(tmp))
            rep = b200fft.jit_probe(**kw)
            new = sorted(set(os.listdir(tmp)) - before)
            if not new:
                out.append((fname, None, None))
                continue
            text = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, new[0])], capture_output=True, text=True).stdout
            body, keep = [], False
            for line in text.splitlines():
                if re.match(r"\s*Function : ", line):
                    keep = "kernel" in line
                if keep:
                    body.append(line)
            out.append((fname, rep.split(": ", 1)[1].split(", smem=")[0] + "  [NVRTC]", body))
        del os.environ["B200FFT_JIT_DUMP_DIR"]
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    # split into functions
    funcs = {}
    cur, buf = None, []
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                funcs[cur] = buf
            cur, buf = m.group(1), [line]
        elif cur:
            buf.append(line)
    if cur:
        funcs[cur] = buf
    names = list(funcs)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for fname, pat in KERNELS:
        hit = [n for n, d in zip(names, dem) if re.search(pat, d)]
        if not hit:
            rows.append((fname, None, None))
            continue
        body = funcs[hit[0]]
        with open(os.path.join(OUT, fname + ".sass"), "w") as f:
            f.write("// %s\n" % dem[names.index(hit[0])])
            f.write("\n".join(body[:MAX_LINES]) + "\n")
            if len(body) > MAX_LINES:
                f.write("// ... %d more lines not kept (tools/dump_sass.py regenerates the full listing)\n" % (len(body) - MAX_LINES))
        ops = collections.Counter()
        for line in body:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
        rows.append((fname, dem[names.index(hit[0])], ops))
    for fname, d, body in jit_listings():
        if body is None:
            rows.append((fname, None, None))
            continue
        with open(os.path.join(OUT, fname + ".sass"), "w") as f:
            f.write("// %s\n" % d)
            f.write("\n".join(body) + "\n")
        ops = collections.Counter()
        for line in body:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
        rows.append((fname, d, ops))
    keys = ["FADD", "FADD2", "FMUL", "FFMA", "DADD", "DFMA", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS",
            "ATOMG", "REDG", "MEMBAR", "CCTL", "NANOSLEEP"]
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# SASS listings (sm_100a) of the kernels on the hot path\n\n")
        f.write("Generated by `tools/dump_sass.py` from `hackathon-fft_b200/lib/libb200fft.so` (`cuobjdump -sass`). One file per\n"
                "kernel; the table counts static instructions per opcode. `UTMALDG`/`UTMASTG` = TMA tensor loads/stores\n"
                "(`cp.async.bulk.tensor`), `UBLKCP` = 1-D `cp.async.bulk`, `SYNCS` = mbarrier operations, `FADD2` = packed fp32x2 adds,\n"
                "`SHFL` = warp shuffles (128-bit row variants' two-lane exchange; the R2C kernels' Hermitian unpack in registers),\n"
                "`ATOMG`/`REDG` = the work-fetch and dependency counters of the fused kernel. No `HMMA`/`UTCMMA`: no tensor cores on\n"
                "this path (fp32 complex butterflies are not a dense contraction).\n\n")
        f.write("| kernel | total | " + " | ".join(keys) + " |\n|---|---|" + "---|" * len(keys) + "\n")
        for fname, d, ops in rows:
            if ops is None:
                f.write("| %s | not found in this build |\n" % fname)
                continue
            f.write("| [`%s`](%s.sass) | %d | " % (fname, fname, sum(ops.values())) + " | ".join(str(ops.get(k, 0)) for k in keys) + " |\n")
        f.write("\n")
        for fname, d, ops in rows:
            if d:
                f.write("* `%s`: `%s`\n" % (fname, d))
    print(open(os.path.join(OUT, "README.md")).read())


if __name__ == "__main__":
    main()
