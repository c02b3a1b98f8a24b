import sys, os, json
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch, b200fft
from bench import time_gpu
st = torch.cuda.current_stream().cuda_stream
for shape in [(100000, 1024), (500000, 128), (100, 64, 64, 64), (100, 640, 480), (20000, 1000)]:
    x = torch.randn(shape + (2,), device="cuda", dtype=torch.float64); out = torch.empty_like(x)
    row = {"shape": list(shape), "dtype": "f64"}
    for name, kw in (("jit", {}), ("generic", {"_test": "generic"})):
        if name == "generic" and x.numel() > 60e6: continue
        plan = b200fft.plan_fft("float64", "float64", x.shape, x.shape, **kw)
        ms = time_gpu(lambda: plan.exec(out, x, st), 2, 5, torch)
        row[name + "_ms"] = round(ms, 4)
        if name == "jit":
            row["gbs"] = round(2 * x.numel() * 8 / ms / 1e6, 1)
            row["plan"] = [l.split(" n=")[0] for l in plan.describe().strip().split("\n")]
        plan.destroy()
    xc = torch.view_as_complex(x)
    t = time_gpu(lambda: torch.fft.fftn(xc, dim=tuple(range(1, xc.dim()))), 2, 5, torch)
    row["cufft_z2z_via_torch_ms"] = round(t, 4)
    print(json.dumps(row), flush=True)
