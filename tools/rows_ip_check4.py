#!/usr/bin/env python3
"""Three-stage rows below 4096 points on the plan-time tier: one-buffer kernel (B200FFT_ROWS_INPLACE_MIN=8192) vs the default
two-buffer geometry vs cuFFT. One process per setting (the threshold is read once)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
LENGTHS = tuple(int(v) for v in os.environ.get("LENGTHS", "1536,2000,2187,2500,3000,3072,3125,3600,4000").split(","))
if len(sys.argv) > 1:
    import torch
    import b200fft
    from bench import CuFFT, time_gpu
    st = torch.cuda.current_stream().cuda_stream
    for n in LENGTHS:
        for batch in (100, (200 << 20) // (8 * n)):
            x = torch.randn(batch, n, 2, device="cuda"); out = torch.empty_like(x)
            plan = b200fft.plan_fft("float32", "float32", x.shape, x.shape)
            ms = time_gpu(lambda: plan.exec(out, x, st), 5, 30, torch)
            want = torch.fft.fft(torch.view_as_complex(x[0].double().contiguous()))
            got = torch.view_as_complex(out[0].double().contiguous())
            row = {"shape": [batch, n], "ms": round(ms, 5), "rel": float((got - want).norm() / want.norm()),
                   "plan": plan.describe().strip().split(" user stages")[0].replace("axis 0: ", "")[:90]}
            plan.destroy()
            if sys.argv[1] == "cufft":
                cf = CuFFT((batch, n)); row["cufft_ms"] = round(time_gpu(lambda: cf.exec(x, out, st), 5, 30, torch), 5); cf.destroy()
            print(json.dumps(row), flush=True)
else:
    ARMS = (("default", {}), ("inplace", {"B200FFT_ROWS_INPLACE_MIN": "8192"}))
    if os.environ.get("ARMS") == "onoff":
        ARMS = (("default", {"B200FFT_ROWS_INPLACE": "0"}), ("inplace", {}))
    if os.environ.get("ARMS") == "narrow":
        ARMS = (("default", {}), ("inplace", {"B200FFT_JIT_MAX_RADIX": "32", "B200FFT_ROWS_INPLACE_MIN": "4096"}))
    for tag, env in ARMS:
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "cufft" if tag == "default" else "x"], env=e, capture_output=True, text=True)
        for l in r.stdout.splitlines():
            print(json.dumps({"arm": tag, **json.loads(l)}), flush=True)
        sys.stderr.write(r.stderr[-2000:])
