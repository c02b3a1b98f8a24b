import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/hackathon-fft_b200/python"]
import numpy as np, torch, b200fft
os.environ["B200FFT_PLANE"] = "1"
rng = np.random.default_rng(3)
for shape, mode, inv in [((5,64,64,64),"c2c",False), ((5,64,64,64),"c2c",True), ((3,64,64,64),"real",False), ((7,64,64,64),"half",False), ((9,64,64),"c2c",False), ((4,64,64),"half",False), ((2,3,64,64,64),"c2c",False)]:
    comps = 2 if mode == "c2c" else 1
    x = rng.standard_normal(shape + (comps,)).astype(np.float32)
    oshape = shape[:-1] + (shape[-1]//2+1, 2) if mode == "half" else shape + (2,)
    plan = b200fft.plan_fft("float32", "float32", x.shape, oshape, inverse=inv, real_mode=b200fft.REAL_HALF if mode=="half" else 0)
    out = torch.full(oshape, float("nan"), device="cuda")
    b200fft.fft(out, torch.from_numpy(x).cuda(), plan=plan); torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64); got = got[...,0] + 1j*got[...,1]
    xd = x.astype(np.float64); xc = xd[...,0] + (1j*xd[...,1] if comps == 2 else 0)
    axes = tuple(range(1, len(shape)))
    want = np.fft.rfftn(xd[...,0], axes=axes) if mode == "half" else (np.fft.ifftn(xc, axes=axes) if inv else np.fft.fftn(xc, axes=axes))
    print(shape, mode, inv, "rel", float(np.linalg.norm(got-want)/np.linalg.norm(want)), plan.describe().strip().split("\n")[0][:70], "| launches", plan.launches)
