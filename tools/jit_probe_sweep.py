#!/usr/bin/env python3
"""Host-only robustness sweep of the plan-time tier's row planner: every smooth length in a range is planned and compiled
with NVRTC (no GPU needed); prints the lengths the tier declines and any compile failure.
    python tools/jit_probe_sweep.py 2048 22000 [step] [float64]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hackathon-fft_b200", "python")]
os.environ.setdefault("B200FFT_JIT_CACHE", "0")
import b200fft

lo, hi = int(sys.argv[1]), int(sys.argv[2])
step = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dt = sys.argv[4] if len(sys.argv) > 4 else "float32"


def smooth(n):
    for p in (2, 3, 5, 7):
        while n % p == 0:
            n //= p
    return n == 1


kinds, declined, failed = {}, [], []
cands = [n for n in range(lo, hi + 1) if smooth(n)][::step]
t0 = time.time()
for n in cands:
    try:
        rep = b200fft.jit_probe(n, in_dtype=dt, out_dtype=dt)
        kinds[rep.split("<")[0].split("::")[-1]] = kinds.get(rep.split("<")[0].split("::")[-1], 0) + 1
    except b200fft.B200FFTError as e:
        (declined if e.status == 4 else failed).append((n, str(e)[:120]))
print("lengths", len(cands), "kernels", kinds, "declined", [n for n, _ in declined], "failed", failed, "%.0f s" % (time.time() - t0))
