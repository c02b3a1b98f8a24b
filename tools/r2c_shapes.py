#!/usr/bin/env python3
"""Half-spectrum R2C of every BASELINE shape, ours vs cuFFT R2C (same buffers, CUDA events): the rows of bench.py's
`shapes[]` whose name ends in _r2c_half, printed one JSON line each (-> gpurun_out/r2c_shapes.jsonl)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "hackathon-fft_b200", "python")]
import torch

import b200fft
import bench

peak = bench.measured_peak()[0]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r2c_shapes.jsonl"), "w") as f:
    for name, shape, real in bench.SHAPES:
        if real != "half":
            continue
        row = bench.bench_shape(name, shape, real, torch, b200fft, 20, 3, peak)
        line = json.dumps(row)
        print(line)
        f.write(line + "\n")
