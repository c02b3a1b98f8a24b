"""Real multi-rank parity of the slab decomposition: when the box has more than one GPU,
spawn one process per GPU under torchrun and check every exchange mode with the distributed checks of
b200fft.verify. On a one-GPU box the test is skipped (virtual ranks: tests/test_gpu_slab.py; the N-GPU bench
prints the same checks in its slab object)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_modes_on_real_ranks():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("one GPU visible: the spawned multi-rank test needs at least two")
    world = 8 if ngpu >= 8 else 4 if ngpu >= 4 else 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_multirank_slab.py"), "128,256"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("MULTIRANK ")][-1]
    res = json.loads(line[len("MULTIRANK "):])
    assert res["world"] == world and len(res["cases"]) == 12
    for c in res["cases"]:
        assert "error" not in c, c
        assert c["repeat_identical"], c
        assert c["bins"] < 2e-6 and c["delta"] < 2e-6, c          # stated fp32 tolerance
        assert c.get("rel_l2_vs_torch_f64", 0.0) < 2e-6, c
        assert c.get("peer_wait_timeouts", 0) == 0, c
