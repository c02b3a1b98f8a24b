import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hackathon-fft_b200", "python")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


# The run-time compiled kernels are cached on disk under ~/.cache/b200fft by default (csrc/jit.cu). The test suite must not
# write outside the repository: no disk cache here (tests/test_host.py::test_jit_disk_cache points it at a tmp_path).
os.environ.setdefault("B200FFT_JIT_CACHE", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (test infrastructure, see oracle/)."""
    import oracle as _oracle
    _oracle.build()
    return _oracle
