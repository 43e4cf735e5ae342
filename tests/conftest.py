import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle and the CUDA library are built in-tree once per session (nvcc cross-compiles
    without a GPU).  If the prebuilt files are present (GPU box) make is a no-op."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libhtm_oracle.so"])
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "hypotremormcmc_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "drivers")])


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
