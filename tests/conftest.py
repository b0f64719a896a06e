import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mls-mpm-godot_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """libmpm_b200.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    import mpm_b200
    if not os.path.exists(mpm_b200.LIB_PATH):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "mls-mpm-godot_b200"), "-j8", "-s"])
    return mpm_b200.load()
