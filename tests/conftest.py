import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build what the selected tests need (cheap no-op when up to date).  The oracle and the host emulator are
    test infrastructure; the CUDA library is cross-compiled here and only *loaded* in the CPU suite."""
    import cases
    if not os.path.exists(cases.DP_ORACLE):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    if os.path.isdir("/root/reference") and not os.path.exists(cases.DP_REF):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    csrc = os.path.join(ROOT, "stringdecomposer_b200", "csrc")
    if not all(os.path.exists(f) for f in (os.path.join(ROOT, "stringdecomposer_b200", "libsd_b200.so"), cases.DP_CUDA)):
        subprocess.run(["make", "-s", "-j8", "-C", csrc, "all"], check=True)
    # host emulator of the kernels: test infrastructure, built outside the package (make is a no-op when up to date)
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "emu"), "all"], check=True)
    return True


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
