import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_or_skip():
    if not has_gpu():
        pytest.skip("no CUDA device")


@pytest.fixture(scope="session", autouse=True)
def _compat_stubs():
    """The pybind11 stub modules named `rela` / `hanalearn` (hanabi_sad_b200/compat/) that subprocess-based drop-in tests put on
    PYTHONPATH: built once if missing (seconds; they normally exist from __graft_entry__.build())."""
    try:
        from hanabi_sad_b200 import build as hb_build

        hb_build.build_compat()
    except Exception as ex:   # no compiler: the tests that need them will say so themselves
        print("compat stubs not built:", ex)
