import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import ctypes
        cudart = None
        for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                cudart = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if cudart is None:
            import torch
            return torch.cuda.is_available()
        n = ctypes.c_int(0)
        return cudart.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


HAVE_GPU = _have_gpu()


@pytest.fixture(scope="session")
def ctx():
    """Device context of the CUDA path; -m gpu tests FAIL (not skip) when it cannot be made."""
    from hubbardtn_b200 import device
    c = device.Context(0)
    yield c
    c.close()
