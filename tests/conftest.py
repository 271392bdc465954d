import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def _cuda_available():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_lib():
    """The oracle: unmodified reference sources built by oracle/Makefile."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("oracle/_ref/libtmr_ref.so not built (needs /root/reference)")
    return ref_loader.load()


@pytest.fixture(scope="session")
def emu_lib():
    """Test-only serial emulation of the kernel bodies (tests/emu)."""
    from tmr_b200 import _capi

    path = os.path.join(ROOT, "tests", "emu", "_build", "libtmr_emu.so")
    if not os.path.exists(path):
        import subprocess

        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    return _capi.bind(ctypes.CDLL(path))


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real GPU."""
    import tmr_b200

    lib = tmr_b200.load_library()
    tmr_b200.require_gpu()
    return lib
