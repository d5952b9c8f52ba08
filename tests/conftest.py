import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    """The product package; lib/libbrt.so is (re)built by its Makefile when sources are newer (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "hardware-ray-tracer_b200", "csrc")])
    return importlib.import_module("hardware-ray-tracer_b200")


@pytest.fixture(scope="session")
def orc_mod():
    """oracle/binding.py with liboracle.so built (test infrastructure only)."""
    from oracle import binding
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    binding.load()
    return binding


@pytest.fixture(scope="session")
def emu_lib():
    """tests/emu/libbrt_emu.so: the product's kernel bodies compiled as sequential host loops (test only)."""
    import ctypes
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    return ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libbrt_emu.so"))
