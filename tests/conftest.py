"""Test tiers:
  -m "not gpu": oracle vs golden vectors, host logic (tokenizer, lowering, grid, slabs), C-ABI symbol check — CPU only.
  -m gpu      : parity of the CUDA path (through the C ABI) against the oracle / golden vectors — needs a B200.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def mcb():
    if os.environ.get("MCB_TEST_EMU") == "1":
        # developer switch: run the `-m gpu` tests' small cases against the kernels' source executed on the host
        # (tests/emu).  Never set by the driver; the GPU tier always goes through libmcb200.so on a real device.
        from tests import emu
        return emu.load()
    return importlib.import_module("marching-cube-for-implicit-surfaces_b200")


@pytest.fixture(scope="session")
def mcb_emu():
    """The package's ctypes mirror bound to tests/emu/_build/libmcb200_emu.so (device code executed on the host)."""
    from tests import emu
    return emu.load()


@pytest.fixture(scope="session")
def refbind():
    from oracle import refbind as R
    if not R.available():
        pytest.skip("oracle/_ref/libmcref.so not built (needs /root/reference)")
    return R


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "cases.npz")
    return np.load(path, allow_pickle=False)
