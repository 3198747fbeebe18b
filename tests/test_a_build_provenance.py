"""-m gpu, first file of the tier: the libmcb200.so this process loaded was compiled from the sources in this tree
(mcb_build_stamp() = sha256 over csrc/ + include/mcb.h, written into the library by build.py at build time), it is the
real device library (not the host emulation of tests/emu), and it finds a CUDA device."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_loaded_library_is_the_nvcc_build_of_these_sources(mcb):
    if os.environ.get("MCB_TEST_EMU") == "1":
        pytest.skip("developer run against the host emulation")
    spec = importlib.util.spec_from_file_location("mcb_build", os.path.join(ROOT, "marching-cube-for-implicit-surfaces_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert os.path.basename(mcb.LIB_PATH) == "libmcb200.so"
    assert mcb.lib.mcb_build_stamp().decode() == b.source_stamp(), "libmcb200.so is stale: rebuild with __graft_entry__.build()"
    with open("/proc/self/maps") as f:
        maps = f.read()
    assert "libmcb200.so" in maps and "libmcb200_emu.so" not in maps
    ctx = mcb.Context(0)   # MCB_E_NODEVICE would raise: there is no CPU path behind the ABI
    ctx.close()
