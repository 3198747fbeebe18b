"""TEST INFRASTRUCTURE ONLY: runs the product's device code on the host (see cuda_runtime.h in this directory)."""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG_DIR = os.path.join(ROOT, "marching-cube-for-implicit-surfaces_b200")
_cached = None


def load():
    """The package's ctypes mirror (its __init__.py, unchanged) bound to the emulated library instead of libmcb200.so.
    Returns a module object with the same API as the package.  Only tests call this."""
    global _cached
    if _cached is not None:
        return _cached
    spec = importlib.util.spec_from_file_location("mcb_emu_build", os.path.join(HERE, "build_emu.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    lib = b.build()
    init = os.path.join(PKG_DIR, "__init__.py")
    src = open(init).read()
    needle = 'LIB_PATH = os.path.join(HERE, "libmcb200.so")'
    assert needle in src
    src = src.replace(needle, "LIB_PATH = %r" % lib)
    mod = types.ModuleType("mcb_emulated")
    mod.__file__ = init
    mod.__package__ = None
    sys.modules["mcb_emulated"] = mod
    exec(compile(src, init, "exec"), mod.__dict__)
    _cached = mod
    return mod
