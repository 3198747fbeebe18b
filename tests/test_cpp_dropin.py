"""The C++ drop-in headers (include/evaluator.h, include/marching.h) used the way the reference's main.cpp / drawer.cpp
use the originals.  CPU tier: they compile and link against libmcb200.so.  GPU tier: Poly_Data (welded vertex_list +
tri_list) equals the UNMODIFIED reference's, byte for byte (FNV-1a-64 of the raw arrays), on the golden cases."""
import os
import subprocess

import numpy as np
import pytest

from .helpers import load_meta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "marching-cube-for-implicit-surfaces_b200")


def fnv1a64(b):
    h = 0xCBF29CE484222325
    for x in b:
        h = ((h ^ x) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _compile(tmp_path):
    exe = str(tmp_path / "dropin_main")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"), "-o", exe, "-L", PKG, "-lmcb200",
                           "-Wl,-rpath," + PKG])
    return exe


def test_dropin_headers_compile_and_link(mcb, tmp_path):
    assert os.path.exists(_compile(tmp_path))


@pytest.mark.gpu
def test_dropin_poly_data_equals_reference(mcb, golden, tmp_path):
    exe = _compile(tmp_path)
    meta = load_meta(golden)
    names = [n for n, c in meta.items() if not c["cons"] and 0.001 <= c["step"] <= 0.5]
    args = []
    for n in names:
        c = meta[n]
        args += [n, c["eq"], repr(c["step"]), repr(c["scale"][0]), repr(c["scale"][1]), repr(c["scale"][2]), repr(c["iso"])]
    out = subprocess.check_output([exe] + args, text=True)
    lines = {l.split()[0]: l.split()[1:] for l in out.strip().splitlines()}
    for n in names:
        v, t = golden[n + "/vertex_list"], golden[n + "/tri_list"]
        assert lines[n][:2] == [str(len(v)), str(len(t))], (n, lines[n])
        assert lines[n][2] == fnv1a64(v.tobytes()) and lines[n][3] == fnv1a64(t.tobytes()), n
    assert lines["evaluate"] == ["5"]
    assert lines["ctor"] == ["THROW"]
    assert lines["step_rejected"] == ["0", "0"]
