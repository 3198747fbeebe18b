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
                           "-Wl,-rpath," + PKG, "-pthread"])
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


# the UNMODIFIED reference at step 2/256 (M=257), scale 1, iso 0:
# welded vertices, triangles, FNV-1a-64 of the raw vertex_list bytes, FNV-1a-64 of the raw tri_list bytes
REF_257 = {
    "eq1": (66306, 263682, "12bc8487aed915e7", "a41206d677ea9131"),
    "eq2": (66820, 266244, "b0461e91f16b54c7", "0293c81f39a3d673"),
    "eq3": (77069, 153868, "fd509bf878cc50dc", "38ee49b70a71ffa4"),
    "eq4": (120533, 240532, "a55f12cfe1353571", "08c839f8da495f9f"),
    "eq5": (185710, 369792, "a22332f80d29cc58", "0d5f1da974b8a315"),
    "eq6": (93081, 186060, "c42d1b6129171513", "cdfa99f20113e545"),
    "eq7": (39435, 78928, "62f319027c4eef7a", "19506a69e993cfff"),
    "eq8": (185340, 370672, "970f97ce2d48298f", "724a4400f6dca358"),
    "sphere": (151398, 302792, "d2b7d15ee850c4f7", "1954e5f6824dc6e8"),
    "torus": (115880, 232128, "5e3bb1e8040fe4d7", "9f132f6bbfdb0626"),
    "gyr78": (1206817, 2396052, "03cde882e0149ecd", "72b5f24c98197d6a"),
}


@pytest.mark.gpu
def test_dropin_poly_data_at_256_cubed_hashes(mcb, tmp_path):
    """BASELINE.json configs[1]: every example equation (plus sphere, torus and the polynomial gyroid) at 256^3 through
    the C++ drop-in; the welded Poly_Data must hash to what the unmodified reference produced (17 M cubes each)."""
    from oracle.refbind import EXAMPLE_EQUATIONS, SPHERE, TORUS, GYR78
    exe = _compile(tmp_path)
    eqs = {"eq%d" % n: EXAMPLE_EQUATIONS[n] for n in range(1, 9)}
    eqs.update(sphere=SPHERE, torus=TORUS, gyr78=GYR78)
    args = []
    for n, e in eqs.items():
        args += [n, e, repr(2.0 / 256), "1", "1", "1", "0"]
    # Appendix B's hashes use the offset basis 1469598103934665603 (the standard one with its last digit missing)
    out = subprocess.check_output([exe] + args, text=True, env=dict(os.environ, FNV_BASIS="1469598103934665603"))
    lines = {l.split()[0]: l.split()[1:] for l in out.strip().splitlines()}
    # Documented deviation (weld): the reference's comparator is not a strict weak ordering, and on three of
    # these fields its std::set fails to find a vertex that the comparator itself calls equal (an unrelated vertex whose
    # x lies within 1e-6 of one point but not of the other diverts the tree descent), so the reference keeps a few
    # duplicates: 18 of 93 081 vertices (eq6), 1 of 39 435 (eq7), 15 of 1 206 817 (gyr78).  The GPU weld merges them.
    # Everything else must be byte-identical.
    merged = {"eq6": 18, "eq7": 1, "gyr78": 15}
    bad = {}
    for n in eqs:
        v, t, hv, ht = REF_257[n]
        if n in merged:
            if lines[n][:2] != [str(v - merged[n]), str(t)]:
                bad[n] = lines[n]
        elif lines[n] != [str(v), str(t), hv, ht]:
            bad[n] = lines[n]
    assert not bad, bad
