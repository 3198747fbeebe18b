"""-m "not gpu": the interval proof of csrc/mcb_interval.h against brute force (oracle/host_interp.cpp runs both on the host).

For every 32 x 4 x 4 box of a grid: when the interval evaluation of the fused grid program says "every vertex is above iso"
or "no vertex is above iso", the sign of f at EVERY vertex of the box (the same fp32 operation sequence the kernels
execute) must agree, and every value must lie inside the interval.  Zero wrong proofs is the bar; how many boxes are
decided is reported, not asserted (it only affects speed)."""
import ctypes as C
import os

import numpy as np
import pytest

from .test_host_logic import _random_equation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    so = os.path.join(ROOT, "oracle", "libmcoracle_host.so")
    if not os.path.exists(so):
        pytest.skip("oracle/libmcoracle_host.so not built")
    L = C.CDLL(so)
    L.mcoh_interval_check.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
    return L


def check(L, eq, n, iso=0.0, scale=(1.0, 1.0, 1.0), box=(32, 4, 4), lo=-1.0, hi=1.0):
    ax = [(np.linspace(lo, hi, n).astype(np.float32) * np.float32(s)).astype(np.float32) for s in scale]
    st = np.zeros(6, np.int64)
    rc = L.mcoh_interval_check(eq.encode(), ax[0].ctypes.data, ax[1].ctypes.data, ax[2].ctypes.data, n, box[0], box[1], box[2], iso, st.ctypes.data)
    assert rc == 0, (eq, rc)
    return dict(boxes=int(st[0]), proven0=int(st[1]), proven1=int(st[2]), uniform=int(st[3]), wrong=int(st[4]), outside=int(st[5]))


EQUATIONS = [
    "x^2+y^2+z^2-0.49",
    "(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)",
    "x+y",
    "x^2*y^2+x^2*z^2+z^2*y^2+x*y*z",
    "(x+0.5)*(x^2+y^2+z^2-0.5^2*0.5^2*0.25)+0.5*z^2",
    "(x^2+y^2-(1/16))^2+(y^2+z^2-(1/16))^2+(z^2+x^2-(1/16))^2-8*(x^2+y^2+z^2-(1/4))^2",
    "x/(y*y+0.01)-z^3", "(x*y)^3+z", "(x*y)^4-z", "2^(x*y)-z-1", "x/y-z", "(x+y)^0.5-z", "(x*y+2)^-1.5-z*z", "-(x-y)^2+z",
    "1/(x*y*z)", "(x*y)^2^(z+1)", "x*y*z/0", "(x*y)^0-z",
]


@pytest.mark.parametrize("eq", EQUATIONS)
def test_no_wrong_proof(host, eq):
    for n, iso, scale in ((65, 0.0, (1.0, 1.0, 1.0)), (97, 0.013, (1.1, 0.9, 1.3)), (40, -0.2, (3.0, 3.0, 3.0))):
        r = check(host, eq, n, iso, scale)
        assert r["wrong"] == 0 and r["outside"] == 0, (eq, n, r)


def test_the_bench_fields_are_mostly_decided(host):
    """not a correctness statement: on the bench fields the proof decides nearly every box that really is uniform"""
    for eq, frac in (("x^2+y^2+z^2-0.49", 0.99), ("(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)", 0.6)):
        r = check(host, eq, 129)
        assert r["wrong"] == 0
        assert r["proven0"] + r["proven1"] >= frac * r["uniform"], (eq, r)


def test_random_equations(host, mcb):
    rng = np.random.default_rng(11)
    done = 0
    stats = dict(boxes=0, decided=0, uniform=0)
    while done < 150:
        eq = _random_equation(rng, int(rng.integers(2, 6)))
        if len(eq) > 120 or not mcb.parse_ok(eq):
            continue
        done += 1
        r = check(host, eq, 33, iso=float(rng.normal() * 0.3), scale=tuple(float(x) for x in rng.uniform(0.5, 2.0, 3)), box=(8, 4, 4))
        assert r["wrong"] == 0 and r["outside"] == 0, (eq, r)
        stats["boxes"] += r["boxes"]; stats["decided"] += r["proven0"] + r["proven1"]; stats["uniform"] += r["uniform"]
    assert stats["decided"] > 0.3 * stats["uniform"], stats   # the proof is not vacuous on random programs either
