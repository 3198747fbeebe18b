"""Generates tests/golden/cases.npz from the UNMODIFIED reference (oracle/_ref/libmcref.so, built by oracle/Makefile
from /root/reference).  Run in the build container only; the vectors are committed because /root/reference does
not exist on the GPU box.

Per case: per-cube cube_code / table_idx / ntri in loop order, the triangle soup (exact per-cube positions from
Step_Data), the welded Poly_Data (vertex_list, tri_list) and normal.h normals, the grid coordinates, and the field at
all grid vertices (Marching::evaluate).  Counts are cross-checked against SURVEY.md Appendix B where it lists them.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402
from tests.mc_numpy import apron_coords  # noqa: E402

E = R.EXAMPLE_EQUATIONS
CASES = {}
for n in range(1, 9):
    CASES["eq%d_gui" % n] = dict(eq=E[n], step=0.2, scale=1.1)      # GUI defaults, drawer.cpp:40,43
    CASES["eq%d_ctor" % n] = dict(eq=E[n], step=0.25, scale=1.0)    # Marching ctor defaults, marching.cpp:23-37
CASES["sphere_17"] = dict(eq=R.SPHERE, step=0.125, scale=1.0)
CASES["sphere_33_iso"] = dict(eq=R.SPHERE, step=0.0625, scale=1.0, iso=0.07)
CASES["torus_33"] = dict(eq=R.TORUS, step=0.0625, scale=1.0)
CASES["saddle_17"] = dict(eq="(x*y-0.03)*(z-0.1)", step=0.125, scale=1.0)
CASES["gyr34_9"] = dict(eq=R.GYR34, step=0.25, scale=1.0)
CASES["gyr78_17"] = dict(eq=R.GYR78, step=0.125, scale=1.0)
CASES["quirk_div"] = dict(eq="x/y*z-0.3", step=0.1, scale=1.0)          # division by zero planes -> inf/nan corners
CASES["quirk_neg"] = dict(eq="-x^2+y*-z+0.2", step=0.125, scale=(1.0, 0.9, 1.3))
CASES["nonuniform_scale"] = dict(eq=R.SPHERE, step=0.15, scale=(1.3, 0.8, 1.1), iso=-0.1)
CASES["constraint_x"] = dict(eq=R.SPHERE, step=0.125, scale=1.0, cons=[("x", ">", -0.5)])
CASES["constraint_2"] = dict(eq=E[8], step=0.125, scale=1.0, cons=[("x+y", "<=", 0.25), ("z^2", "<", 0.36)])

# SURVEY.md Appendix B (welded verts, tris) for the cases it lists
SURVEY_COUNTS = {"eq1_gui": (132, 462), "eq2_gui": (0, 0), "eq3_gui": (104, 200), "eq4_gui": (179, 340), "eq5_gui": (306, 576),
                 "eq6_gui": (180, 306), "eq7_gui": (54, 104), "eq8_gui": (204, 400), "eq1_ctor": (90, 306), "eq2_ctor": (108, 388),
                 "eq3_ctor": (89, 164), "eq4_ctor": (135, 252), "eq5_ctor": (218, 386), "eq6_ctor": (129, 248),
                 "eq7_ctor": (38, 72), "eq8_ctor": (180, 352), "gyr34_9": (742, 1366), "gyr78_17": (5311, 10296)}


def scale3(s):
    return (s, s, s) if not isinstance(s, tuple) else s


def main():
    out = {}
    meta = {}
    for name, c in CASES.items():
        sc = scale3(c["scale"])
        r = R.Ref(c["eq"], c["step"], sc, c.get("iso", 0.0))
        for i, (lhs, op, rhs) in enumerate(c.get("cons", [])):
            assert r.set_constraint(i, lhs, op, rhs, True)
        v, t = r.recalculate()
        nrm = r.normals()
        sw = r.sweep(corners=False, soup=True, weld=True)
        v2, t2 = r.mesh()
        assert np.array_equal(v, v2) and np.array_equal(t, t2), name  # harness loop == Marching::recalculate
        M, coords = r.coords()
        if name in SURVEY_COUNTS:
            assert (len(v), len(t)) == SURVEY_COUNTS[name], (name, len(v), len(t))
        # field through Marching::evaluate (with scaling) at the grid vertices plus a one-vertex apron
        # (index v+1 holds vertex v in [-1, M+1]); the apron feeds the central-difference normals
        cs = apron_coords(coords, c["step"])
        Z, Y, X = np.meshgrid(cs, cs, cs, indexing="ij")
        pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1).astype(np.float32)
        field = r.eval_points(pts, scaled=True).reshape(M + 3, M + 3, M + 3)
        out[name + "/code"] = sw["code"]
        out[name + "/table_idx"] = sw["table_idx"]
        out[name + "/ntri"] = sw["ntri"]
        out[name + "/soup"] = sw["soup"]
        out[name + "/vertex_list"] = v
        out[name + "/tri_list"] = t
        out[name + "/normals"] = nrm
        out[name + "/coords"] = coords
        out[name + "/field_ext"] = field
        meta[name] = dict(eq=c["eq"], step=c["step"], scale=list(sc), iso=c.get("iso", 0.0), cons=c.get("cons", []), M=M,
                          active=sw["active"], ambiguous=sw["ambiguous"], redirected=sw["redirected"], T=sw["T"],
                          welded=len(v))
        print(name, meta[name]["M"], "T", sw["T"], "A", sw["active"], "amb", sw["ambiguous"], "red", sw["redirected"], "welded", len(v))
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cases.npz"), **out)


if __name__ == "__main__":
    main()
