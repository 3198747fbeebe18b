import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
from oracle import refbind as R
from tests.helpers import same_bits
ctx = m.Context(0)
for eq, n in [(R.SPHERE, 64), (R.SPHERE, 128), ("(x*y-0.03)*(z-0.1)", 16), ("x/y*z-0.3", 20)]:
    step = 2.0 / n
    r = R.Ref(eq, step)
    sw = r.sweep(soup=False)
    M = sw["M"]
    ctx.set_equation(eq); ctx.set_grid_step(step); ctx.set_scaling(1, 1, 1); ctx.set_surface_constant(0)
    cnt = ctx.polygonise()
    code, tidx = ctx.get_cases()
    print(eq[:20], "M", M, "T", cnt.triangles, sw["T"], "A", cnt.active, sw["active"], "red", cnt.redirected, sw["redirected"])
    c3 = code.reshape(M, M, M); r3 = sw["code"].reshape(M, M, M)
    bad = c3 != r3
    print("  code mismatches", bad.sum(), "per z", np.flatnonzero(bad.any(axis=(1, 2)))[:20], "per y", np.flatnonzero(bad.any(axis=(0, 2)))[:20], "per x", np.flatnonzero(bad.any(axis=(0, 1)))[:40])
    t3 = tidx.reshape(M, M, M); rt3 = sw["table_idx"].reshape(M, M, M)
    badt = (t3 != rt3) & ~bad
    print("  tidx mismatches", badt.sum())
    for (k, j, i) in np.argwhere(badt)[:6]:
        print("   cube", i, j, k, "code", c3[k, j, i], "mine", t3[k, j, i], "ref", rt3[k, j, i])
    rec, off = ctx.get_active()
    act_ref = np.flatnonzero((sw["code"] != 0) & (sw["code"] != 255))
    lin = (rec & 0xFFF).astype(np.int64) + M * (((rec >> 12) & 0xFFF).astype(np.int64) + M * ((rec >> 24) & 0xFFF).astype(np.int64))
    print("  records", len(rec), "ref active", len(act_ref), "sorted", bool(np.all(np.diff(lin) > 0)), "subset", np.isin(lin, act_ref).all())
    if len(lin) != len(act_ref):
        missing = np.setdiff1d(act_ref, lin)
        mk = missing // (M * M); mj = (missing // M) % M; mi = missing % M
        print("  missing: k range", mk.min(), mk.max(), "j range", mj.min(), mj.max(), "i range", mi.min(), mi.max(), "i hist words", np.bincount(mi // 32))
        dact = np.flatnonzero((code != 0) & (code != 255))
        print("  dense-code active", len(dact), "vs records", len(lin))
