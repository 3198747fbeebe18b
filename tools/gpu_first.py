"""Diagnostic script for GPU bring-up: prints parity summaries and stage timings (not a test, not a bench)."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
from tests.helpers import configure, load_meta, same_bits
g = np.load(os.path.join(ROOT, "tests/golden/cases.npz"))
meta = load_meta(g)
ctx = m.Context(0)
for name, case in meta.items():
    configure(ctx, case)
    cnt = ctx.polygonise()
    code, tidx = ctx.get_cases()
    F = ctx.get_field()
    pos, nrm = ctx.get_mesh()
    soup = g[name + "/soup"]
    print("%-18s M=%3d T=%6d/%6d A=%5d/%5d amb=%d/%d red=%d/%d field_ok=%s code_ok=%s tidx_ok=%s soup_ok=%s" % (
        name, cnt.M, cnt.triangles, case["T"], cnt.active, case["active"], cnt.ambiguous, case["ambiguous"], cnt.redirected,
        case["redirected"], same_bits(F, g[name + "/field_ext"][1:-1, 1:-1, 1:-1]), np.array_equal(code, g[name + "/code"]),
        np.array_equal(tidx, g[name + "/table_idx"]), pos.shape[0] == soup.shape[0] and same_bits(pos[:, :, :3], soup)))
for eq, n in [("x^2+y^2+z^2-0.49", 256), ("x^2+y^2+z^2-0.49", 1024), ("(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)", 1024)]:
    ctx.set_equation(eq); ctx.set_grid_step(2.0 / n); ctx.set_scaling(1, 1, 1); ctx.set_surface_constant(0)
    for i in range(3): ctx.set_constraint(i, '>', 0.0, False)
    for it in range(3):
        t0 = time.time(); cnt = ctx.polygonise(); dt = time.time() - t0
        print(n, eq[:20], json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in cnt.as_dict().items()}), "wall %.1f ms" % (dt * 1e3))
