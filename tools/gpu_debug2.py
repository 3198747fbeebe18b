import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
from oracle import refbind as R
ctx = m.Context(0)
eq = R.SPHERE
for n in (192, 256):
    step = 2.0 / n
    r = R.Ref(eq, step)
    ctx.set_equation(eq); M = ctx.set_grid_step(step)
    cnt = ctx.polygonise()
    code, tidx = ctx.get_cases()
    c3 = code.reshape(M, M, M)
    dact = (c3 != 0) & (c3 != 255)
    rec, off = ctx.get_active()
    lin = (rec & 0xFFF).astype(np.int64) + M * (((rec >> 12) & 0xFFF).astype(np.int64) + M * ((rec >> 24) & 0xFFF).astype(np.int64))
    print("n", n, "M", M, "T", cnt.triangles, "A", cnt.active, "dense active", dact.sum(), "records", len(rec), "unique", len(np.unique(lin)), "sorted", bool(np.all(np.diff(lin) > 0)))
    # per-layer active counts: dense vs records
    per_layer_dense = dact.sum(axis=(1, 2))
    per_layer_rec = np.bincount((lin // (M * M)).astype(np.int64), minlength=M)
    bad_layers = np.flatnonzero(per_layer_dense != per_layer_rec)
    print("  layers where records != dense:", len(bad_layers), bad_layers[:10], bad_layers[-10:])
    for k in (M // 2, M // 2 + 40):
        sw = r.sweep(k, k + 1, soup=False)
        print("  layer", k, "ref active", sw["active"], "dense", per_layer_dense[k], "records", per_layer_rec[k], "code equal", np.array_equal(sw["code"], c3[k].ravel()))
    if len(bad_layers):
        k = bad_layers[0]
        miss = np.setdiff1d(np.flatnonzero(dact[k].ravel()), lin[(lin // (M * M)) == k] - k * M * M)
        print("  first bad layer", k, "missing", len(miss), "j", np.unique(miss // M)[:20], "i words", np.bincount((miss % M) // 32))
    # item/tile analysis of missing
    WC = (M + 31) // 32
    allact = np.flatnonzero(dact.ravel())
    missing = np.setdiff1d(allact, lin)
    if len(missing):
        k = missing // (M * M); j = (missing // M) % M; i = missing % M
        item = (k * M + j) * WC + i // 32
        tile = item // 512
        print("  missing total", len(missing), "tiles with missing", len(np.unique(tile)), "of", (M * M * M * 0 + (M * M * WC + 511) // 512), "first tiles", np.unique(tile)[:20])
        print("  tid of missing (first 20)", ((item % 512) // 2)[:20], "item parity hist", np.bincount(item % 2))
