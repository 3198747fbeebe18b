"""Generates tests/golden/seed_cases.npz from the UNMODIFIED reference in seed mode (Marching::set_seed + seed_mode(true)
+ recalculate(), marching.cpp:42-137, 310-331) through oracle/_ref.  Run in the build container only.
Per case: the expanded triangles (vertex_list[tri_list], BFS order) — the GPU path keeps the same SET of triangles."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402

TWO = "((x-0.5)^2+y^2+z^2-0.04)*((x+0.5)^2+y^2+z^2-0.04)"
CASES = {
    "two_right": dict(eq=TWO, step=2.0 / 32, scale=1.0, seed=(0.5, 0.0, 0.2)),
    "two_left": dict(eq=TWO, step=2.0 / 32, scale=1.0, seed=(-0.69, 0.0, 0.0)),
    "two_miss": dict(eq=TWO, step=2.0 / 32, scale=1.0, seed=(0.0, 0.0, 0.0)),
    "torus_all": dict(eq=R.TORUS, step=2.0 / 32, scale=1.0, seed=(0.75, 0.0, 0.0)),
    "sphere_bound": dict(eq="x^2+y^2+z^2-0.98", step=2.0 / 16, scale=1.0, seed=(0.0, 0.0, 0.98)),  # reaches the bound check
    "gui_seed_misses": dict(eq=R.EXAMPLE_EQUATIONS[8], step=0.2, scale=1.1, seed=(0.8, 0.8, 0.9)),  # GUI defaults, drawer.cpp:40-44
    "eq8_gui": dict(eq=R.EXAMPLE_EQUATIONS[8], step=0.2, scale=1.1, seed="vertex"),                 # non-dyadic step, scaled
    "gyr34": dict(eq=R.GYR34, step=2.0 / 16, scale=1.0, seed="vertex"),                             # many components
    "corner_plane": dict(eq="x+y-1.9", step=2.0 / 32, scale=1.0, seed=(0.95, 0.95, 0.0)),           # the bound check cuts it
}
out, meta = {}, {}
for name, c in CASES.items():
    r = R.Ref(c["eq"], c["step"], scale=(c["scale"],) * 3)
    full_v, full_t = r.recalculate()
    if c["seed"] == "vertex":  # a point on the surface: the middle vertex of the full mesh, in scaled coordinates
        p = full_v[len(full_v) // 2] * c["scale"]
        c["seed"] = tuple(float(np.clip(x, -1, 1)) for x in p)
    m = r.seed_recalculate(*c["seed"])
    v, t = m
    tris = v[t.astype(np.int64)] if len(t) else np.zeros((0, 3, 3), np.float32)
    out[name + "/tris"] = tris.astype(np.float32)
    meta[name] = dict(eq=c["eq"], step=c["step"], scale=[c["scale"]] * 3, seed=list(c["seed"]), T=int(len(t)), V=int(len(v)),
                      T_full=int(len(full_t)))
    print(name, meta[name]["T"], "of", meta[name]["T_full"])
# step-by-step mode run to completion: the reference's own traversal (`< 1.0` bounds, carried coordinates)
STEP_CASES = {
    "step_eq1_ctor": dict(eq=R.EXAMPLE_EQUATIONS[1], step=0.25, scale=1.0),
    "step_sphere": dict(eq=R.SPHERE, step=0.25, scale=1.0),
    "step_eq8_gui": dict(eq=R.EXAMPLE_EQUATIONS[8], step=0.2, scale=1.1),
}
for name, c in STEP_CASES.items():
    r = R.Ref(c["eq"], c["step"], scale=(c["scale"],) * 3)
    n, v, t = r.step_all()
    out[name + "/vertex_list"] = v
    out[name + "/tri_list"] = t
    meta[name] = dict(eq=c["eq"], step=c["step"], scale=[c["scale"]] * 3, calls=int(n), V=int(len(v)), T=int(len(t)))
    print(name, n, "calls", len(v), "vertices", len(t), "triangles")
out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "seed_cases.npz"), **out)
