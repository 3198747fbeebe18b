"""A/B of kernel variants selected by environment switches (read at mcb_create): device time per stage, a few workloads.
Not a test, not a bench: a development aid whose output goes to profiles/ when a decision is taken from it."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
import bench
import torch
out = []
for wl, n in (("sphere", 1024), ("gyr78", 1024), ("torus", 1024), ("sphere", 2048)):
    for env in ({"MCB_EMIT": "2"}, {"MCB_EMIT": "3"}, {"MCB_EMIT": "4"}):
        for k in ("MCB_EMIT", "MCB_WELD_EXACT", "MCB_NO_INTERVAL"):
            os.environ.pop(k, None)
        os.environ.update(env)
        if env.get("MCB_NO_INTERVAL") and n > 1024:
            continue
        ctx = m.Context(0)
        ctx.set_field_mode(m.FIELD_AUTO)
        assert ctx.set_equation(bench.WORKLOADS[wl]) == 0
        ctx.jit_wait()
        ctx.set_grid_step(2.0 / n)
        ctx.set_normals(1)
        for mesh in (m.MESH_SOUP, m.MESH_INDEXED):
            if mesh == m.MESH_INDEXED and env.get("MCB_EMIT") != "2":
                continue
            ctx.set_mesh_mode(mesh)
            for _ in range(3):
                ctx.polygonise()
            acc = {}
            reps = 10
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                c = ctx.polygonise()
                for k in ("ms_tables", "ms_eval", "ms_classify", "ms_fill", "ms_emit", "ms_weld", "ms_total"):
                    acc[k] = acc.get(k, 0.0) + getattr(c, k) / reps
            wall = (time.perf_counter() - t0) / reps * 1e3
            row = {"workload": wl, "n": n, "env": env, "mesh": "soup" if mesh == m.MESH_SOUP else "indexed", "wall_ms": round(wall, 4),
                   "T": int(c.triangles), "A": int(c.active), "blocks": int(c.field_blocks), **{k: round(v, 4) for k, v in acc.items()}}
            print(json.dumps(row), flush=True)
        ctx.close()
