"""A/B of the default emitter's grid size (blocks per SM, $MCB_EMIT_BLOCKS_PER_SM): device time of the emission stage."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
import bench
for wl, n in (("sphere", 1024), ("gyr78", 1024), ("sphere", 2048)):
    for bps in (8, 10, 12, 16, 20, 24, 32, 48):
        os.environ["MCB_EMIT_BLOCKS_PER_SM"] = str(bps)
        ctx = m.Context(0)
        ctx.set_field_mode(m.FIELD_AUTO)
        assert ctx.set_equation(bench.WORKLOADS[wl]) == 0
        ctx.jit_wait()
        ctx.set_grid_step(2.0 / n)
        ctx.set_normals(1)
        for _ in range(3):
            ctx.polygonise()
        reps, acc = 10, 0.0
        tot = 0.0
        for _ in range(reps):
            c = ctx.polygonise()
            acc += c.ms_emit / reps
            tot += c.ms_total / reps
        print(json.dumps({"workload": wl, "n": n, "blocks_per_sm": bps, "ms_emit": round(acc, 4), "ms_total": round(tot, 4)}), flush=True)
        ctx.close()
