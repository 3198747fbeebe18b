"""One workload, a few polygonisations: the command profiled under ncu (not a test, not a bench).
usage: prof_step.py [workload|equation] [resolution] [repetitions] [mesh mode 1 soup / 2 indexed / 3 both] [field mode 0 dense / 2 auto]"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
import bench
eq = sys.argv[1] if len(sys.argv) > 1 else "sphere"
eq = bench.WORKLOADS.get(eq, eq)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
field = int(sys.argv[5]) if len(sys.argv) > 5 else m.FIELD_AUTO
ctx = m.Context(0)
assert ctx.set_equation(eq) == 0
ctx.jit_wait()
ctx.set_grid_step(2.0 / n)
ctx.set_normals(1)
ctx.set_mesh_mode(mode)
ctx.set_field_mode(field)
for it in range(reps):
    c = ctx.polygonise()
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in c.as_dict().items()}))
ctx.close()
