"""Where the end-to-end time of one reference-facing call goes (not a test, not a bench)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
eq, n = "x^2+y^2+z^2-0.49", 1024
ctx = m.Context(0)
ctx.set_mesh_mode(m.MESH_INDEXED); ctx.set_normals(1)
ctx.set_equation(eq); ctx.set_grid_step(2.0 / n)
c = ctx.polygonise()
capV, capT = int(c.vertices) + 1024, int(c.triangles) + 1024
bv = torch.empty((capV, 3), dtype=torch.float32).pin_memory(); bt = torch.empty((capT, 3), dtype=torch.int32).pin_memory(); bn = torch.empty((capV, 3), dtype=torch.float32).pin_memory()
acc = {}
def tick(name, t0):
    torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
for it in range(12):
    if it == 2: acc.clear()
    t = time.perf_counter(); ctx.set_equation(eq); tick("set_equation", t)
    t = time.perf_counter(); ctx.set_grid_step(2.0 / n); ctx.set_slab(0, 1025); tick("set_grid", t)
    t = time.perf_counter(); c = ctx.polygonise(); tick("polygonise", t)
    t = time.perf_counter(); ctx.get_indexed_mesh_into(bv.data_ptr(), bt.data_ptr(), bn.data_ptr(), capV, capT); tick("get_mesh", t)
print({k: round(v / 10, 3) for k, v in acc.items()}, "device ms_total", round(c.ms_total, 3))
acc.clear()
for it in range(10):
    t = time.perf_counter(); c = ctx.polygonise(); tick("polygonise_only", t)
print({k: round(v / 10, 3) for k, v in acc.items()}, "device ms_total", round(c.ms_total, 3))
ctx.set_mesh_mode(m.MESH_SOUP)
ctx.polygonise()
acc.clear()
for it in range(10):
    t = time.perf_counter(); c = ctx.polygonise(); tick("polygonise_soup", t)
print({k: round(v / 10, 3) for k, v in acc.items()}, "device ms_total", round(c.ms_total, 3))
