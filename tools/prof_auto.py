"""The command profiled under ncu for profiles/r01_ncu_full_summary_step9.txt: N indexed-mesh polygonisations of the sphere
at 1024^3 in field mode auto (the first writes the whole field, the following ones only its signs); not a test, not a bench."""
import importlib, sys
sys.path.insert(0, ".")
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
c = m.Context(0); c.set_equation("x^2+y^2+z^2-0.49"); c.set_grid_step(2.0/1024); c.set_normals(1); c.set_mesh_mode(3); c.set_field_mode(m.FIELD_AUTO)
for it in range(int(sys.argv[1])):
    cnt = c.polygonise(); print(cnt.field_mode, cnt.jit, round(cnt.ms_total, 4))
