"""-m gpu: the per-cube view of the path — mcb_inspect_cube (= Marching::calculate_step, marching.cpp:456-595) against the
per-cube golden data of the unmodified reference, and the C++ drop-in's step-by-step mode (marching.cpp:386-428, one cube
per recalculate() call) run to completion against the reference's own step-by-step run."""
import os
import subprocess

import numpy as np
import pytest

from .helpers import configure, load_meta, same_bits
from .test_cpp_dropin import _compile, fnv1a64

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["eq8_ctor", "sphere_17", "saddle_17", "gyr34_9", "constraint_x", "nonuniform_scale"])
def test_inspect_cube_equals_calculate_step(mcb, golden, name):
    case = load_meta(golden)[name]
    c = mcb.Context(0)
    configure(c, case)
    M, cs = mcb.grid_axis(case["step"])
    code_g, tidx_g, ntri_g = golden[name + "/code"], golden[name + "/table_idx"], golden[name + "/ntri"]
    soup_g = golden[name + "/soup"]
    F = golden[name + "/field_ext"][1:-1, 1:-1, 1:-1]
    first = np.concatenate([[0], np.cumsum(ntri_g.astype(np.int64))])
    active = np.flatnonzero(ntri_g)
    inactive = np.flatnonzero(ntri_g == 0)
    picks = list(active[:: max(1, len(active) // 40)]) + list(inactive[:: max(1, len(inactive) // 10)])
    for lin in picks:
        i, j, k = int(lin % M), int((lin // M) % M), int(lin // (M * M))
        sd = c.inspect_cube(float(cs[i]), float(cs[j]), float(cs[k]))
        if case.get("cons") and code_g[lin] == 0 and sd.skipped:
            continue  # a constraint rejected a corner: the reference returns before computing anything
        assert sd.cube_code == code_g[lin] and (sd.table_idx == tidx_g[lin] or ntri_g[lin] == 0)
        # corners: 0:(0,0,0) 1:(1,0,0) 2:(1,1,0) 3:(0,1,0) 4..7 the same at z1 (marching.cpp:471-472)
        for v, (dx, dy, dz) in enumerate([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]):
            assert same_bits(np.float32(sd.corner_values[v]), F[k + dz, j + dy, i + dx])
        assert sd.n_tri_idx == 3 * ntri_g[lin]
        if ntri_g[lin]:
            pts = np.array(sd.intersect_coord[:3 * sd.n_edges], np.float32).reshape(-1, 3)
            tri = pts[np.array(sd.tri_vlist[:sd.n_tri_idx])].reshape(-1, 3, 3)
            assert same_bits(tri, soup_g[first[lin]:first[lin + 1]])
            assert list(sd.edge_list[:sd.n_edges]) == sorted(sd.edge_list[:sd.n_edges])
    c.close()


def test_dropin_step_by_step_mode_equals_reference(mcb, tmp_path):
    seeds = np.load(os.path.join(ROOT, "tests", "golden", "seed_cases.npz"))
    meta = load_meta(seeds)
    exe = _compile(tmp_path)
    names = [n for n in meta if n.startswith("step_")]
    args = []
    for n in names:
        c = meta[n]
        args += [n, c["eq"], repr(c["step"]), repr(c["scale"][0]), repr(c["scale"][1]), repr(c["scale"][2]), "0"]
    out = subprocess.check_output([exe] + args, text=True)
    lines = {l.split()[0]: l.split()[1:] for l in out.strip().splitlines()}
    for n in names:
        v, t = seeds[n + "/vertex_list"], seeds[n + "/tri_list"]
        assert lines[n + "_calls"] == [str(meta[n]["calls"])]
        assert lines[n][:2] == [str(len(v)), str(len(t))], (n, lines[n])
        assert lines[n][2] == fnv1a64(v.tobytes()) and lines[n][3] == fnv1a64(t.tobytes()), n
