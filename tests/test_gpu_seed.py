"""-m gpu: seed mode (Marching::seed_mode / set_seed, marching.cpp:42-137, 310-331) against the UNMODIFIED reference.

tests/golden/seed_cases.npz (made by tests/golden/make_seed_golden.py) holds the reference's seed-mode triangles in
BFS order; the GPU path emits the same set of triangles in the full-grid loop order, so the comparison is as sets:
a one-to-one nearest-neighbour match of the triangles within the weld tolerance (the reference's triangles are expanded
from its welded Poly_Data, and for the non-dyadic GUI step its seed-mode cube origins are -1 + i*h, not the
accumulated loop values)."""
import os

import numpy as np
import pytest

from .helpers import load_meta

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def seeds():
    return np.load(os.path.join(ROOT, "tests", "golden", "seed_cases.npz"))


def same_triangle_set(a, b, tol=2e-6):
    """Same number of triangles, and every triangle of either list has a partner in the other within tol per coordinate
    (both lists may contain exact duplicates — zero-area triangles where the surface touches grid vertices — so a strict
    bijection through nearest neighbours is not required)."""
    from scipy.spatial import cKDTree
    a = np.asarray(a, np.float64).reshape(-1, 9)
    b = np.asarray(b, np.float64).reshape(-1, 9)
    if a.shape != b.shape:
        return False
    if len(a) == 0:
        return True
    d1, _ = cKDTree(b).query(a, k=1, p=np.inf)
    d2, _ = cKDTree(a).query(b, k=1, p=np.inf)
    return bool(d1.max() < tol and d2.max() < tol)


NAMES = ["two_right", "two_left", "two_miss", "torus_all", "sphere_bound", "gui_seed_misses", "eq8_gui", "gyr34", "corner_plane"]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", [1, 2])
def test_seed_mode_keeps_the_reference_component(mcb, seeds, name, mode):
    case = load_meta(seeds)[name]
    c = mcb.Context(0)
    c.set_mesh_mode(mode)
    assert c.set_equation(case["eq"]) == 0
    c.set_grid_step(case["step"])
    c.set_scaling(*case["scale"])
    assert c.set_seed(True, *case["seed"]) == 0
    cnt = c.polygonise()
    assert cnt.triangles == case["T"], (cnt.triangles, case["T"], case["T_full"])
    if mode == 1:
        pos, _ = c.get_mesh(normals=False)
        got = pos[:, :, :3]
    else:
        vl, tl = c.get_indexed_mesh()
        got = vl[tl.astype(np.int64)] if len(tl) else np.zeros((0, 3, 3), np.float32)
        assert cnt.vertices <= max(case["V"], 0) + 2 and (case["T"] == 0 or cnt.vertices > 0)
    assert same_triangle_set(got, seeds[name + "/tris"])
    # switching the seed off gives the full grid again
    c.set_seed(False)
    assert c.polygonise().triangles == case["T_full"]
    c.close()


def test_seed_outside_the_cube_is_rejected(mcb):
    c = mcb.Context(0)
    assert c.set_seed(True, 1.5, 0.0, 0.0) == mcb.MCB_E_ARG  # Marching::set_seed returns false (marching.cpp:128-137)
    assert c.set_seed(True, 1.0, -1.0, 0.0) == 0
    c.close()
