"""-m gpu: the block-field mode (MCB_FIELD_SPARSE / MCB_FIELD_AUTO: an interval proof per 32 x 4 x 4 vertex block decides
which blocks can hold a sign change; only those are evaluated, classify skips the rest) against
MCB_FIELD_DENSE, which the other -m gpu tests pin to the unmodified reference.

Bar: every output — counts, triangle soup, gradient normals, welded Poly_Data, normal.h normals, seed-mode component,
constrained and slabbed runs — is byte for byte what the dense mode produces.  The sparse contexts run with
$MCB_POISON_FIELD=1: the field buffer is NaN-filled before every run, so a read outside the evaluated blocks cannot
go unnoticed.
"""
import os

import numpy as np
import pytest

from .helpers import configure, load_meta, same_bits
from .test_gpu_parity import CASE_NAMES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(mcb):
    dense = mcb.Context(0)
    os.environ["MCB_POISON_FIELD"] = "1"
    try:
        sparse = mcb.Context(0)
    finally:
        del os.environ["MCB_POISON_FIELD"]
    sparse.set_field_mode(mcb.FIELD_SPARSE)
    for c in (dense, sparse):
        c.set_mesh_mode(mcb.MESH_SOUP | mcb.MESH_INDEXED)
    yield dense, sparse
    dense.close()
    sparse.close()


def run_both(pair, setup, normals):
    out = []
    for c in pair:
        setup(c)
        c.set_normals(normals)
        cnt = c.polygonise()
        pos, nrm = c.get_mesh(normals=(normals == 1))
        if normals:
            vl, tl, vn = c.get_indexed_mesh(normals=True)
        else:
            (vl, tl), vn = c.get_indexed_mesh(), None
        out.append((cnt, pos, nrm, vl, tl, vn))
    return out


def assert_same(d, s):
    (cd, pd, nd, vd, td, vnd), (cs, ps, ns, vs, ts, vns) = d, s
    for k in ("cubes", "active", "triangles", "ambiguous", "redirected", "vertices", "M"):
        assert getattr(cd, k) == getattr(cs, k), k
    assert same_bits(pd, ps), "soup positions"
    assert (nd is None and ns is None) or same_bits(nd, ns), "soup normals"
    assert same_bits(vd, vs), "vertex_list"
    assert np.array_equal(td, ts), "tri_list"
    assert (vnd is None and vns is None) or same_bits(vnd, vns), "vertex normals"


@pytest.mark.parametrize("name", CASE_NAMES)
@pytest.mark.parametrize("normals", [1, 2])
def test_sparse_field_equals_dense_on_the_golden_cases(mcb, pair, golden, name, normals):
    case = load_meta(golden)[name]
    d, s = run_both(pair, lambda c: configure(c, case), normals)
    assert_same(d, s)
    assert s[0].field_mode == mcb.FIELD_SPARSE and d[0].field_mode == mcb.FIELD_DENSE
    assert s[0].field_blocks > 0 or s[0].active == 0   # undecided blocks need not hold a sign change, active cubes need blocks


EQS = {
    "sphere": "x^2+y^2+z^2-0.49",
    "torus": "(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)",
    "eq8": None,  # example_files/equation_8.txt via the golden meta
    "quirk": "x^2-y/z+x*y^3",
}


@pytest.mark.parametrize("n,eq", [(96, "sphere"), (161, "torus"), (130, "quirk"), (257, "sphere")])
def test_sparse_field_equals_dense_at_odd_sizes(mcb, pair, n, eq):
    """grid sizes that leave partial blocks on every side; scale != 1; iso != 0"""
    def setup(c):
        assert c.set_equation(EQS[eq]) == 0
        c.set_grid_step(2.0 / n)
        c.set_scaling(1.1, 0.9, 1.0)
        c.set_surface_constant(0.01)
        for i in range(3):
            c.set_constraint(i, ">", 0.0, False)
    d, s = run_both(pair, setup, 1)
    assert d[0].triangles > 0
    assert_same(d, s)


def test_sparse_field_with_constraints_slabs_and_seed(mcb, pair):
    def setup(c):
        assert c.set_equation("(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)") == 0
        c.set_grid_step(2.0 / 100)
        c.set_scaling(1.0, 1.0, 1.0)
        c.set_surface_constant(0.0)
        assert c.set_equation("x+y", slot=1) == 0
        assert c.set_constraint(0, "<", 0.3, True) == 0
        for i in (1, 2):
            c.set_constraint(i, ">", 0.0, False)
    d, s = run_both(pair, setup, 1)
    assert 0 < d[0].triangles
    assert_same(d, s)
    M = d[0].M
    for r in range(3):
        k0, k1 = mcb.slab_range(M, r, 3)
        for c in pair:
            c.set_slab(k0, k1)
        dd, ss = run_both(pair, lambda c: None, 1)
        assert_same(dd, ss)
    for c in pair:
        c.set_slab(0, M)
        c.set_constraint(0, ">", 0.0, False)
        assert c.set_seed(True, 0.74, 0.0, 0.0) == 0  # on the outer equator of the torus
    dd, ss = run_both(pair, lambda c: None, 1)
    assert 0 < dd[0].triangles
    assert_same(dd, ss)
    for c in pair:
        c.set_seed(False)


def test_sparse_field_has_no_dense_field_to_read(mcb, pair):
    dense, sparse = pair
    for c in pair:
        assert c.set_equation("x^2+y^2+z^2-0.49") == 0
        c.set_grid_step(0.1)
        c.polygonise()
    assert dense.get_field().shape[0] > 0
    with pytest.raises(mcb.McbError):
        sparse.get_field()
    # switching back materialises it again
    sparse.set_field_mode(mcb.FIELD_DENSE)
    sparse.polygonise()
    assert same_bits(sparse.get_field(), dense.get_field())
    sparse.set_field_mode(mcb.FIELD_SPARSE)


def test_field_auto_is_the_block_mode_from_the_first_call(mcb):
    """MCB_FIELD_AUTO needs no earlier run of a configuration: the first polygonisation of a new equation, a moved iso
    value and a coarse grid all run the block-field mode, with the dense mode's outputs."""
    c = mcb.Context(0)
    d = mcb.Context(0)
    c.set_field_mode(mcb.FIELD_AUTO)
    d.set_field_mode(mcb.FIELD_DENSE)
    for x in (c, d):
        x.set_normals(1)

    def both():
        a, b = c.polygonise(), d.polygonise()
        pa, na = c.get_mesh(normals=True)
        pb, nb = d.get_mesh(normals=True)
        assert a.field_mode == mcb.FIELD_SPARSE and b.field_mode == mcb.FIELD_DENSE
        assert (a.active, a.triangles, a.ambiguous, a.redirected) == (b.active, b.triangles, b.ambiguous, b.redirected)
        assert same_bits(pa, pb) and same_bits(na, nb)
        return a
    for x in (c, d):
        assert x.set_equation("x^2+y^2+z^2-0.49") == 0
        x.set_grid_step(2.0 / 320)
    first = both()
    nblocks = ((first.M + 3 + 31) // 32) * ((first.M + 3 + 3) // 4) ** 2
    assert 0 < first.field_blocks < 0.25 * nblocks        # most of the grid is proven empty, not evaluated
    for x in (c, d):
        x.set_surface_constant(0.1)                        # another surface of the same equation
    both()
    for x in (c, d):
        assert x.set_equation("(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)") == 0   # another equation
    both()
    for x in (c, d):
        x.set_grid_step(2.0 / 24)                          # a coarse grid: the surface is everywhere
    both()
    c.close()
    d.close()


@pytest.mark.parametrize("eq,n", [("x^2+y^2+z^2-0.49", 200), ("(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)", 160),
                                  ("x^2-y/z+x*y^3", 130), ("x/(y*y+0.01)-z^3", 96), ("(x*y)^3+z", 96), ("2^(x*y)-z-1", 64)])
def test_the_interval_proof_changes_nothing(mcb, eq, n):
    """Deciding blocks by interval arithmetic against evaluating every block ($MCB_NO_INTERVAL=1): identical outputs;
    includes programs the proof has to give up on (division by an interval around zero, a variable exponent)."""
    decided = mcb.Context(0)
    os.environ["MCB_NO_INTERVAL"] = "1"
    try:
        everything = mcb.Context(0)
    finally:
        del os.environ["MCB_NO_INTERVAL"]
    out = []
    for c in (decided, everything):
        c.set_field_mode(mcb.FIELD_SPARSE)
        c.set_mesh_mode(mcb.MESH_SOUP | mcb.MESH_INDEXED)
        c.set_normals(1)
        assert c.set_equation(eq) == 0
        c.set_grid_step(2.0 / n)
        cnt = c.polygonise()
        out.append((cnt, c.get_active(), c.get_mesh(normals=True), c.get_indexed_mesh(normals=True), c.get_cases()))
    (ca, ra, ma, ia, ka), (cb, rb, mb, ib, kb) = out
    assert (ca.active, ca.triangles, ca.ambiguous, ca.redirected, ca.vertices) == (cb.active, cb.triangles, cb.ambiguous, cb.redirected, cb.vertices)
    assert ca.field_blocks <= cb.field_blocks
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    assert same_bits(ma[0], mb[0]) and same_bits(ma[1], mb[1])
    assert same_bits(ia[0], ib[0]) and np.array_equal(ia[1], ib[1]) and same_bits(ia[2], ib[2])
    assert np.array_equal(ka[0], kb[0]) and np.array_equal(ka[1], kb[1])
    decided.close()
    everything.close()
