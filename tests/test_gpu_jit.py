"""-m gpu: the run-time specialised evaluator (mcb_set_jit: the equation's fused grid program compiled by NVRTC into a
straight-line kernel) against the interpreter, which the other tests pin to the unmodified reference.

Bar: the field at every grid vertex, the sign planes' consequences (codes, counts) and every mesh output are byte for
byte what the interpreter produces — the generated kernel executes the same fp32 operations in the same order."""
import numpy as np
import pytest

from .helpers import configure, load_meta, same_bits
from .test_gpu_parity import CASE_NAMES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(mcb):
    a, b = mcb.Context(0), mcb.Context(0)
    a.set_jit(mcb.JIT_OFF)
    b.set_jit(mcb.JIT_ON)
    for c in (a, b):
        c.set_mesh_mode(mcb.MESH_SOUP | mcb.MESH_INDEXED)
        c.set_normals(1)
    yield a, b
    a.close()
    b.close()


@pytest.mark.parametrize("name", CASE_NAMES)
def test_jit_equals_interpreter_on_the_golden_cases(mcb, pair, golden, name):
    case = load_meta(golden)[name]
    res = []
    for c in pair:
        configure(c, case)
        cnt = c.polygonise()
        res.append((cnt, c.get_field(), c.get_cases(), c.get_mesh(normals=True), c.get_indexed_mesh(normals=True)))
    (ci, fi, (codei, tidxi), (pi, ni), (vli, tli, vni)), (cj, fj, (codej, tidxj), (pj, nj), (vlj, tlj, vnj)) = res
    assert ci.jit == 0 and cj.jit == 1
    assert same_bits(fi, fj), "field"
    assert same_bits(fj, golden[name + "/field_ext"][1:-1, 1:-1, 1:-1]), "field vs the reference evaluator"
    assert np.array_equal(codei, codej) and np.array_equal(tidxi, tidxj)
    for k in ("active", "triangles", "ambiguous", "redirected", "vertices"):
        assert getattr(ci, k) == getattr(cj, k), k
    assert same_bits(pi, pj) and same_bits(ni, nj) and same_bits(vli, vlj) and np.array_equal(tli, tlj) and same_bits(vni, vnj)


def test_jit_compiles_once_per_equation_and_survives_parameter_changes(mcb):
    c = mcb.Context(0)
    assert c.polygonise().jit == 1          # JIT_AUTO is the default and NVRTC is part of the image
    c.set_jit(mcb.JIT_ON)
    assert c.set_equation("x^2+y^2+z^2-0.49") == 0
    c.set_grid_step(2.0 / 64)
    first = c.polygonise()
    assert first.jit == 1 and first.ms_compile > 0
    c.set_grid_step(2.0 / 200)          # another grid: table offsets and sizes are kernel arguments
    c.set_scaling(1.1, 0.9, 1.0)
    c.set_surface_constant(0.05)
    again = c.polygonise()
    assert again.jit == 1 and again.ms_compile == 0
    ref = mcb.Context(0)
    ref.set_jit(mcb.JIT_OFF)
    assert ref.set_equation("x^2+y^2+z^2-0.49") == 0
    ref.set_grid_step(2.0 / 200); ref.set_scaling(1.1, 0.9, 1.0); ref.set_surface_constant(0.05)
    r = ref.polygonise()
    assert (again.triangles, again.active) == (r.triangles, r.active) and same_bits(c.get_field(), ref.get_field())
    assert c.set_equation("x*y-z^3+0.1") == 0    # a new equation compiles again; a constraint does not
    assert c.set_equation("x", slot=1) == 0 and c.set_constraint(0, "<", 0.5, True) == 0
    third = c.polygonise()
    assert third.ms_compile > 0 and c.polygonise().ms_compile == 0
    c.set_jit(mcb.JIT_OFF)
    assert c.polygonise().jit == 0
    c.close(); ref.close()


def test_jit_at_1024_matches_the_interpreter(mcb):
    res = []
    for jit in (mcb.JIT_OFF, mcb.JIT_ON):
        c = mcb.Context(0)
        c.set_jit(jit)
        assert c.set_equation("(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)") == 0
        c.set_grid_step(2.0 / 512)
        c.set_normals(1)
        cnt = c.polygonise()
        cnt = c.polygonise()
        pos, nrm = c.get_mesh(normals=True)
        res.append((cnt.triangles, cnt.active, cnt.ambiguous, pos, nrm, cnt.ms_eval))
        c.close()
    assert res[0][:3] == res[1][:3] and same_bits(res[0][3], res[1][3]) and same_bits(res[0][4], res[1][4])


def test_auto_leaves_pow_heavy_programs_to_the_interpreter(mcb):
    eq = "+".join("(x*%d.5+y*z)^%d" % (i, 2 + i % 3) for i in range(1, 11)) + "-3"
    res = []
    for mode in (mcb.JIT_AUTO, mcb.JIT_ON):
        c = mcb.Context(0)
        c.set_jit(mode)
        assert c.set_equation(eq) == 0
        c.set_grid_step(2.0 / 40)
        cnt = c.polygonise()
        res.append((cnt.jit, cnt.triangles, c.get_field()))
        c.close()
    assert res[0][0] == 0 and res[1][0] == 1          # 10 general powers: auto interprets, ON compiles anyway
    assert res[0][1] == res[1][1] and same_bits(res[0][2], res[1][2])
