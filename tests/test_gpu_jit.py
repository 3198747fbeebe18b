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
    # JIT_AUTO is the default and NVRTC is part of the image: the compile runs in the background, the calls made meanwhile
    # interpret, and once it is over the compiled kernels run
    early = c.polygonise()
    assert c.jit_wait() is True
    late = c.polygonise()
    assert late.jit == 1 and (early.jit == 1 or late.ms_compile > 0)
    assert (early.triangles, early.active) == (late.triangles, late.active)
    assert c.polygonise().ms_compile == 0
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
    third = c.polygonise()              # JIT_ON: compiled inside the call
    assert third.jit == 1 and third.ms_compile > 0 and c.polygonise().ms_compile == 0
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


def test_background_compile_switches_over_without_changing_a_bit(mcb):
    """MCB_JIT_AUTO: the first mesh of a new equation does not wait for NVRTC.  Whatever ran — the interpreter before the
    compile finished, the compiled kernels after — field, codes and soup are the same bytes."""
    import time
    c = mcb.Context(0)
    c.set_field_mode(mcb.FIELD_DENSE)
    c.set_normals(1)
    assert c.set_equation("x*x*y-z*y*y+0.3*x*z*z-0.05") == 0   # nothing else in the suite compiles this one
    c.set_grid_step(2.0 / 96)
    t0 = time.perf_counter()
    first = c.polygonise()
    ms_first = (time.perf_counter() - t0) * 1e3
    f0, p0 = c.get_field(), c.get_mesh(normals=True)
    seen = [first.jit]
    for _ in range(2000):
        cnt = c.polygonise()
        seen.append(cnt.jit)
        if cnt.jit:
            break
        time.sleep(0.002)
    assert seen[-1] == 1 and sorted(seen) == seen          # switches over once and stays
    assert cnt.ms_compile > 0 or first.jit == 1
    # (first.jit == 0 is itself the proof that the first call did not sit through the compile: a call that waits for the
    # module runs it; ms_first against cnt.ms_compile is informative only — both vary with the box)
    print("first call %.1f ms, compile %.1f ms, switched after %d calls" % (ms_first, cnt.ms_compile, len(seen)))
    f1, p1 = c.get_field(), c.get_mesh(normals=True)
    assert (first.triangles, first.active) == (cnt.triangles, cnt.active)
    assert same_bits(f0, f1) and same_bits(p0[0], p1[0]) and same_bits(p0[1], p1[1])
    # a second context with the same equation while the first one's module exists: its own cache, same behaviour
    d = mcb.Context(0)
    assert d.set_equation("x*x*y-z*y*y+0.3*x*z*z-0.05") == 0
    d.set_grid_step(2.0 / 96)
    assert d.jit_wait() is True and d.polygonise().jit == 1
    d.set_jit(mcb.JIT_OFF)
    assert d.jit_wait() is False and d.polygonise().jit == 0
    c.close(); d.close()
