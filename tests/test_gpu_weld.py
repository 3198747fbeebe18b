"""-m gpu: the indexed, welded mesh (MCB_MESH_INDEXED) against Poly_Data of the UNMODIFIED reference.

Golden vectors: vertex_list / tri_list exactly as Marching::recalculate() left them (add_step_to_poly_data / add_point,
marching.cpp:599-654, std::set with the tolerance comparator of marching.h:38-54).  Bar: byte for byte — same vertex
numbering (first insertion), same coordinates (first inserted wins), same index list — including the degenerate
`x+y` cases where the surface runs through grid corners and many crossing points coincide.
"""
import numpy as np
import pytest

from .helpers import configure, load_meta, rel_close, same_bits
from .test_gpu_parity import CASE_NAMES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(mcb):
    c = mcb.Context(0)
    c.set_mesh_mode(mcb.MESH_SOUP | mcb.MESH_INDEXED)
    yield c
    c.close()


@pytest.mark.parametrize("name", CASE_NAMES)
def test_indexed_mesh_equals_reference_poly_data(mcb, ctx, golden, name):
    case = load_meta(golden)[name]
    configure(ctx, case)
    ctx.set_normals(1)
    cnt = ctx.polygonise()
    vref = golden[name + "/vertex_list"].reshape(-1, 3)
    tref = golden[name + "/tri_list"].reshape(-1, 3)
    assert cnt.triangles == len(tref)
    assert cnt.vertices == len(vref), "welded vertex count"
    vl, tl, vn = ctx.get_indexed_mesh(normals=True)
    assert same_bits(vl, vref), "vertex_list"
    assert np.array_equal(tl, tref.astype(np.uint32)), "tri_list"
    # the soup of the same run expands to the same triangles (first-inserted twin within the weld tolerance)
    pos, nrm = ctx.get_mesh(normals=True)
    if len(tl):
        assert rel_close(vl[tl.astype(np.int64)], pos[:, :, :3], 1e-5)
        # welded normals are the soup normals of the inserting cube's edge: same definition, so compare loosely
        ok = ~np.isnan(nrm[:, :, :3]).any(axis=2) & ~np.isnan(vn[tl.astype(np.int64)]).any(axis=2)
        d = np.abs(vn[tl.astype(np.int64)][ok] - nrm[:, :, :3][ok]).max() if ok.any() else 0.0
        # (quirk_div has poles: the blend of two huge gradients is ill-conditioned there, so it is exempt)
        assert d < 1e-3 or name.startswith("quirk"), d


def test_indexed_only_mode_and_slabs(mcb, golden):
    """MESH_INDEXED alone (no soup buffers); per-slab welds expand to the same triangles as the full grid."""
    c = mcb.Context(0)
    c.set_mesh_mode(mcb.MESH_INDEXED)
    case = load_meta(golden)["gyr78_17"]
    configure(c, case)
    cnt = c.polygonise()
    vl, tl = c.get_indexed_mesh()
    assert same_bits(vl, golden["gyr78_17/vertex_list"].reshape(-1, 3))
    assert np.array_equal(tl, golden["gyr78_17/tri_list"].reshape(-1, 3).astype(np.uint32))
    with pytest.raises(mcb.McbError):
        c.get_mesh()
    full = vl[tl.astype(np.int64)]
    parts = []
    M = cnt.M
    for r in range(3):
        k0, k1 = mcb.slab_range(M, r, 3)
        c.set_slab(k0, k1)
        c.polygonise()
        v, t = c.get_indexed_mesh()
        parts.append(v[t.astype(np.int64)])
    got = np.concatenate(parts)
    assert got.shape == full.shape and rel_close(got, full, 1e-5)
    c.close()


def test_indexed_mesh_large_sphere_properties(mcb):
    """257^3 sphere: counts of the unmodified reference (151 398 welded vertices, 302 792
    triangles), closed-surface Euler characteristic, every vertex referenced."""
    c = mcb.Context(0)
    c.set_mesh_mode(mcb.MESH_INDEXED)
    assert c.set_equation("x^2+y^2+z^2-0.49") == 0
    assert c.set_grid_step(2.0 / 256) == 257
    cnt = c.polygonise()
    assert (cnt.vertices, cnt.triangles) == (151398, 302792)
    vl, tl = c.get_indexed_mesh()
    assert np.array_equal(np.unique(tl), np.arange(cnt.vertices, dtype=np.uint32))
    e = np.sort(np.concatenate([tl[:, [0, 1]], tl[:, [1, 2]], tl[:, [2, 0]]]).astype(np.int64), axis=1)
    n_edges = len(np.unique(e[:, 0] * (1 << 32) + e[:, 1]))
    assert cnt.vertices - n_edges + cnt.triangles == 2
    c.close()


def test_python_mirror_reads_like_the_reference(mcb, golden):
    """Evaluator + Marching used as main.cpp:11-20 / drawer.cpp:785-831 use them; Poly_Data equals the reference's."""
    case = load_meta(golden)["eq8_gui"]
    evaluator = mcb.Evaluator()
    march_maker = mcb.Marching()
    assert march_maker.set_evaluator(evaluator)
    assert not evaluator.set_equation("(x(y)") and evaluator.set_equation(case["eq"])
    assert not march_maker.set_grid_step_size(0.6) and march_maker.set_grid_step_size(case["step"])
    march_maker.set_scaling_x(1.1); march_maker.set_scaling_y(1.1); march_maker.set_scaling_z(1.1)
    assert march_maker.recalculate()
    p = march_maker.get_poly_data()
    assert same_bits(p.vertex_list, golden["eq8_gui/vertex_list"].reshape(-1))
    assert np.array_equal(p.tri_list, golden["eq8_gui/tri_list"].reshape(-1))


@pytest.mark.parametrize("name", ["eq6", "gyr78"])
def test_deviating_fields_weld_within_tolerance(mcb, name):
    """The fields on which the reference keeps a few unwelded duplicates (see tests/test_cpp_dropin.py): the GPU mesh has
    the same triangles, and every indexed corner lies within the weld tolerance (1e-6 per axis) of the soup corner."""
    from oracle.refbind import EXAMPLE_EQUATIONS, GYR78
    c = mcb.Context(0)
    c.set_mesh_mode(mcb.MESH_SOUP | mcb.MESH_INDEXED)
    assert c.set_equation(GYR78 if name == "gyr78" else EXAMPLE_EQUATIONS[6]) == 0
    assert c.set_grid_step(2.0 / 256) == 257
    cnt = c.polygonise()
    vl, tl = c.get_indexed_mesh()
    pos, _ = c.get_mesh(normals=False)
    assert cnt.triangles == {"eq6": 186060, "gyr78": 2396052}[name]
    d = np.abs(vl[tl.astype(np.int64)].astype(np.float64) - pos[:, :, :3].astype(np.float64)).max()
    assert d < 1e-6, d
    c.close()


@pytest.mark.parametrize("name", CASE_NAMES)
def test_normal_h_normals_bit_exact(mcb, golden, name):
    """mcb_set_normals(2): CalculateNormal (normal.h:3-42) on the GPU — the sum per welded vertex runs in the reference's
    triangle order, so the fp32 result is the reference's bit for bit (isolated / degenerate vertices: NaN in both)."""
    case = load_meta(golden)[name]
    c = mcb.Context(0)
    c.set_mesh_mode(mcb.MESH_INDEXED)
    configure(c, case)
    c.set_normals(2)
    c.polygonise()
    vl, tl, vn = c.get_indexed_mesh(normals=True)
    ref = golden[name + "/normals"].reshape(-1, 3)
    assert vn.shape == ref.shape and same_bits(vn, ref)
    c.close()


def test_normal_h_normals_need_the_indexed_mesh(mcb):
    c = mcb.Context(0)
    c.set_normals(2)
    with pytest.raises(mcb.McbError):
        c.polygonise()
    c.close()


def test_streamed_host_output_equals_the_plain_copy(mcb):
    """mcb_set_host_output: the mesh streamed into registered (pinned) buffers while weld_emit is still running equals
    what mcb_get_indexed_mesh copies afterwards; too-small buffers and the first (buffer-growing) call fall back."""
    import torch
    c = mcb.Context(0)
    c.set_mesh_mode(mcb.MESH_INDEXED)
    c.set_normals(1)
    assert c.set_equation("x^2+y^2+z^2-0.49") == 0 and c.set_grid_step(2.0 / 256) == 257
    cnt = c.polygonise()
    v0, t0, n0 = c.get_indexed_mesh(normals=True)
    capV, capT = int(cnt.vertices) + 100, int(cnt.triangles) + 100
    bv = torch.full((capV, 3), -7.0, dtype=torch.float32).pin_memory()
    bt = torch.full((capT, 3), -7, dtype=torch.int32).pin_memory()
    bn = torch.full((capV, 3), -7.0, dtype=torch.float32).pin_memory()
    c.set_host_output(bv.data_ptr(), bt.data_ptr(), bn.data_ptr(), capV, capT)
    cnt2 = c.polygonise()
    assert c.host_output_filled() and (cnt2.vertices, cnt2.triangles) == (cnt.vertices, cnt.triangles)
    V, T = int(cnt.vertices), int(cnt.triangles)
    assert same_bits(bv.numpy()[:V], v0) and same_bits(bn.numpy()[:V], n0)
    assert np.array_equal(bt.numpy()[:T].view(np.uint32), t0)
    assert float(bv[V:].min()) == -7.0 and int(bt[T:].min()) == -7  # nothing written past the mesh
    # host buffers too small: nothing is streamed, the plain getter still works
    c.set_host_output(bv.data_ptr(), bt.data_ptr(), bn.data_ptr(), 10, 10)
    c.polygonise()
    assert not c.host_output_filled()
    v1, t1 = c.get_indexed_mesh()
    assert same_bits(v1, v0) and np.array_equal(t1, t0)
    c.set_host_output(0, 0, 0, 0, 0)
    c.close()


def test_index_base_shifts_tri_list_on_the_device(mcb):
    """mcb_set_index_base: the slab's tri_list as it sits in a mesh assembled from several slabs (Marching::set_devices) —
    shifted once on the device, idempotent over repeated reads, reset by the next polygonisation."""
    c = mcb.Context(0)
    c.set_mesh_mode(mcb.MESH_INDEXED)
    assert c.set_equation("x^2+y^2+z^2-0.49") == 0 and c.set_grid_step(2.0 / 40) > 0
    c.polygonise()
    v0, t0 = c.get_indexed_mesh()
    c.set_index_base(1000)
    for _ in range(2):
        v1, t1 = c.get_indexed_mesh()
        assert np.array_equal(t1.astype(np.int64), t0.astype(np.int64) + 1000) and np.array_equal(v0.view(np.uint32), v1.view(np.uint32))
    c.set_index_base(7)
    assert np.array_equal(c.get_indexed_mesh()[1].astype(np.int64), t0.astype(np.int64) + 7)
    c.polygonise()                                    # tri_list rewritten from zero: the base is applied to the new one
    assert np.array_equal(c.get_indexed_mesh()[1].astype(np.int64), t0.astype(np.int64) + 7)
    c.set_index_base(0)
    assert np.array_equal(c.get_indexed_mesh()[1], t0)
    c.close()
