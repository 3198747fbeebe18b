"""CPU tier, world_size 2 and 3 over gloo: the z-slab exchange (slabs.py).  Each rank polygonises its slab with the CPU
oracle (test infrastructure — there is no GPU here), the product's exchange_counts() turns the per-slab triangle counts
into global offsets, and the slabs placed at those offsets must reproduce the full-grid soup of the golden vectors."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slabs = importlib.import_module("marching-cube-for-implicit-surfaces_b200.slabs")
        from oracle import oraclebind
        from tests.helpers import load_meta
        golden = np.load(os.path.join(ROOT, "tests", "golden", "cases.npz"))
        case = load_meta(golden)[name]
        o = oraclebind.Oracle(case["eq"], case["step"], tuple(case["scale"]), case["iso"])
        k0, k1 = slabs.slab_of(case["M"], rank, world)
        sw = o.sweep(k0, k1, nthreads=1, soup=True)
        offset, total, counts = slabs.exchange_counts(sw["T"])
        assert total == case["T"] and counts[rank] == sw["T"]
        full = golden[name + "/soup"]
        ok = np.array_equal(full[offset:offset + sw["T"]].view(np.uint32), sw["soup"].view(np.uint32))
        gathered = slabs.gather_soup_to_rank0(torch.from_numpy(np.ascontiguousarray(sw["soup"])), offset, total)
        if rank == 0:
            ok = ok and np.array_equal(gathered.numpy().view(np.uint32), full.view(np.uint32))
        q.put((rank, bool(ok), offset, sw["T"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "gyr78_17"), (3, "eq8_ctor"), (2, "sphere_33_iso")])
def test_slab_offsets_reproduce_global_order(world, name):
    from oracle import oraclebind
    if not oraclebind.available():
        pytest.skip("oracle/libmcoracle.so not built")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _, _ in res)
    offs = [o for _, _, o, _ in res]
    cnts = [t for _, _, _, t in res]
    assert offs == [sum(cnts[:r]) for r in range(world)]


def _balance_worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slabs = importlib.import_module("marching-cube-for-implicit-surfaces_b200.slabs")
        from oracle import oraclebind
        from tests.helpers import load_meta
        golden = np.load(os.path.join(ROOT, "tests", "golden", "cases.npz"))
        case = load_meta(golden)[name]
        M = case["M"]
        o = oraclebind.Oracle(case["eq"], case["step"], tuple(case["scale"]), case["iso"])
        k0, k1 = slabs.slab_of(M, rank, world)
        sw = o.sweep(k0, k1, nthreads=1, soup=True)
        ntri = golden[name + "/ntri"].reshape(M, M, M).sum(axis=(1, 2))        # per layer, from the reference's own run
        mine = ntri[k0:k1]
        assert int(mine.sum()) == sw["T"]
        (b0, b1), cuts = slabs.balanced_slab(mine, k0, M, fixed_cost_per_layer=1.0)
        sw2 = o.sweep(b0, b1, nthreads=1, soup=True)
        offset, total, counts = slabs.exchange_counts(sw2["T"])
        full = golden[name + "/soup"]
        ok = total == case["T"] and np.array_equal(full[offset:offset + sw2["T"]].view(np.uint32), sw2["soup"].view(np.uint32))
        q.put((rank, bool(ok), cuts, sw["T"], sw2["T"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "sphere_33_iso"), (3, "torus_33")])
def test_balanced_slabs_keep_the_global_order_and_even_out_the_work(world, name):
    """slabs cut by measured triangles per layer (the torch.distributed twin of mcb_comm_balance): same global soup, and the
    heaviest slab is lighter than with slabs of equal thickness"""
    from oracle import oraclebind
    if not oraclebind.available():
        pytest.skip("oracle/libmcoracle.so not built")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_balance_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, ok, _, _, _ in res)
    cuts = res[0][2]
    assert all(r[2] == cuts for r in res) and cuts[0] == 0 and all(b > a for a, b in zip(cuts, cuts[1:]))
    uniform, balanced = [r[3] for r in res], [r[4] for r in res]
    assert sum(uniform) == sum(balanced) and max(balanced) <= max(uniform)


def test_balance_slabs_host_function(mcb):
    rng = np.random.default_rng(5)
    for M, w in ((9, 2), (33, 3), (1025, 8), (2049, 8), (16, 16)):
        tri = (rng.random(M) ** 4 * 1000).astype(np.uint32)
        tri[: M // 5] = 0
        tri[-(M // 7):] = 0
        cuts = mcb.balance_slabs(tri, w, 2.0)
        assert cuts[0] == 0 and cuts[-1] == M and all(b > a for a, b in zip(cuts, cuts[1:]))
        cost = [float(tri[a:b].sum()) + 2.0 * (b - a) for a, b in zip(cuts, cuts[1:])]
        ideal = sum(cost) / w
        heaviest_layer = float(tri.max()) + 2.0
        assert max(cost) <= ideal + heaviest_layer + 1e-6      # no slab overshoots by more than one layer
    # a sphere-like profile: the empty caps end up in thick slabs, the middle in thin ones
    M = 1025
    z = (np.arange(M) - M / 2) / (M / 2)
    tri = np.where(np.abs(z) < 0.7, 6000.0, 0.0).astype(np.uint32)
    cuts = mcb.balance_slabs(tri, 8, -1.0)
    thick = [b - a for a, b in zip(cuts, cuts[1:])]
    assert thick[0] > thick[3] and thick[-1] > thick[4]


def test_rebalance_slabs_converges_on_a_time_model(mcb):
    """mcb_rebalance_slabs (the host half of mcb_comm_rebalance): with slab time = launch floor + per-layer cost + per-triangle
    cost (coefficients the first cut does not know, the triangle cost depending on z as measured on the 2048^3 sphere),
    two refinements bring the slowest slab close to the mean."""
    M, w = 2049, 8
    z = (np.arange(M) + 0.5 - M / 2) / (M / 2)
    tri = np.where(np.abs(z) < 0.7, 13500.0, 0.0)
    per_tri = 0.145e-6 * (1.0 + 0.3 * (1 - np.abs(z)))          # ms per triangle: dearer near the equator
    def slab_ms(cuts):
        return [0.06 + 6e-5 * (b - a) + float((tri[a:b] * per_tri[a:b]).sum()) for a, b in zip(cuts, cuts[1:])]
    cuts = mcb.balance_slabs(tri.astype(np.uint32), w, 0.0015 * M * M)   # a deliberately bad fixed cost
    cost = tri + 0.0015 * M * M
    ms0 = slab_ms(cuts)
    spread = [max(ms0) / (sum(ms0) / w)]
    for _ in range(3):
        cost, cuts = mcb.rebalance_slabs(cost, cuts, slab_ms(cuts))
        assert cuts[0] == 0 and cuts[-1] == M and all(b > a for a, b in zip(cuts, cuts[1:]))
        ms = slab_ms(cuts)
        spread.append(max(ms) / (sum(ms) / w))
    assert spread[0] > 1.15 and spread[1] < 1.04 and spread[2] < 1.01 and spread[3] < 1.01 and max(slab_ms(cuts)) < 0.9 * max(ms0)
    # bad arguments
    with pytest.raises(mcb.McbError):
        mcb.rebalance_slabs(cost, cuts, [0.0] * w)
    with pytest.raises(mcb.McbError):
        mcb.rebalance_slabs(cost, [0, 5, 5] + cuts[3:], [1.0] * w)


def test_slab_of_matches_c_abi(mcb):
    slabs = importlib.import_module("marching-cube-for-implicit-surfaces_b200.slabs")
    for M in (9, 257, 1025, 2049):
        for w in (1, 2, 3, 4, 8):
            for r in range(w):
                assert slabs.slab_of(M, r, w) == mcb.slab_range(M, r, w)
