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


def test_slab_of_matches_c_abi(mcb):
    slabs = importlib.import_module("marching-cube-for-implicit-surfaces_b200.slabs")
    for M in (9, 257, 1025, 2049):
        for w in (1, 2, 3, 4, 8):
            for r in range(w):
                assert slabs.slab_of(M, r, w) == mcb.slab_range(M, r, w)
