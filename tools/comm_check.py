"""Run under torchrun on N >= 2 GPUs: the z-slab path through the C ABI's own NCCL communicator (mcb_comm_*) against one
context polygonising the whole grid.  Every rank polygonises its slab (uniform cut, the cut balanced by measured triangles per layer, that cut refined by measured time),
the triangle counts are all-gathered by mcb_comm_exchange, and the slabs' soups placed at the offsets mcb_comm_offsets
returns must be, bit for bit, the soup of the full grid (rank 0 computes that one alone).  Prints one JSON line on rank 0."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

CUTS = ("uniform", "balanced", "rebalanced")
_stdout = os.dup(1)   # NCCL prints its version banner on stdout: everything but the JSON line goes to stderr
os.dup2(2, 1)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = importlib.import_module("marching-cube-for-implicit-surfaces_b200")
import bench
out = {"world": world, "cases": []}
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(m.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
comm_id = bytes(idt.cpu().numpy().tobytes())
ctx = m.Context(local)
ctx.set_field_mode(m.FIELD_AUTO)
first = True
for wl, n in (("sphere", 512), ("gyr78", 256), ("torus", 384)):
    eq = bench.WORKLOADS[wl]
    assert ctx.set_equation(eq) == 0
    M = ctx.set_grid_step(2.0 / n)
    ctx.set_normals(1)
    if first:
        ctx.comm_init(comm_id, rank, world)
        first = False
    res = {}
    for cut in CUTS:
        if cut == "uniform":
            k0, k1 = m.slab_range(M, rank, world)
            ctx.set_slab(k0, k1)
        elif cut == "balanced":
            k0, k1 = ctx.comm_balance()
        else:   # the balanced cut refined by the measured time of the balanced slabs (mcb_comm_rebalance)
            k0, k1 = ctx.comm_rebalance(c.ms_total)
        c = ctx.polygonise()
        ctx.comm_exchange()
        off, tot, per = ctx.comm_offsets(world)
        pos, nrm = ctx.get_mesh(normals=True)
        assert per[rank] == int(c.triangles) and off == sum(per[:rank]) and tot == sum(per)
        # gather the slabs on rank 0 at their offsets (not part of the timed path: SURVEY 8e)
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([pos.shape[0]], dtype=torch.int64, device="cuda"))
        full_pos = torch.zeros((tot, 3, 4), dtype=torch.float32, device="cuda") if rank == 0 else None
        full_nrm = torch.zeros((tot, 3, 4), dtype=torch.float32, device="cuda") if rank == 0 else None
        mine_p, mine_n = torch.from_numpy(np.ascontiguousarray(pos)).cuda(), torch.from_numpy(np.ascontiguousarray(nrm)).cuda()
        if rank == 0:
            full_pos[off:off + pos.shape[0]] = mine_p; full_nrm[off:off + pos.shape[0]] = mine_n
            o = per[0]
            for r in range(1, world):
                if per[r]:
                    dist.recv(full_pos[o:o + per[r]], src=r); dist.recv(full_nrm[o:o + per[r]], src=r)
                o += per[r]
        elif pos.shape[0]:
            dist.send(mine_p, dst=0); dist.send(mine_n, dst=0)
        res[cut] = (k0, k1, per, full_pos, full_nrm)
    if rank == 0:
        ref = m.Context(local)
        ref.set_field_mode(m.FIELD_DENSE)
        ref.set_equation(eq); ref.set_grid_step(2.0 / n); ref.set_normals(1)
        rc = ref.polygonise()
        rp, rn = ref.get_mesh(normals=True)
        ref.close()
        row = {"workload": wl, "n": n, "triangles": int(rc.triangles)}
        for cut in CUTS:
            k0, k1, per, fp, fn = res[cut]
            same_p = bool(np.array_equal(fp.cpu().numpy().view(np.uint32), rp.view(np.uint32)))
            a, b = fn.cpu().numpy(), rn
            same_n = bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.all((a == b) | (np.isnan(a) & np.isnan(b))))
            row[cut] = {"rank0_slab": [k0, k1], "per_rank_triangles": per, "soup_equals_single_gpu": same_p, "normals_equal": same_n}
            assert same_p and same_n and sum(per) == rc.triangles, (wl, cut)
        out["cases"].append(row)
    dist.barrier()
if rank == 0:
    out["ok"] = True
    os.write(_stdout, (json.dumps(out) + "\n").encode())
ctx.close()
dist.barrier()
dist.destroy_process_group()
