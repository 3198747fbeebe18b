"""z-slab sharding of the grid over ranks (BASELINE.json configs[4]).

The reference emits triangles in cube-loop order with z slowest (marching.cpp:375-383), so cutting the M cube layers
into contiguous z-slabs and concatenating the slabs' outputs in rank order reproduces the single-device output
exactly.  Each rank recomputes its one-vertex halo plane (the field is analytic), so the ONLY exchange on the path is
the all-gather of one integer per rank — the slab's triangle count — whose exclusive prefix is the slab's offset in the
global triangle list.  One process per GPU; the collective goes through torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU test tier).
"""
import torch
import torch.distributed as dist


def slab_of(M, rank, world):
    """Cube layers [k0,k1) of rank `rank` — same arithmetic as mcb_slab_range in the C ABI."""
    return (M * rank) // world, (M * (rank + 1)) // world


def exchange_counts(local_triangles, device=None, group=None):
    """All-gather of the per-slab triangle counts.  Returns (offset_of_this_rank, total, counts_per_rank)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0, int(local_triangles), [int(local_triangles)]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(local_triangles)], dtype=torch.int64, device=device)
    allc = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine, group=group)
    counts = [int(x) for x in allc.tolist()]
    return sum(counts[:rank]), sum(counts), counts


def balanced_slab(layer_triangles, k0, M, fixed_cost_per_layer=-1.0, group=None, device=None):
    """Slabs of equal cost instead of equal thickness (SURVEY §8e "optionally by measured active count"): every rank
    contributes the triangles per cube layer of the slab [k0, k0 + len) it has just polygonised, the histogram of the whole
    grid is all-reduced, and every rank makes the same cut (mcb_balance_slabs).  Returns this rank's (k_begin, k_end) and the
    cut points of all ranks.  The torch.distributed twin of mcb_comm_balance (which does the same over NCCL behind the C ABI)."""
    from . import balance_slabs
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if world > 1 and dist.get_backend(group) == "nccl" else torch.device("cpu")
    hist = torch.zeros(M, dtype=torch.int64, device=device)
    lt = torch.as_tensor([int(x) for x in layer_triangles], dtype=torch.int64, device=device)
    hist[k0:k0 + len(lt)] = lt
    if world > 1:
        dist.all_reduce(hist, group=group)
    cuts = balance_slabs(hist.cpu().numpy().astype("uint32"), world, fixed_cost_per_layer)
    return (cuts[rank], cuts[rank + 1]), cuts


class DeviceCounts:
    """The same exchange without a host round trip (NCCL only): the slab's triangle count is all-gathered straight from
    the context's device counters, and the placement (offset of this slab, total) stays on the device as int64
    tensors.  The collective runs on a side stream, so the next polygonisation does not wait for the slowest rank's
    count; a consumer of `offset` / `total` calls `wait()` (stream order) or `result()` (Python ints) first.
    The context is switched to torch's current stream (mcb_set_stream): its counters are read in that stream's order.
    mcb_comm_exchange / mcb_comm_offsets do the same behind the C ABI without torch."""

    def __init__(self, ctx, device):
        # the staging copy below and the context's next counter reset must be ordered: both go on torch's current stream
        ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.device = device
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        ptr = ctx.counts_device_ptr()
        n = 5  # uint64 active, triangles, ambiguous, redirected, vertices (mcb_counts_device)

        class _Holder:  # minimal __cuda_array_interface__ view of the context's counters (no copy, no ownership)
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
        self._holder = _Holder()
        self.counters = torch.as_tensor(self._holder, device=device)
        self.all = torch.empty(self.world, dtype=torch.int64, device=device)
        self.offset = torch.zeros(1, dtype=torch.int64, device=device)
        self.total = torch.zeros(1, dtype=torch.int64, device=device)
        # the next polygonisation resets the live counters: the count is first copied aside, in stream order
        self._staged = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(2)]
        self._flip = 0
        self._side = torch.cuda.Stream(device=device)
        self._ready = torch.cuda.Event()
        self._done = torch.cuda.Event()

    def exchange(self):
        cur = torch.cuda.current_stream(self.device)
        stg = self._staged[self._flip]
        self._flip ^= 1
        stg.copy_(self.counters[1:2])
        self._ready.record(cur)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ready)
            dist.all_gather_into_tensor(self.all, stg)
            torch.sum(self.all[:self.rank], dim=0, keepdim=True, out=self.offset)
            torch.sum(self.all, dim=0, keepdim=True, out=self.total)
            self._done.record(self._side)

    def wait(self, stream=None):
        """Make `stream` (default: the current one) wait for the last exchange."""
        (stream or torch.cuda.current_stream(self.device)).wait_event(self._done)

    def result(self):
        self._done.synchronize()
        counts = [int(x) for x in self.all.tolist()]
        return sum(counts[:self.rank]), sum(counts), counts


def gather_soup_to_rank0(local_soup, offset, total, group=None):
    """Optional (NOT part of the timed path): place every slab's soup at its global offset on rank 0.
    local_soup: [T_r, 3, C] tensor.  Returns the [total, 3, C] tensor on rank 0, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_soup
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    shape_tail = tuple(local_soup.shape[1:])
    if rank == 0:
        out = torch.empty((total,) + shape_tail, dtype=local_soup.dtype, device=local_soup.device)
        out[offset:offset + local_soup.shape[0]] = local_soup
        counts = torch.empty(world, dtype=torch.int64, device=local_soup.device)
    else:
        out, counts = None, None
    sizes = torch.tensor([local_soup.shape[0]], dtype=torch.int64, device=local_soup.device)
    gathered = [torch.empty_like(sizes) for _ in range(world)] if rank == 0 else None
    dist.gather(sizes, gathered, dst=0, group=group)
    if rank == 0:
        pos = int(local_soup.shape[0]) + offset
        for r in range(1, world):
            n = int(gathered[r].item())
            if n:
                dist.recv(out[pos:pos + n], src=r, group=group)
            pos += n
        return out
    if local_soup.shape[0]:
        dist.send(local_soup.contiguous(), dst=0, group=group)
    return None
