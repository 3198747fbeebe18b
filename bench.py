#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: Gvoxels/s and Mtriangles/s at 1024^3 / 2048^3).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference's own CPU implementation (oracle/_ref)

One "step" = one full polygonisation of the workload through the C ABI: axis tables -> field + sign planes ->
classify/scan/compact -> emit (positions + normals), inputs (bytecode, coordinates) already resident in HBM, and the
48-byte counts read back.  For N>1 each rank owns a z-slab and the step ends with the all-gather of
the per-slab triangle counts (NCCL) that gives every slab its global output offset.

Workload: sphere x^2+y^2+z^2-0.49 (BASELINE.json configs[2]); N=1 -> 1024^3 (M=1025 cubes per axis in the reference's
loop semantics); N ranks -> the grid with N times the voxels, step 2/n with n = round(1024*N^(1/3)) (N=8: 2048^3,
configs[4]), i.e. weak scaling at ~1.08e9 voxels per GPU.  "voxel" = one cube visited by the reference loop.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "marching-cube-for-implicit-surfaces_b200"

WORKLOADS = {
    "sphere": "x^2+y^2+z^2-0.49",
    "torus": "(x^2+y^2+z^2+0.25-0.0625)^2-(x^2+y^2)",
    "eq8": "(x^2+y^2-(1/16))^2+(y^2+z^2-(1/16))^2+(z^2+x^2-(1/16))^2-8*(x^2+y^2+z^2-(1/4))^2",
    # BASELINE.json configs[3]: a true gyroid is not expressible in the reference grammar (no sin/cos); this is the
    # polynomial gyroid of SURVEY.md Appendix B (Chebyshev T7/T8 in Horner form), the high-triangle-density stress field
    "gyr78": ("((x*(-7+x^2*(56+x^2*(-112+64*x^2))))*(1+y^2*(-32+y^2*(160+y^2*(-256+128*y^2)))))"
              "+((y*(-7+y^2*(56+y^2*(-112+64*y^2))))*(1+z^2*(-32+z^2*(160+z^2*(-256+128*z^2)))))"
              "+((z*(-7+z^2*(56+z^2*(-112+64*z^2))))*(1+x^2*(-32+x^2*(160+x^2*(-256+128*x^2)))))"),
}


def resolution_for(ngpus, base):
    n = int(round(base * ngpus ** (1.0 / 3.0)))
    return n + (n & 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def count_since(self, t0):
        return sum(1 for (t, _) in self.rows if t >= t0)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for (t, line) in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            inside = t0 - 0.05 <= t <= t1 + 0.05
            try:
                if inside:
                    sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            if inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take the nearest samples
            for (t, line) in self.rows[-3:]:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(eq, step, seconds_target, threads):
    """Times the UNMODIFIED reference (oracle/_ref: calculate_step + add_step_to_poly_data per cube, marching.cpp:375-383)
    on a bounded sample: the middle cube layers of the same grid, split over `threads` independent Evaluator+Marching
    pairs.  Returns (Gvoxels/s, Mtriangles/s, sample description, seconds)."""
    from oracle import refbind
    if not refbind.available():
        raise RuntimeError("oracle/_ref/libmcref.so missing: run __graft_entry__.build() where /root/reference exists")
    mcb = importlib.import_module(PKG)
    M, _ = mcb.grid_axis(step)
    # calibrate on a few rows through the middle of the grid, then size the sample for ~seconds_target of wall time
    mid_row = (M // 2) * M + M // 4
    t_probe, cubes_p, _ = refbind.timed_rows_mt(eq, step, row0=mid_row, nrows=threads * 2, nthreads=threads)
    rate = cubes_p / max(t_probe, 1e-6)
    nrows = int(max(threads * 2, min(M * M, round(seconds_target * rate / M))))
    # centre the sample on the middle layers so that it crosses the surface like the full grid does
    row0 = max(0, min(M * M - nrows, (M // 2) * M + M // 2 - nrows // 2))
    sec, cubes, tris = refbind.timed_rows_mt(eq, step, row0=row0, nrows=nrows, nthreads=threads)
    sample = "cube rows [%d,%d) (row=k*M+j, M=%d: %d cubes around the middle layers of the %d^3-cube grid), %d threads, unmodified Marching::calculate_step+add_step_to_poly_data" % (
        row0, row0 + nrows, M, cubes, M, threads)
    return cubes / sec / 1e9, tris / sec / 1e6, sample, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    eq = WORKLOADS[args.workload]
    n = resolution_for(args.gpus, args.base_res)
    step = 2.0 / n
    threads = os.cpu_count() or 1
    per_step = max(1.0, min(6.0, 120.0 / max(1, args.steps + args.warmup)))
    vals, tris = [], []
    sample = ""
    t_all = time.time()
    for i in range(args.warmup + args.steps):
        gv, mt, sample, sec = cpu_reference_rate(eq, step, per_step, threads)
        if i >= args.warmup:
            vals.append(gv); tris.append(mt)
    v = statistics.mean(vals)
    out = {"impl": "reference", "metric": "Gvoxels/s", "value": v, "unit": "Gvoxels/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "mtriangles_per_s": statistics.mean(tris),
           "config": {"workload": "%s %s at %d^3 (step 2/%d), CPU reference on a bounded sample" % (args.workload, eq, n, n),
                      "timing": "wall clock around the reference loop"},
           "cpu_baseline": {"value": v, "unit": "Gvoxels/s", "cores": threads, "kind": "reference", "sample": sample},
           "e2e": {"value": v, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.time() - t_all}
    print(json.dumps(out))
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("MCB_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mcb = importlib.import_module(PKG)
    eq = WORKLOADS[args.workload]
    n = resolution_for(world, args.base_res)
    step = 2.0 / n

    ctx = mcb.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    assert ctx.set_equation(eq) == 0
    M = ctx.set_grid_step(step)
    k0, k1 = mcb.slab_range(M, rank, world)
    ctx.set_slab(k0, k1)
    ctx.set_normals(1)
    ctx.set_field_mode(mcb.FIELD_AUTO)  # what the drop-in class uses: the field write is dropped once the surface is known to be sparse
    slabs = importlib.import_module(PKG + ".slabs")
    placement = slabs.DeviceCounts(ctx, torch.device("cuda", local)) if world > 1 else None

    def step_fn():
        c = ctx.polygonise()
        if world > 1:  # the path's only exchange: per-slab triangle counts -> global output offsets (NCCL all-gather,
            placement.exchange()  # straight from the device counters; offset/total stay on the device)
        return c

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None  # nvidia-smi takes ~0.1 s to start: it runs from the warm-up on
    for _ in range(max(3, args.warmup)):
        c = step_fn()
    sync_all()

    # ---- timed region: device-resident inputs ---------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = {"ms_tables": 0.0, "ms_eval": 0.0, "ms_classify": 0.0, "ms_emit": 0.0, "ms_fill": 0.0}
    launches = 0
    t0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        c = step_fn()
        for k in stage:
            stage[k] += getattr(c, k)
        launches += c.launches
    if world > 1:
        placement.wait(stream)  # the last exchange is inside the timed region
    e1.record(stream)
    sync_all()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    soak = 0
    if sampler and sampler.proc and world == 1:
        # a timed region shorter than the sampling period holds no sample: keep the same load running (untimed) until
        # one has been taken, so that the clocks reported are clocks under this load
        t_soak = time.time()
        while sampler.count_since(t0) < 2 and time.time() - t_soak < 1.0:
            step_fn()
            soak += 1
        torch.cuda.synchronize()
        t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    if clocks is not None:
        clocks["window"] = "timed region" if soak == 0 else "timed region + %d untimed steps of the same load after it" % soak
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(c.cubes), float(c.triangles), float(c.active), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(tms.item())
    cubes, tris, active, launches_all = [float(x) for x in tot.tolist()]
    ms_per_step = ms / args.steps
    value = cubes / (ms_per_step * 1e-3) / 1e9

    # ---- variants (informational): the field fully materialised; the bytecode interpreter ---------------------------------
    variants = {}
    if world == 1:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nv = max(1, min(args.steps, 10))
        ctx.set_field_mode(mcb.FIELD_DENSE)
        for _ in range(3):
            cs_ = ctx.polygonise()
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(nv):
            cs_ = ctx.polygonise()
        ev1.record(stream)
        torch.cuda.synchronize()
        dn_ms = ev0.elapsed_time(ev1) / nv
        V_ = (M + 1) * (M + 1) * (k1 - k0 + 1)
        variants["dense_field"] = {"value": float(cs_.cubes) / (dn_ms * 1e-3) / 1e9, "ms_per_step": dn_ms, "ms_eval": cs_.ms_eval, "steps": nv,
                                   "eval_GBps": (4.0 * V_ + V_ / 8.0) / (cs_.ms_eval * 1e-3) / 1e9,
                                   "what": "mcb_set_field_mode(MCB_FIELD_DENSE): every vertex's value written to HBM (4 B/vertex); "
                                           "eval_GBps is that kernel's real write rate"}
        # the bytecode interpreter instead of the kernel NVRTC compiled for this equation (same results bit for bit)
        ctx.set_jit(mcb.JIT_OFF)
        for _ in range(3):
            ci_ = ctx.polygonise()
        torch.cuda.synchronize()
        ev0.record(stream)
        for _ in range(nv):
            ci_ = ctx.polygonise()
        ev1.record(stream)
        torch.cuda.synchronize()
        in_ms = ev0.elapsed_time(ev1) / nv
        variants["interpreter"] = {"value": float(ci_.cubes) / (in_ms * 1e-3) / 1e9, "ms_per_step": in_ms, "ms_eval": ci_.ms_eval, "steps": nv,
                                   "what": "mcb_set_jit(MCB_JIT_OFF): eval_field_kernel interpreting the fused bytecode"}
        ctx.set_jit(mcb.JIT_AUTO)
        ctx.set_field_mode(mcb.FIELD_AUTO)
        for _ in range(2):
            ctx.polygonise()

    # ---- e2e: through the reference-facing call with HOST buffers: equation text in, Poly_Data out ----------------
    # Marching::recalculate() leaves a welded, indexed mesh in Poly_Data (vertex_list + tri_list, marching.h:26-30);
    # that is what comes back here (MCB_MESH_INDEXED, + gradient normals per vertex).  The float4 soup variant is
    # timed as well and reported next to it.
    def make_e2e(mode, normals=1):
        ctx.set_mesh_mode(mode)
        ctx.set_normals(normals)
        cc0 = ctx.polygonise()
        capT = int(cc0.triangles) + 1024
        if mode == mcb.MESH_INDEXED:
            capV = int(cc0.vertices) + 1024
            bufs = [torch.empty((capV, 3), dtype=torch.float32).pin_memory(), torch.empty((capT, 3), dtype=torch.int32).pin_memory(),
                    torch.empty((capV, 3), dtype=torch.float32).pin_memory()]
        else:
            bufs = [torch.empty((capT, 3, 4), dtype=torch.float32).pin_memory(), torch.empty((capT, 3, 4), dtype=torch.float32).pin_memory()]

        if mode == mcb.MESH_INDEXED:  # registered once: polygonise() streams the mesh into these pinned buffers
            ctx.set_host_output(bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr() if normals else 0, capV, capT)
        else:
            ctx.set_host_output(0, 0, 0, 0, 0)

        def one():
            assert ctx.set_equation(eq) == 0          # tokenise + lower + upload bytecode (H2D) + fold constants
            ctx.set_grid_step(step)                   # host coordinate loop + upload (H2D)
            ctx.set_slab(k0, k1)
            cc = step_fn()                            # MESH_INDEXED: returns when Poly_Data is in the host buffers (D2H inside)
            if mode == mcb.MESH_INDEXED:
                if not ctx.host_output_filled():      # first call after a buffer had to grow: plain copy
                    ctx.get_indexed_mesh_into(bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr() if normals else 0, capV, capT)
            else:
                ctx.get_mesh_into(bufs[0].data_ptr(), bufs[1].data_ptr(), capT)  # D2H of positions + normals
            return cc
        return one

    def time_e2e(mode, normals=1):
        one = make_e2e(mode, normals)
        for _ in range(2):
            one()
        sync_all()
        nst = max(1, min(args.steps, 10))
        w0 = time.perf_counter()
        e0.record(stream)
        for _ in range(nst):
            cc = one()
        e1.record(stream)
        sync_all()
        wall_ms = (time.perf_counter() - w0) * 1e3
        t_ms = torch.tensor([max(e0.elapsed_time(e1), wall_ms) / nst], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return float(t_ms.item()), cc, nst

    soup_ms, cc_s, e2e_steps = time_e2e(mcb.MESH_SOUP)
    pd_ms, cc_pd, _ = time_e2e(mcb.MESH_INDEXED, 0)  # exactly what Marching::recalculate() leaves: vertex_list + tri_list
    idx_ms, cc, e2e_steps = time_e2e(mcb.MESH_INDEXED)
    ctx.set_host_output(0, 0, 0, 0, 0)
    ctx.set_mesh_mode(mcb.MESH_SOUP)
    e2e_value = cubes / (idx_ms * 1e-3) / 1e9
    h2d = 4 * (mcb.lib.mcb_grid_axis(step, None, 0) + 3 + 64) + 2052 * 2 + 512
    d2h = int(cc.vertices) * 24 + int(cc.triangles) * 12 + 48
    d2h_soup = int(cc_s.triangles) * 96 + 48
    d2h_pd = int(cc_pd.vertices) * 12 + int(cc_pd.triangles) * 12 + 48
    if world > 1:  # bytes of all ranks, and a consistency check of the device-side placement
        bt = torch.tensor([float(d2h), float(d2h_soup), float(d2h_pd)], dtype=torch.float64, device="cuda")
        dist.all_reduce(bt, op=dist.ReduceOp.SUM)
        d2h, d2h_soup, d2h_pd = int(bt[0].item()), int(bt[1].item()), int(bt[2].item())
        off, total, per_rank = placement.result()
        assert total == sum(per_rank) and off == sum(per_rank[:rank]) and per_rank[rank] == int(cc.triangles), (off, total, per_rank)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        hbm = float(peaks["hbm_gbs"])
        Ml = k1 - k0
        V = (M + 1) * (M + 1) * (Ml + 1)
        A_r, T_r, C_r = float(c.active), float(c.triangles), float(c.cubes)
        per = {k: v / args.steps for k, v in stage.items()}
        sparse_run = c.field_mode == mcb.FIELD_SPARSE
        kern = {
            # algorithmic bytes per launch, SURVEY.md §8(d) / DESIGN.md §roofline
            "eval_field": {"ms": per["ms_eval"], "bytes": 4.0 * V + V / 8.0,
                           "written_bytes": (V / 8.0) if sparse_run else (4.0 * V + V / 8.0),
                           "what": "SURVEY 8(d) accounting: 4 B field + 1 bit sign per grid vertex (%s)%s" % (
                               "mcb_eval_jit: the equation's program compiled by NVRTC" if c.jit else "eval_field_kernel: bytecode interpreter",
                               "; sparse-field mode: every vertex is evaluated but only the sign bit is written, so this "
                               "algorithmic rate exceeds the HBM peak - as SURVEY 8(d) anticipates for a variant that never "
                               "writes the field; variants.dense_field.eval_GBps is the kernel that really writes 4 B/vertex" if sparse_run else "")},
            "classify+compact": {"ms": per["ms_classify"], "bytes": V / 8.0 + 12.0 * A_r,
                                 "what": "1 bit per vertex read + 12 B per active cube written (our layout; SURVEY's 4V+C accounting is in classify_scan_emit_vs_survey_bytes)"},
            "emit": {"ms": per["ms_emit"], "bytes": 12.0 * A_r + 32.0 * A_r + 96.0 * T_r,
                     "what": "12 B record + 8 corner values per active cube read, 96 B per triangle written"},
        }
        if sparse_run:
            kern["field_refill"] = {"ms": per["ms_fill"], "bytes": 2048.0 * float(c.field_blocks) + 8.0 * A_r,
                                    "what": "sparse-field mode: records read, 32x4x4-vertex blocks around the active cubes evaluated again and written (4 B/vertex)"}
        for k, d in kern.items():
            d["GBps"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else None
            d["frac_of_hbm_peak"] = d["GBps"] / hbm if d["GBps"] else None
        # the roofline object is about HBM: in a sparse-field run the evaluation writes one bit per vertex and is bound by
        # instruction issue (one add, one compare, one vote per vertex), so the dominant HBM-streaming kernel is picked among
        # the others; kernels["eval_field"] still carries its time and its SURVEY 8(d) accounting
        cands = [k for k in kern if not (sparse_run and k == "eval_field")]
        dom = max(cands, key=lambda k: kern[k]["ms"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("%s@%d" % (args.workload, n), {}).get(dom)
            except Exception:
                traffic = None
        survey_bytes = 6.0 * C_r + 40.0 * A_r + 96.0 * T_r  # SURVEY.md §8(d) classify+scan+emit accounting
        pipe_ms = per["ms_classify"] + per["ms_emit"]
        out = {
            "metric": "Gvoxels/s", "value": value, "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "mtriangles_per_s": tris / (ms_per_step * 1e-3) / 1e6,
            "config": {"workload": "%s %s at %d^3 (step 2/%d, M=%d cubes/axis), iso 0, scale 1, positions+gradient normals, field mode auto, z-slabs over %d GPU(s)" % (
                args.workload, eq, n, n, M, world), "cubes": cubes, "triangles": tris, "active_cubes": active,
                "field_mode": "sparse (MCB_FIELD_AUTO after the first run of this configuration)" if sparse_run else "dense",
                "l2": ("nothing is read back between steps; per step the sign planes (%.2f GB) and the soup (%.2f GB) alone exceed L2, "
                       "every grid vertex is re-evaluated" % ((M + 3) ** 2 * (Ml + 3) / 8e9, 96.0 * T_r / 1e9)) if sparse_run else
                      "inputs larger than L2 (field %.2f GB per GPU)" % (4.0 * (M + 3) ** 2 * (Ml + 3) / 1e9),
                "timing": "CUDA events on the launching stream, max over ranks"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["GBps"], "peak": hbm, "unit": "GB/s",
                         "frac": kern[dom]["frac_of_hbm_peak"], "traffic": traffic, "peak_source": peak_src,
                         "kernels": kern,
                         # the same evaluation with the field really written (variants.dense_field): the number to hold against the HBM peak
                         "field_write_kernel": ({"GBps": variants["dense_field"]["eval_GBps"], "frac": variants["dense_field"]["eval_GBps"] / hbm,
                                                 "ms": variants["dense_field"]["ms_eval"]} if "dense_field" in variants else None),
                         "classify_scan_emit_vs_survey_bytes": {"bytes": survey_bytes, "ms": pipe_ms,
                                                                "GBps": survey_bytes / (pipe_ms * 1e-3) / 1e9,
                                                                "frac": survey_bytes / (pipe_ms * 1e-3) / 1e9 / hbm}},
            "e2e": {"value": e2e_value, "unit": "Gvoxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": idx_ms, "steps": e2e_steps,
                    "what": "equation text in -> Poly_Data on the host (welded vertex_list + tri_list + per-vertex normals), MCB_MESH_INDEXED",
                    "poly_data_only": {"value": cubes / (pd_ms * 1e-3) / 1e9, "ms_per_step": pd_ms, "d2h_bytes_per_step": d2h_pd,
                                       "what": "same without the per-vertex normals: exactly the Poly_Data the reference's recalculate() leaves"},
                    "soup": {"value": cubes / (soup_ms * 1e-3) / 1e9, "ms_per_step": soup_ms, "d2h_bytes_per_step": d2h_soup,
                             "what": "same, float4 triangle soup + float4 normals out (MCB_MESH_SOUP)"}},
            "variants": variants,
            "gpu_launches": int(launches_all), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                threads = os.cpu_count() or 1
                gv, mt, sample, sec = cpu_reference_rate(eq, step, args.cpu_seconds, threads)
                out["cpu_baseline"] = {"value": gv, "unit": "Gvoxels/s", "cores": threads, "kind": "reference", "sample": sample,
                                       "mtriangles_per_s": mt, "seconds": sec}
            except Exception as ex:  # the oracle is test infrastructure: its absence must not fail the product bench
                out["cpu_baseline"] = {"value": None, "unit": "Gvoxels/s", "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: %s" % ex}
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sphere", choices=sorted(WORKLOADS))
    ap.add_argument("--base-res", type=int, default=1024)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
