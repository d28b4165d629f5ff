#!/usr/bin/env python
"""bench.py -- RMCRT divQ throughput on B200 (metric of BASELINE.json: ray-cell steps/s, divQ cells/s, % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays R] [--n N]

A "step" is one radiation solve (one pass of the hot path) over synthetic Burns & Christon property fields:
  N=1   BASELINE configs[1]: single level 128^3, 100 rays/cell, float properties, one B200.
  N>1   one 128^3 patch of origin cells per GPU on a (128*px, 128*py, 128*pz) level (N=8: configs[2], 256^3);
        every step each rank packs its own patch, the packed records are all-gathered over NCCL, assembled into the
        replicated level and the rank traces its own patch ("scaling": "weak": origin cells per GPU are fixed).
`value`  = executed ray-cell steps of all ranks / max-over-ranks device time, inputs resident in HBM.
`e2e`    = the same metric through the host-buffer C-ABI calls (rmcrt_upload_level + rmcrt_finalize_level +
           rmcrt_trace_divq): pinned host arrays in, host arrays out, H2D/D2H inside the timed region.
--impl reference times the CPU restatement of the reference's own Ray::rayTrace (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_STEP = 12.0      # float abskg + float sigmaT4/pi + int32 cellType
ALG_BYTES_PER_CELL = 28.0      # 12 B read + divQ and radiationVolq (2 x f64) written


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def profiled_traffic(n, rays, rec_bytes, kernel="divq"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one divq_kernel launch, from the committed ncu --set full capture of
    the same workload (profiles/traffic.json names the report it was read from); None when this workload was not profiled"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(f"{kernel}_{n}cubed_{rays}rays_rec{rec_bytes}")
        return (float(e["dram_bytes_per_launch"]), e["source"]) if e else (None, None)
    except Exception:
        return None, None


# --------------------------------------------------------------------------------------------------- synthetic fields
def burns_christon_fields(shape_cells, lo, hi, dtype=np.float32):
    """Burns & Christon benchmark field (RMCRT_Test.cc:676-679 generalised to a box of `shape_cells` cells with
    dx = 1/128... per axis length L_d = n_d*dx) on index window [lo,hi); walls: cellType 8, abskg 1, T 0."""
    n = np.array(shape_cells)
    idx = [np.arange(lo[d], hi[d]) for d in range(3)]
    f = []
    for d in range(3):
        x = (idx[d] + 0.5) / n[d]
        f.append(1.0 - 2.0 * np.abs(x - 0.5))
    ab = 0.9 * f[2][:, None, None] * f[1][None, :, None] * f[0][None, None, :] + 0.1
    wall = np.zeros(ab.shape, bool)
    for d, ax in ((0, 2), (1, 1), (2, 0)):
        w = (idx[d] < 0) | (idx[d] >= n[d])
        sl = [None, None, None]
        sl[ax] = slice(None)
        wall |= w[tuple(sl)]
    sig_over_pi = 5.67051e-8 / np.pi * 64.804 ** 4
    s4 = np.where(wall, 0.0, sig_over_pi)
    ab = np.where(wall, 1.0, ab)
    ct = np.where(wall, 8, -1).astype(np.int32)
    return ab.astype(dtype), s4.astype(dtype), ct


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------- reference arm
CPU_KIND_NOTE = {
    "reference": "; the reference's own RMCRTCommon::findRayDirection / ray_Origin / updateSumI<T> compiled from its tree (oracle/_ref/libref_rmcrt.so) "
                 "under the Ray::rayTrace divQ cell loop",
    "port": "; oracle restatement of Ray::rayTrace (oracle/_ref/libref_rmcrt.so absent)",
}


def cpu_reference_sample(n_cells_side, rays, shape_cells, slab_cells, threads, repeats=1, warm=0):
    """Time the oracle's Ray::rayTrace restatement on a z-slab of `slab_cells` planes in the middle of the level,
    split into `threads` patch tasks run on `threads` host threads (reference partitioning: one task per patch)."""
    from oracle import pyoracle as po
    po.build()
    lo = (-1, -1, -1)
    hi = tuple(s + 1 for s in shape_cells)
    ab, s4, ct = burns_christon_fields(shape_cells, lo, hi, np.float32)
    lv = po.Level(ab, s4, ct, lo, hi, tuple(1.0 / s for s in shape_cells), interior_lo=(0, 0, 0), interior_hi=tuple(shape_cells))
    p = po.default_params(nDivQRays=rays, threshold=0.05, isSeedRandom=1)      # reference semantics: allowReflect default
    z0 = shape_cells[2] // 2 - slab_cells // 2
    plo, phi = (0, 0, z0), (shape_cells[0], shape_cells[1], z0 + slab_cells)
    times, steps = [], 0
    # the reference's own findRayDirection / ray_Origin / updateSumI<T> (oracle/_ref/libref_rmcrt.so, built from the reference tree by
    # `make -C oracle ref`) under the Ray::rayTrace cell loop when that library is there, else the oracle's restatement of the same
    trace = po.ref_ray_trace if po.ref_lib() is not None else po.ray_trace
    cpu_reference_sample.kind = "reference" if po.ref_lib() is not None else "port"
    for it in range(warm + repeats):
        t0 = time.perf_counter()
        r = trace(p, lv, plo, phi, seed=1 + it, patches=(1, threads, slab_cells), threads=threads)
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
            steps = r["steps"]
    cells = shape_cells[0] * shape_cells[1] * slab_cells
    return times, steps, cells, f"z-slab [{z0},{z0 + slab_cells}) of the {shape_cells} level = {cells} origin cells x {rays} rays"


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = args.n
    world = max(env_int("WORLD_SIZE", 1), args.gpus)
    shape = tuple(n * g for g in patch_grid(world))
    rays = args.rays
    # ~1/128 of the level per step: one xy plane of origin cells
    times, steps, cells, sample = cpu_reference_sample(n, rays, shape, 1, threads, repeats=args.steps, warm=args.warmup)
    total = sum(times)
    value = steps * len(times) / total
    line = {
        "impl": "reference", "metric": "rmcrt_ray_cell_steps_per_s", "value": value, "unit": "ray-cell steps/s",
        "divq_cells_per_s": cells * len(times) / total,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, shape, rays, world),
                   "sample_per_step": sample, "semantics": "reference (allowReflect default true: zero-weight legs marched)"},
        "cpu_baseline": {"value": value, "unit": "ray-cell steps/s", "cores": threads, "kind": cpu_reference_sample.kind,
                         "sample": sample + CPU_KIND_NOTE[cpu_reference_sample.kind]},
        "e2e": {"value": value, "unit": "ray-cell steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------- our arm
def patch_grid(n_gpus):
    from uintah_b200.partition import patch_grid as pg
    return pg(n_gpus)


def workload_name(P, shape, rays, world):
    """same string in both arms (the driver compares the arms on `config.workload`)"""
    if world == 1:
        return f"Burns-Christon single level {P}^3, {rays} rays/cell, float properties, 1 B200 (BASELINE configs[1])"
    return (f"Burns-Christon single level {shape[0]}x{shape[1]}x{shape[2]}, {rays} rays/cell, one {P}^3 patch per GPU, "
            f"packed-record all-gather per solve (BASELINE configs[2] at 8 GPUs)")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from uintah_b200 import RMCRT, LevelData

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the RMCRT hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = args.n                                # patch edge = origin cells per GPU per axis
    pg = patch_grid(world)
    shape = tuple(P * g for g in pg)          # level interior
    lo, hi = (-1, -1, -1), tuple(s + 1 for s in shape)
    dx = tuple(1.0 / s for s in shape)       # unit-cube domain like every Burns & Christon deck
    from uintah_b200 import partition
    plo, phi = partition.rank_patch(rank, pg, P)
    rays = args.rays

    if args.half_records < 0:     # policy: keep the replicated level L2-resident
        args.half_records = 1 if int(np.prod([h - l for l, h in zip(lo, hi)])) * 8 > 64e6 else 0
    skip = 0 if args.reference_semantics else 1
    g = RMCRT(device=local, nDivQRays=rays, threshold=0.05, seed=1234, skipZeroWeightLegs=skip,
              rayCoherence=args.coherence, l2Persist=args.l2persist, warpsPerTile=args.warps, tileX=args.tile_x, cellsPerLane=args.cells_per_lane,
              genericMarch=args.generic_march, maxBlocksPerSM=args.max_blocks_per_sm, half_records=bool(args.half_records))
    stream = torch.cuda.current_stream()
    desc = LevelData(None, None, None, lo, hi, dx, interior_lo=(0, 0, 0), interior_hi=shape).desc()
    g.define_level(0, desc)
    g.levels[0] = LevelData(None, None, None, lo, hi, dx, interior_lo=(0, 0, 0), interior_hi=shape)

    # ---- the rank's own property box (its patch plus the extra cells it touches), resident on the device
    blo, bhi = partition.owner_box(plo, phi, shape)
    ab, s4, ct = burns_christon_fields(shape, blo, bhi, np.float32)
    d_ab, d_s4, d_ct = (torch.from_numpy(x).to(dev) for x in (ab, s4, ct))
    boxes_lo, boxes_hi = partition.all_owner_boxes(world, pg, P, shape)
    rec_bytes = 4 if args.half_records else 8
    slot = partition.slot_bytes(boxes_lo, boxes_hi, rec_bytes)
    send = torch.zeros(slot, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(slot * world, dtype=torch.uint8, device=dev) if world > 1 else None

    out_lo, out_hi = blo, bhi                 # allocateAndPut window of the patch incl. its extra cells
    oshape = (out_hi[2] - out_lo[2], out_hi[1] - out_lo[1], out_hi[0] - out_lo[0])
    d_divQ = torch.empty(oshape, dtype=torch.float64, device=dev)
    d_volq = torch.empty_like(d_divQ)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def replicate():
        """pack own box -> (all-gather) -> assemble the replicated level"""
        if world == 1:
            g.pack_level_device(0, d_ab, d_s4, d_ct, blo, bhi, stream=stream)
        else:
            g.pack_box_device(0, d_ab, d_s4, d_ct, blo, bhi, blo, bhi, send, stream=stream)
            dist.all_gather_into_tensor(gathered, send)
            g.assemble_level(0, gathered, slot, boxes_lo, boxes_hi, stream=stream)
        g.finalize_level_async(0, stream)     # device-side validation scan; no host round trip

    def device_step(ts):
        replicate()
        g.trace_divq_device(0, plo, phi, d_divQ, d_volq, None, out_lo, out_hi, timestep=ts, zero_init=True, stream=stream)

    # one validated set-up solve
    replicate()
    g.finalize_level(0, stream)
    dual = args.cells_per_lane in (2, 4) and args.coherence >= 1
    # pack + finalize + divq (+ assemble per rank box); the two-cells-per-lane solve launches the pair kernel AND the gated one-cell
    # kernel, of which one returns at once (launch_gate)
    kernels_per_step = (3 if world == 1 else 3 + world) + (1 if dual else 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        device_step(w + 1)
    barrier()
    g.reset_counters(stream)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed: K steps, each bracketed by events on the launching stream; L2 flushed between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        e0, e1, e2 = ev[k]
        e0.record(stream)
        replicate()
        e1.record(stream)
        g.trace_divq_device(0, plo, phi, d_divQ, d_volq, None, out_lo, out_hi, timestep=100 + k, zero_init=True, stream=stream)
        e2.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [e[0].elapsed_time(e[2]) for e in ev]
    trace_ms = [e[1].elapsed_time(e[2]) for e in ev]
    c = g.counters(stream)
    steps_exec, cells_exec = int(c.nRaySteps), int(c.nDivQCells)
    clocks = sampler.stop() if rank == 0 else None

    t_total = torch.tensor([sum(step_ms) * 1e-3, sum(trace_ms) * 1e-3], dtype=torch.float64, device=dev)
    work = torch.tensor([steps_exec, cells_exec], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    t_step_total, t_trace_total = (float(x) for x in t_total.cpu())
    all_steps, all_cells = (float(x) for x in work.cpu())
    value = all_steps / t_step_total

    # ---- e2e: host buffers through the reference-facing C-ABI calls, H2D and D2H inside the timed region
    lv_lo, lv_hi = lo, hi
    if world == 1:
        h_ab, h_s4, h_ct = (torch.from_numpy(x).pin_memory() for x in (ab, s4, ct))
        h_divQ = torch.empty(oshape, dtype=torch.float64).pin_memory()
        h_volq = torch.empty(oshape, dtype=torch.float64).pin_memory()
        import ctypes as C
        from uintah_b200.capi import check, i3

        def e2e_step(ts):
            check(g.ctx, g.lib.rmcrt_upload_level(g.ctx, 0, h_ab.data_ptr(), h_s4.data_ptr(), h_ct.data_ptr(), 1, i3(blo), i3(bhi),
                                                  stream.cuda_stream))
            g.finalize_level(0, stream)
            check(g.ctx, g.lib.rmcrt_trace_divq(g.ctx, 0, i3(plo), i3(phi), h_divQ.data_ptr(), h_volq.data_ptr(), None, i3(out_lo),
                                                i3(out_hi), 1, C.c_uint64(ts), stream.cuda_stream))
        h2d = (h_ab.numel() + h_s4.numel()) * 4 + h_ct.numel() * 4
        d2h = (h_divQ.numel() + h_volq.numel()) * 8
        e2e_step(1)
        g.reset_counters(stream)
        e2e_n = max(1, min(args.steps, 5))
        es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        es.record(stream)
        for k in range(e2e_n):
            e2e_step(200 + k)
        ee.record(stream)
        torch.cuda.synchronize()
        e2e_s = es.elapsed_time(ee) * 1e-3
        c2 = g.counters(stream)
        e2e_val = int(c2.nRaySteps) / e2e_s
        e2e = {"value": e2e_val, "unit": "ray-cell steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": e2e_n, "ms_per_step": 1e3 * e2e_s / e2e_n, "divq_cells_per_s": int(c2.nDivQCells) / e2e_s,
               "api": "rmcrt_upload_level + rmcrt_finalize_level + rmcrt_trace_divq (pinned host buffers)"}
    else:
        # per rank: host property box in, host divQ/radiationVolq of the patch out, every step
        h_ab, h_s4, h_ct = (torch.from_numpy(x).pin_memory() for x in (ab, s4, ct))
        h_divQ = torch.empty(oshape, dtype=torch.float64).pin_memory()
        h_volq = torch.empty(oshape, dtype=torch.float64).pin_memory()

        def e2e_step(ts):
            d_ab.copy_(h_ab, non_blocking=True); d_s4.copy_(h_s4, non_blocking=True); d_ct.copy_(h_ct, non_blocking=True)
            device_step(ts)
            h_divQ.copy_(d_divQ, non_blocking=True); h_volq.copy_(d_volq, non_blocking=True)
            stream.synchronize()
        h2d = (h_ab.numel() + h_s4.numel() + h_ct.numel()) * 4
        d2h = (h_divQ.numel() + h_volq.numel()) * 8
        e2e_step(1)
        g.reset_counters(stream)
        e2e_n = max(1, min(args.steps, 5))
        barrier()
        es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es.record(stream)
        for k in range(e2e_n):
            e2e_step(200 + k)
        ee.record(stream)
        barrier()
        tt = torch.tensor([es.elapsed_time(ee) * 1e-3], dtype=torch.float64, device=dev)
        ww = torch.tensor([float(g.counters(stream).nRaySteps), float(g.counters(stream).nDivQCells)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        e2e_s = float(tt.item())
        e2e = {"value": float(ww[0].item()) / e2e_s, "unit": "ray-cell steps/s", "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "steps": e2e_n, "ms_per_step": 1e3 * e2e_s / e2e_n,
               "divq_cells_per_s": float(ww[1].item()) / e2e_s,
               "api": "pinned host property box -> pack/all-gather/assemble/trace -> pinned host divQ, per rank"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (divq_kernel): algorithmic bytes / CUDA-event duration of the launch
    peaks, peak_src = measured_peaks()
    per_rank_steps = all_steps / world / args.steps
    per_rank_cells = all_cells / world / args.steps
    alg_bytes = ALG_BYTES_PER_STEP * per_rank_steps + ALG_BYTES_PER_CELL * per_rank_cells
    kern_s = (t_trace_total / args.steps)
    achieved = alg_bytes / kern_s / 1e9
    traffic, traffic_src = (profiled_traffic(P, rays, rec_bytes, {2: "divq_dual", 4: "divq_quad"}.get(args.cells_per_lane if args.coherence >= 1 else 1, "divq"))
                            if world == 1 else (None, None))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "traffic_source": traffic_src, "kernel": {2: "divq_dual_kernel", 4: "divq_quad_kernel"}.get(args.cells_per_lane if args.coherence >= 1 else 1, "divq_kernel") + ("<Rec4>" if args.half_records else "<Rec8>"), "kernel_ms": 1e3 * kern_s,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "12 B per executed ray-cell step + 28 B per divQ cell; the property records are L2-resident, so the "
                        "binding roofline is L2 (see roofline_l2), HBM peak shown because MEASURED_PEAKS.json has no L2 figure"}
    # L2 denominators measured live with the library's microbenchmarks (16 MiB scratch, L2-resident)
    try:
        l2_stream, _ = g.measure_read_bandwidth(16 << 20, repeats=10, pattern=0)
        l2_gather, l2_gacc = g.measure_read_bandwidth(16 << 20, repeats=20, pattern=1)
        hbm_stream, _ = g.measure_read_bandwidth(2 << 30, repeats=5, pattern=0)
        roofline_l2 = {"bound": "l2", "achieved": achieved, "peak": l2_stream, "unit": "GB/s", "frac": achieved / l2_stream,
                       "l2_sector_gather_gbs": l2_gather, "l2_gathers_per_s": l2_gacc * 1e9,
                       "steps_per_s_vs_gather_rate": (per_rank_steps / kern_s) / (l2_gacc * 1e9),
                       "hbm_read_gbs_live": hbm_stream,
                       "how": "rmcrt_measure_read_bandwidth: 16 MiB buffer, coalesced 16-B loads (peak) and per-lane hashed 8-B "
                              "gathers counted as 32-B sectors (incoherent-ray ceiling)"}
    except Exception as ex:      # noqa: BLE001
        roofline_l2 = {"error": str(ex)}

    # ---- CPU baseline: the oracle restatement of Ray::rayTrace on the host cores, bounded sample (rank 0, N=1 only)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        slab = max(1, int(round(12 * threads / 8)))          # ~10-15 s of CPU work
        times, csteps, ccells, sample = cpu_reference_sample(P, rays, shape, slab, threads, repeats=1, warm=0)
        cpu_baseline = {"value": csteps / times[0], "unit": "ray-cell steps/s", "cores": threads, "kind": cpu_reference_sample.kind,
                        "sample": sample + f"; {times[0]:.1f} s wall; {threads} patch tasks on {threads} threads" + CPU_KIND_NOTE[cpu_reference_sample.kind],
                        "divq_cells_per_s": ccells / times[0], "steps_per_ray": csteps / (ccells * rays)}

    line = {
        "metric": "rmcrt_ray_cell_steps_per_s", "value": value, "unit": "ray-cell steps/s",
        "divq_cells_per_s": all_cells / t_step_total,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step_total / args.steps,
        "trace_ms_per_step": 1e3 * t_trace_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(P, shape, rays, world),
                   "executed_steps_per_ray": all_steps / (all_cells * rays),
                   "zero_weight_legs": "marched (reference semantics)" if not skip else "skipped (exact: cold black walls, fs==0)",
                   "ray_coherence": {0: "independent per cell", 1: "warp-shared direction", 2: "warp-shared direction and in-cell offset"}[args.coherence],
                   "cells_per_lane": args.cells_per_lane if args.coherence >= 1 else 1,
                   "l2": "flushed between steps (256 MiB fill); level records (%.1f MB, %d B/cell) re-packed every step" % (np.prod([h - l for l, h in zip(lo, hi)]) * rec_bytes / 1e6, rec_bytes),
                   "timing": "CUDA events on the launching stream per step, max over ranks"},
        "roofline": roofline, "roofline_l2": roofline_l2, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": kernels_per_step * args.steps, "clocks": clocks,
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=128, help="cells per side of the per-GPU patch")
    ap.add_argument("--rays", type=int, default=100)
    ap.add_argument("--coherence", type=int, default=2, help="rayCoherence (library default 2)")
    ap.add_argument("--l2persist", type=int, default=1)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--tile-x", type=int, default=0, help="cellsPerLane 4: origin cells of a warp in x (32, 16, 8; 0 = library default)")
    ap.add_argument("--cells-per-lane", type=int, default=4, help="4 (library default): a lane marches 2 x 2 cells with one DDA; 2: a pair; 1: one cell per lane (A/B)")
    ap.add_argument("--half-records", type=int, default=-1,
                    help="1: 4-B half-packed level records (fp16 abskg|wall, fp16 sigmaT4/pi); 0: 8-B float records; "
                         "-1 (default): half-packed when the float records of the level would exceed 64 MB (beyond the L2)")
    ap.add_argument("--max-blocks-per-sm", type=int, default=0, help="cap the divQ grid (0: one tile per block)")
    ap.add_argument("--generic-march", type=int, default=0, help="1: generic march template instead of the tuned fast path (A/B)")
    ap.add_argument("--reference-semantics", action="store_true", help="march the zero-weight reflected legs like the reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
