#!/usr/bin/env python
"""bench.py -- RMCRT divQ throughput on B200 (metric of BASELINE.json: ray-cell steps/s, divQ cells/s, % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rays R] [--n N]

A "step" is one radiation solve (one pass of the hot path) over synthetic Burns & Christon property fields:
  N=1   BASELINE configs[1]: single level 128^3, 100 rays/cell, float properties, one B200.
  N>1   one 128^3 patch of origin cells per GPU on a (128*px, 128*py, 128*pz) level (N=8: configs[2], 256^3);
        every step each rank packs its own patch, the packed records are all-gathered over NCCL, assembled into the
        replicated level and the rank traces its own patch ("scaling": "weak": origin cells per GPU are fixed).
`value`  = executed ray-cell steps of all ranks / max-over-ranks device time, inputs resident in HBM, LIBRARY DEFAULTS: rayCoherence 0 =
           the reference's sampling (independent Philox rays per (cell, ray, timestep)).  `by_coherence` repeats the measurement for the
           opt-in batched modes (1: warp-shared direction, 2: lockstep translates, four cells per lane) so the price of the conformant
           keying is visible; `configs` (N=1) adds driver-timed records for 256^3 on one GPU, C4 (Data Onion) and C5 (boundary flux +
           reflections); at N>1 `configs.c4` times the Data Onion across the ranks.
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


def cpu_reference_sample(n_cells_side, rays, shape_cells, n_planes, threads, repeats=1, warm=0):
    """Time the reference's Ray::rayTrace divQ loop on `n_planes` xy-planes of origin cells spread evenly over z (so that rays per cell
    and steps per ray are representative of the whole level), each plane split into patch tasks run on `threads` host threads
    (reference partitioning: one task per patch)."""
    from oracle import pyoracle as po
    po.build()
    lo = (-1, -1, -1)
    hi = tuple(s + 1 for s in shape_cells)
    ab, s4, ct = burns_christon_fields(shape_cells, lo, hi, np.float32)
    lv = po.Level(ab, s4, ct, lo, hi, tuple(1.0 / s for s in shape_cells), interior_lo=(0, 0, 0), interior_hi=tuple(shape_cells))
    p = po.default_params(nDivQRays=rays, threshold=0.05, isSeedRandom=1)      # reference semantics: allowReflect default
    nz = shape_cells[2]
    n_planes = max(1, min(n_planes, nz))
    planes = sorted({min(nz - 1, int((2 * i + 1) * nz / (2 * n_planes))) for i in range(n_planes)})
    times, steps = [], 0
    # the reference's own findRayDirection / ray_Origin / updateSumI<T> (oracle/_ref/libref_rmcrt.so, built from the reference tree by
    # `make -C oracle ref`) under the Ray::rayTrace cell loop when that library is there, else the oracle's restatement of the same
    trace = po.ref_ray_trace if po.ref_lib() is not None else po.ray_trace
    cpu_reference_sample.kind = "reference" if po.ref_lib() is not None else "port"
    from concurrent.futures import ThreadPoolExecutor
    per_plane = max(1, threads // len(planes))
    for it in range(warm + repeats):
        t0 = time.perf_counter()

        def one(z):
            return trace(p, lv, (0, 0, z), (shape_cells[0], shape_cells[1], z + 1), seed=1 + it + 1000 * z, patches=(1, per_plane, 1),
                         threads=per_plane)["steps"]
        with ThreadPoolExecutor(max_workers=max(1, min(len(planes), threads // per_plane))) as ex:
            st = sum(ex.map(one, planes))
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
            steps = st
    cells = shape_cells[0] * shape_cells[1] * len(planes)
    return times, steps, cells, (f"xy-planes z in {planes} of the {shape_cells} level = {cells} origin cells x {rays} rays, "
                                 f"{len(planes) * per_plane} patch tasks on {threads} threads")


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = args.n
    world = max(env_int("WORLD_SIZE", 1), args.gpus)
    shape = tuple(n * g for g in patch_grid(world))
    rays = args.rays
    # four xy-planes of origin cells spread over z per step (1/32 of a 128^3 level, ~2 s on 16 threads)
    times, steps, cells, sample = cpu_reference_sample(n, rays, shape, 4, threads, repeats=args.steps, warm=min(args.warmup, 1))
    total = sum(times)
    value = steps * len(times) / total
    line = {
        "impl": "reference", "metric": "rmcrt_ray_cell_steps_per_s", "value": value, "unit": "ray-cell steps/s",
        "divq_cells_per_s": cells * len(times) / total,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, shape, rays, world),
                   "sample_per_step": sample, "semantics": "reference (allowReflect default true: zero-weight legs marched)",
                   "executed_steps_per_ray": steps / (cells * rays),
                   "note": "the reference marches its zero-weight second legs (SURVEY 3.5-Q2), so it executes ~2.2x the ray-cell steps per "
                           "divQ cell of the GPU arm, which skips them exactly; divq_cells_per_s compares like with like"},
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


KERNEL_OF = {0: "divq_kernel<float,{r},ML=0,FAST,F32GEN> (one cell per lane, independent rays)",
             1: "divq_quad_kernel<{r}> (warp-shared direction)", 2: "divq_quad_kernel<{r}> (lockstep translates)"}
COHERENCE_NAME = {0: "independent per (cell, ray, timestep) -- the reference's sampling", 1: "warp-shared direction",
                  2: "warp-shared direction and in-cell offset (lockstep translates)"}


class SingleLevelSolve:
    """one rank's share of a single-level solve: pack own box -> (all-gather) -> assemble -> finalize -> trace own patch"""

    def __init__(self, args, P, rays, coherence, cells_per_lane, half, world, rank, local, torch, dist):
        from uintah_b200 import RMCRT, LevelData, partition
        self.torch, self.dist, self.world = torch, dist, world
        dev = self.dev = torch.device(f"cuda:{local}")
        pg = patch_grid(world)
        self.shape = shape = tuple(P * g for g in pg)
        self.lo, self.hi = lo, hi = (-1, -1, -1), tuple(s + 1 for s in shape)
        dx = tuple(1.0 / s for s in shape)       # unit-cube domain like every Burns & Christon deck
        self.plo, self.phi = partition.rank_patch(rank, pg, P)
        if half < 0:     # policy: keep the replicated level L2-resident
            half = 1 if int(np.prod([h - l for l, h in zip(lo, hi)])) * 8 > 64e6 else 0
        self.rec_bytes = 4 if half else 8
        kw = {} if coherence < 0 else {"rayCoherence": coherence}
        if args.blocked_march >= 0:
            kw["blockedMarch"] = args.blocked_march
        skip = 0 if args.reference_semantics else 1
        self.g = g = RMCRT(device=local, nDivQRays=rays, threshold=0.05, seed=1234, skipZeroWeightLegs=skip, l2Persist=args.l2persist,
                           warpsPerTile=args.warps, tileX=args.tile_x, cellsPerLane=cells_per_lane, genericMarch=args.generic_march,
                           maxBlocksPerSM=args.max_blocks_per_sm, half_records=bool(half),
                           refillLanes=args.refill_lanes, originBrickHalo=args.origin_brick_halo, **kw)
        self.coherence = int(g.params.rayCoherence)
        self.stream = stream = torch.cuda.current_stream()
        lvl = LevelData(None, None, None, lo, hi, dx, interior_lo=(0, 0, 0), interior_hi=shape)
        g.define_level(0, lvl.desc())
        g.levels[0] = lvl
        # the rank's own property box (its patch plus the extra cells it touches), resident on the device
        self.blo, self.bhi = blo, bhi = partition.owner_box(self.plo, self.phi, shape)
        self.fields = burns_christon_fields(shape, blo, bhi, np.float32)
        self.d_props = [tuple(torch.from_numpy(x).to(dev) for x in self.fields)]
        self.boxes_lo, self.boxes_hi = partition.all_owner_boxes(world, pg, P, shape)
        self.slot = partition.slot_bytes(self.boxes_lo, self.boxes_hi, self.rec_bytes)
        self.send = torch.zeros(self.slot, dtype=torch.uint8, device=dev)
        self.gathered = torch.zeros(self.slot * world, dtype=torch.uint8, device=dev) if world > 1 else None
        self.oshape = (bhi[2] - blo[2], bhi[1] - blo[1], bhi[0] - blo[0])      # allocateAndPut window of the patch incl. its extra cells
        self.d_divQ = torch.empty(self.oshape, dtype=torch.float64, device=dev)
        self.d_volq = torch.empty_like(self.d_divQ)
        dual = self.coherence >= 1 and cells_per_lane in (2, 4)
        # pack + finalize + divq (+ assemble per rank box); the batched solves launch a second, gated kernel that returns at once
        self.kernels_per_step = (3 if world == 1 else 3 + world) + (1 if dual else 0)
        self.kernel_name = KERNEL_OF[self.coherence if cells_per_lane == 4 or self.coherence == 0 else 0].format(r="Rec4" if half else "Rec8")
        if self.coherence >= 1 and cells_per_lane == 2:
            self.kernel_name = f"divq_dual_kernel<{'Rec4' if half else 'Rec8'}>"
        if self.coherence >= 1 and cells_per_lane == 1:
            self.kernel_name = f"divq_kernel<float,{'Rec4' if half else 'Rec8'},ML=0,FAST,F32GEN> (one cell per lane, warp-shared direction)"
        self.blocked = int(g.params.blockedMarch) if self.coherence == 0 and not args.generic_march else 0
        if self.blocked:
            r = "Rec4" if half else "Rec8"
            self.kernel_name = {1: f"divq_indep_kernel<{r}> (independent rays, blocked half-packed records, per-lane refill)",
                                2: f"divq_blk_kernel<{r}> (independent rays, blocked half-packed records, one cell per lane)",
                                3: f"divq_sorted_kernel<{r}> (independent rays sorted by length per 4 x 4 x 2-cell tile, blocked half-packed records)"}[self.blocked]
            self.kernels_per_step += 1          # blk_build_kernel: the blocked copy is rebuilt after every re-pack
            if self.blocked == 3 and int(g.params.allowReflect):
                self.kernels_per_step += 1      # divq_sorted_kernel<RESORT>: launched behind the plain instantiation, the level's status words pick one

    def replicate(self, props=None):
        g, st = self.g, self.stream
        d_ab, d_s4, d_ct = props or self.d_props[0]
        if self.world == 1:
            g.pack_level_device(0, d_ab, d_s4, d_ct, self.blo, self.bhi, stream=st)
        else:
            g.pack_box_device(0, d_ab, d_s4, d_ct, self.blo, self.bhi, self.blo, self.bhi, self.send, stream=st)
            self.dist.all_gather_into_tensor(self.gathered, self.send)
            g.assemble_level(0, self.gathered, self.slot, self.boxes_lo, self.boxes_hi, stream=st)
        g.finalize_level_async(0, st)     # device-side validation scan; no host round trip

    def trace(self, ts, divQ=None, volq=None, zero_init=True):
        self.g.trace_divq_device(0, self.plo, self.phi, self.d_divQ if divQ is None else divQ, self.d_volq if volq is None else volq, None,
                                 self.blo, self.bhi, timestep=ts, zero_init=zero_init, stream=self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, steps, warmup, flush, sampler=None):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed between steps; returns per-rank sums"""
        torch, st = self.torch, self.stream
        self.replicate()
        self.g.finalize_level(0, st)      # one validated set-up solve
        for w in range(warmup):
            self.replicate(); self.trace(w + 1)
        self.barrier()
        self.g.reset_counters(st)
        if sampler:
            sampler.start()
        ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(steps)]
        self.barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            flush.fill_(k & 0xFF)
            e0, e1, e2 = ev[k]
            e0.record(st)
            self.replicate()
            e1.record(st)
            self.trace(100 + k)
            e2.record(st)
        self.barrier()
        wall = time.perf_counter() - t0
        c = self.g.counters(st)
        return (sum(e[0].elapsed_time(e[2]) for e in ev) * 1e-3, sum(e[1].elapsed_time(e[2]) for e in ev) * 1e-3,
                int(c.nRaySteps), int(c.nDivQCells), wall)

    def reduce(self, t_step, t_trace, steps, cells):
        torch, dist = self.torch, self.dist
        t_total = torch.tensor([t_step, t_trace], dtype=torch.float64, device=self.dev)
        work = torch.tensor([steps, cells], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
            dist.all_reduce(work, op=dist.ReduceOp.SUM)
        return tuple(float(x) for x in t_total.cpu()) + tuple(float(x) for x in work.cpu())

    def e2e(self, n_steps):
        """the same metric through host buffers: page-locked property arrays in, page-locked divQ / radiationVolq out, every step.
        N=1: the reference-facing C-ABI calls rmcrt_upload_level + rmcrt_finalize_level + rmcrt_trace_divq.  N>1: per rank the H2D of the
        next step's property box runs on a copy stream while this step traces (two device buffers), and the kernels write the mapped
        page-locked outputs in place (no staging buffer, no D2H after the solve)."""
        import ctypes as C
        from uintah_b200.capi import check, i3
        torch, g, st = self.torch, self.g, self.stream
        ab, s4, ct = self.fields
        h_props = tuple(torch.from_numpy(x).pin_memory() for x in (ab, s4, ct))
        h_divQ = torch.empty(self.oshape, dtype=torch.float64).pin_memory()
        h_volq = torch.empty(self.oshape, dtype=torch.float64).pin_memory()
        h2d = sum(x.numel() * 4 for x in h_props)
        d2h = (h_divQ.numel() + h_volq.numel()) * 8
        if self.world == 1:
            def step(ts):
                check(g.ctx, g.lib.rmcrt_upload_level(g.ctx, 0, h_props[0].data_ptr(), h_props[1].data_ptr(), h_props[2].data_ptr(), 1,
                                                      i3(self.blo), i3(self.bhi), st.cuda_stream))
                g.finalize_level(0, st)
                check(g.ctx, g.lib.rmcrt_trace_divq(g.ctx, 0, i3(self.plo), i3(self.phi), h_divQ.data_ptr(), h_volq.data_ptr(), None,
                                                    i3(self.blo), i3(self.bhi), 1, C.c_uint64(ts), st.cuda_stream))
            step(1)
            api = "rmcrt_upload_level + rmcrt_finalize_level + rmcrt_trace_divq (pinned host buffers, outputs written in place)"
        else:
            copy = torch.cuda.Stream()
            if len(self.d_props) == 1:
                self.d_props.append(tuple(torch.empty_like(x) for x in self.d_props[0]))
            ready = [torch.cuda.Event(), torch.cuda.Event()]      # H2D of buffer b complete
            freed = [torch.cuda.Event(), torch.cuda.Event()]      # pack of buffer b complete (it may be overwritten)

            def upload(b):
                with torch.cuda.stream(copy):
                    copy.wait_event(freed[b])
                    for d, h in zip(self.d_props[b], h_props):
                        d.copy_(h, non_blocking=True)
                    ready[b].record(copy)
            for b in (0, 1):
                freed[b].record(st)
            upload(0)
            self._k = 0

            def step(ts):
                b = self._k & 1
                upload(b ^ 1)                      # next step's inputs cross the bus while this step packs / gathers / traces
                st.wait_event(ready[b])
                self.replicate(self.d_props[b])
                freed[b].record(st)
                self.trace(ts, h_divQ, h_volq, zero_init=2)      # mapped page-locked outputs, border-only zero fill
                st.synchronize()                   # the caller owns the results of step k before step k+1 starts
                self._k += 1
            step(1)
            api = ("pinned host property box -> H2D on a copy stream (double-buffered) -> pack / all-gather / assemble / trace -> "
                   "divQ, radiationVolq written in place into mapped pinned host buffers, per rank")
        g.reset_counters(st)
        self.barrier()
        es, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es.record(st)
        for k in range(n_steps):
            step(200 + k)
        ee.record(st)
        self.barrier()
        c = g.counters(st)
        sec, _, steps, cells = self.reduce(es.elapsed_time(ee) * 1e-3, 0.0, float(c.nRaySteps), float(c.nDivQCells))
        return {"value": steps / sec, "unit": "ray-cell steps/s", "h2d_bytes_per_step": int(h2d) * self.world,
                "d2h_bytes_per_step": int(d2h) * self.world, "steps": n_steps, "ms_per_step": 1e3 * sec / n_steps,
                "divq_cells_per_s": cells / sec, "api": api}


def issue_model(kernel_key):
    """warp instructions per ray-cell step of the committed ncu capture of this kernel (profiles/traffic.json), for the issue-rate roofline"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(kernel_key)
        return e if e and "warp_instructions_per_launch" in e else None
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the RMCRT hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P, rays = args.n, args.rays
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    S = SingleLevelSolve(args, P, rays, args.coherence, args.cells_per_lane, args.half_records, world, rank, local, torch, dist)
    sampler = ClockSampler(local) if rank == 0 else None
    t_step, t_trace, n_steps, n_cells, t_wall = S.timed(args.steps, args.warmup, flush, sampler)
    clocks = sampler.stop() if rank == 0 else None
    t_step_total, t_trace_total, all_steps, all_cells = S.reduce(t_step, t_trace, n_steps, n_cells)
    value = all_steps / t_step_total
    e2e = S.e2e(max(1, min(args.steps, 5)))
    g, shape, rec_bytes = S.g, S.shape, S.rec_bytes

    # ---- the opt-in batched modes on the same workload (and 256^3 / C4 / C5 at N=1; C4 across the ranks at N>1)
    by_coherence, configs = {}, {}
    sub_steps, sub_warm = max(1, min(args.steps, 5)), 3
    if not args.no_sweeps:
        for coh in (0, 1, 2):
            if coh == S.coherence and args.cells_per_lane == 4:
                by_coherence[str(coh)] = {"ray_cell_steps_per_s": value, "ms_per_step": 1e3 * t_step_total / args.steps,
                                          "kernel": S.kernel_name, "sampling": COHERENCE_NAME[coh], "is_default": coh == 0}
                continue
            S2 = SingleLevelSolve(args, P, rays, coh, 4, args.half_records, world, rank, local, torch, dist)
            a, b, c, d, _ = S2.timed(sub_steps, sub_warm, flush)
            ts, tt, st_, ce = S2.reduce(a, b, c, d)
            by_coherence[str(coh)] = {"ray_cell_steps_per_s": st_ / ts, "ms_per_step": 1e3 * ts / sub_steps, "divq_cells_per_s": ce / ts,
                                      "kernel": S2.kernel_name, "sampling": COHERENCE_NAME[coh], "is_default": coh == 0}
            S2.g.close()
            del S2
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs as bc
        if world == 1:
            cfg_steps = max(1, min(args.steps, 2))
            CN = args.configs_n
            for coh in (0, 2):
                S3 = SingleLevelSolve(args, CN, rays, coh, 4, -1, 1, 0, local, torch, dist)
                a, b, c, d, _ = S3.timed(cfg_steps, 1, flush)
                configs[f"single_level_{CN}_1gpu_coherence{coh}"] = {
                    "workload": f"Burns-Christon single level {CN}^3, {rays} rays/cell, 1 B200 (north_star target; BASELINE configs[2] on one GPU)",
                    "ray_cell_steps_per_s": c / a, "divq_cells_per_s": d / a, "ms_per_solve": 1e3 * a / cfg_steps, "trace_ms": 1e3 * b / cfg_steps,
                    "steps_per_ray": c / (d * rays), "record_bytes": S3.rec_bytes, "kernel": S3.kernel_name, "ray_coherence": coh}
                S3.g.close()
                del S3
                torch.cuda.empty_cache()
            for coh in (0, 2):
                configs[f"c4_dataonion_{CN}_{CN // 4}_coherence{coh}"] = bc.c4_single_gpu(CN, rays, cfg_steps, -1, coh, 4, warm=1, flush=flush)
                torch.cuda.empty_cache()
            for coh in (0, 2):
                configs[f"c5_flux_reflect_{CN}_coherence{coh}"] = bc.c5_single_gpu(CN, rays, cfg_steps, -1, coh, 4, warm=1, flush=flush)
                torch.cuda.empty_cache()
            # the same solve through the C++ plugin: deck -> DataWarehouse -> Ray::rayTrace<float> task on 8 patches of 64^3 -> C ABI
            exe = os.path.join(ROOT, "uintah_b200", "host", "rmcrt_test")
            deck = os.path.join(ROOT, "tests", "decks", "bm1_1L_float_128_perf.ups")
            if os.path.exists(exe) and os.path.exists(deck) and P == 128 and rays == 100:
                try:
                    import re
                    r = subprocess.run([exe, deck, "--timesteps", "3"], capture_output=True, text=True, timeout=300)
                    used = [float(x) for x in re.findall(r"Used ([0-9.eE+-]+) milliseconds", r.stdout)]
                    whole = [float(x) for x in re.findall(r"; ([0-9.eE+-]+) ms for the whole task", r.stdout)]
                    size = [float(x) for x in re.findall(r"Size: ([0-9.eE+-]+) ray-cell steps", r.stdout)]
                    if used and size:
                        configs["cpp_plugin_128"] = {
                            "workload": "tests/decks/bm1_1L_float_128_perf.ups through rmcrt_test: Ray::rayTrace<float> task, 8 patches of 64^3, "
                                        "DataWarehouse grid variables in pooled page-locked memory (rmcrt_host_alloc)",
                            "trace_ms_per_solve": used[-1], "task_ms_per_solve": whole[-1] if whole else None,
                            "ray_cell_steps_per_s": size[-1] / (used[-1] * 1e-3), "solves": len(used),
                            "timing": "the task's own RMCRT REPORT (host clock around the trace calls, results in the DataWarehouse when it stops)"}
                except Exception as ex:      # noqa: BLE001
                    configs["cpp_plugin_128"] = {"error": str(ex)}
        elif world in (2, 4, 8):
            class A:      # the Data Onion sharded by fine patch over the ranks (BASELINE configs[3] at 8 GPUs)
                pass
            a_ = A(); a_.rays = rays; a_.steps = max(1, min(args.steps, 3))
            for coh in (0, 2):
                rec = bc.c4_multi_gpu(a_, 2 * P, 1 if (2 * P + 2) ** 3 * 8 > 64e6 else 0, world, own_pg=False, coherence=coh)
                if rank == 0:
                    configs[f"c4_dataonion_coherence{coh}"] = rec

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: algorithmic bytes / CUDA-event duration of the launch
    peaks, peak_src = measured_peaks()
    per_rank_steps = all_steps / world / args.steps
    per_rank_cells = all_cells / world / args.steps
    alg_bytes = ALG_BYTES_PER_STEP * per_rank_steps + ALG_BYTES_PER_CELL * per_rank_cells
    kern_s = (t_trace_total / args.steps)
    achieved = alg_bytes / kern_s / 1e9
    cpl_eff = args.cells_per_lane if S.coherence >= 1 else 1
    kname = {2: "divq_dual", 4: "divq_quad"}.get(cpl_eff, "divq_indep" if S.coherence == 0 else "divq")
    if S.coherence == 0 and S.blocked:
        kname = {1: "divq_indepblk", 2: "divq_blk", 3: "divq_sorted"}[S.blocked]
    kkey = kname + f"_{P}cubed_{rays}rays_rec{rec_bytes}"
    prof = issue_model(kkey) or {}
    # L2 denominators measured live with the library's microbenchmarks (L2-resident scratch buffers)
    try:
        l2_stream, _ = g.measure_read_bandwidth(16 << 20, repeats=10, pattern=0)
        l2_cg, _ = g.measure_read_bandwidth(32 << 20, repeats=10, pattern=2)
        l2_gather, l2_gacc = g.measure_read_bandwidth(16 << 20, repeats=20, pattern=1)
        hbm_stream, _ = g.measure_read_bandwidth(2 << 30, repeats=5, pattern=0)
        l2_peak = max(l2_stream, l2_cg, l2_gather)
        l2_live = {"l2_stream_ldg128_gbs": l2_stream, "l2_stream_ldcg_8deep_gbs": l2_cg, "l2_sector_gather_gbs": l2_gather,
                   "l2_gathers_per_s": l2_gacc * 1e9, "hbm_read_gbs_live": hbm_stream,
                   "how": "rmcrt_measure_read_bandwidth on L2-resident scratch: coalesced 16-B loads, eight-deep L2-only 16-B loads, and "
                          "per-lane hashed 8-B gathers counted as 32-B sectors; peak = the largest of the three"}
    except Exception as ex:      # noqa: BLE001
        l2_peak, l2_live = None, {"error": str(ex)}
    sm_hz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    sm_hz *= 1e6
    if S.coherence == 0 and l2_peak:
        # independent rays: a dependent 32-B sector gather per ray-cell step; the L1 -> L2 request path (one miss request per clock per SM)
        # and the L2 -> SM bandwidth bind (ncu: l1tex__m_l1tex2xbar_req_cycles_active 96 %, lts__throughput 77 %)
        roofline = {"bound": "l2", "achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak,
                    "traffic": prof.get("dram_bytes_per_launch"), "l2_traffic": prof.get("l2_bytes_per_launch"),
                    "frac_of_l2_peak_by_actual_sector_traffic": (prof["l2_bytes_per_launch"] / kern_s / 1e9 / l2_peak) if prof.get("l2_bytes_per_launch") else None,
                    "sector_gathers_per_s_vs_measured_gather_rate": (per_rank_steps / kern_s) / (l2_gacc * 1e9),
                    # the resource that binds: L1 -> L2 miss requests (one 32-B sector per request, one request per clock per SM).  Requests
                    # per step from the committed capture x steps/s, against the request rate the gather microbenchmark reaches live
                    "l1_miss_requests_per_step": (prof["l2_bytes_per_launch"] / 32.0 / prof["steps_per_launch"])
                                                 if prof.get("l2_bytes_per_launch") and prof.get("steps_per_launch") else None,
                    "frac_of_measured_miss_request_rate": ((prof["l2_bytes_per_launch"] / 32.0 / prof["steps_per_launch"]) * (per_rank_steps / kern_s)
                                                           / (l2_gacc * 1e9)) if prof.get("l2_bytes_per_launch") and prof.get("steps_per_launch") else None,
                    "peak_source": "measured live (rmcrt_measure_read_bandwidth, see l2_live); MEASURED_PEAKS.json has no L2 figure"}
    else:
        # batched rays: neither HBM (records are read once) nor L2 (25 %) binds; the SM issue rate does
        ipl = prof.get("warp_instructions_per_launch")
        spl = prof.get("steps_per_launch")
        ips = (ipl / spl) if ipl and spl else None
        issue_peak = 148 * 4 * sm_hz
        roofline = {"bound": "issue", "achieved": (ips * per_rank_steps / kern_s / 1e9) if ips else None, "peak": issue_peak / 1e9,
                    "unit": "G warp-instructions/s", "frac": (ips * per_rank_steps / kern_s / issue_peak) if ips else None,
                    "warp_instructions_per_step": ips, "traffic": prof.get("dram_bytes_per_launch"),
                    "algorithmic_GBps": achieved, "algorithmic_over_hbm_peak": achieved / peaks["hbm_gbs"],
                    "peak_source": "148 SMs x 4 schedulers x SM clock sampled during the run; instructions per step from the committed ncu capture"}
    roofline.update({"kernel": S.kernel_name, "kernel_ms": 1e3 * kern_s, "algorithmic_bytes_per_launch": alg_bytes,
                     "traffic_source": prof.get("source"), "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_peak_source": peak_src,
                     "note": "algorithmic bytes = 12 B per executed ray-cell step + 28 B per divQ cell (SURVEY 8d)"})

    # ---- CPU baseline: the reference's Ray::rayTrace divQ loop on the host cores, bounded sample (rank 0, N=1 only)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        planes = max(1, int(round(12 * threads / 8)))          # ~10-15 s of CPU work
        times, csteps, ccells, sample = cpu_reference_sample(P, rays, shape, planes, threads, repeats=1, warm=0)
        cpu_baseline = {"value": csteps / times[0], "unit": "ray-cell steps/s", "cores": threads, "kind": cpu_reference_sample.kind,
                        "sample": sample + f"; {times[0]:.1f} s wall" + CPU_KIND_NOTE[cpu_reference_sample.kind],
                        "divq_cells_per_s": ccells / times[0], "steps_per_ray": csteps / (ccells * rays)}

    launches = S.kernels_per_step * args.steps
    line = {
        "metric": "rmcrt_ray_cell_steps_per_s", "value": value, "unit": "ray-cell steps/s",
        "divq_cells_per_s": all_cells / t_step_total,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step_total / args.steps,
        "trace_ms_per_step": 1e3 * t_trace_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(P, shape, rays, world),
                   "executed_steps_per_ray": all_steps / (all_cells * rays),
                   "zero_weight_legs": "marched (reference semantics)" if args.reference_semantics else "skipped (exact: cold black walls, fs==0)",
                   "ray_coherence": S.coherence, "sampling": COHERENCE_NAME[S.coherence],
                   "cells_per_lane": cpl_eff,
                   "l2": "flushed between steps (256 MiB fill); level records (%.1f MB, %d B/cell) re-packed every step%s" % (
                       np.prod([h - l for l, h in zip(S.lo, S.hi)]) * rec_bytes / 1e6, rec_bytes,
                       " and their blocked half-packed copy (4 B/cell, 2 x 2 x 2 cells per sector) rebuilt" if S.blocked else ""),
                   "blocked_march": S.blocked,
                   "timing": "CUDA events on the launching stream per step, max over ranks"},
        "roofline": roofline, "l2_live": l2_live, "by_coherence": by_coherence, "configs": configs,
        "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks,
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
    ap.add_argument("--coherence", type=int, default=-1, help="rayCoherence; -1 (default) = the library default 0, the reference's sampling")
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
    ap.add_argument("--blocked-march", type=int, default=-1, help="rayCoherence 0: -1 = library default (3: divq_sorted_kernel); 0 = one-cell-per-lane kernel on the level's own records; 1 = divq_indep_kernel; 2 = divq_blk_kernel (A/B)")
    ap.add_argument("--refill-lanes", type=int, default=0, help="divq_indep_kernel: waiting lanes that trigger ray set-up (0 = library default)")
    ap.add_argument("--origin-brick-halo", type=int, default=0, help="with --generic-march 1: stage the origin tile + halo in shared memory (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs-n", type=int, default=256, help="cells per side of the configs sub-records (256^3 single level, C4, C5)")
    ap.add_argument("--no-sweeps", action="store_true", help="skip by_coherence and the configs sub-records (A/B runs, profiling)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
