#!/usr/bin/env python
"""Throughput of the BASELINE configs that are not the bench.py headline, on ONE B200 (results go to profiles/):

  C4  RMCRT_bm1_DO pattern: Data Onion, fine 256^3 in 2x2x2 patches of 128^3, coarse 64^3 (RR 4), halo [4,4,4],
      patch_based ROI, 100 rays/cell -- the coarse level is derived on the device (coarsen_Q + computeCellType)
  C5  RMCRT_1L_perf pattern: single level 256^3, benchmark 2 (abskg 1, 1500 K), walls 1000 K, x-/y- mirrors,
      solveBoundaryFlux + allowReflect, nDivQRays = nFluxRays = 100, Threshold 0.001

    python tools/bench_configs.py [--n 256] [--rays 100] [--steps 2] [--config C4 C5]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --config C4     # C4 across 8 GPUs

One JSON line per config: executed ray-cell steps/s, divQ cells/s, ms per solve (CUDA events on the launching stream,
property records re-packed every solve, L2 flushed between solves).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def plates_fields(n, dtype=np.float32, T_gas=1500.0, T_wall=1000.0, mirrors=("x-", "y-")):
    """benchmark 2 of RMCRT_Test (abskg == 1) with the wall set-up of RMCRT_1L_perf.ups:24-90"""
    M = n + 2
    ab = np.ones((M, M, M), np.float64)
    T = np.full((M, M, M), T_gas, np.float64)
    ct = np.full((M, M, M), -1, np.int32)
    faces = {"x-": (2, 0), "x+": (2, M - 1), "y-": (1, 0), "y+": (1, M - 1), "z-": (0, 0), "z+": (0, M - 1)}
    for name in ("x-", "x+", "y-", "y+", "z-", "z+"):
        ax, idx = faces[name]
        sl = [slice(None)] * 3
        sl[ax] = idx
        ct[tuple(sl)] = 8
        T[tuple(sl)] = T_wall
        ab[tuple(sl)] = 0.0 if name in mirrors else 1.0
    s4 = 5.67051e-8 / np.pi * T ** 4
    return ab.astype(dtype), s4.astype(dtype), ct


def c4_multi_gpu(args, N, half, world, own_pg=True, coherence=-1, cells_per_lane=4):
    """BASELINE configs[3]: the Data Onion across `world` GPUs of one box (torchrun, one rank per GPU).  The fine level is sharded by
    patch: every rank packs the records of the cells it owns, ONE all-gather replicates the packed fine level (the halo and the
    coarsening both need neighbours' cells), each rank derives the coarse level on its own device and traces its own fine patch.
    Weak-scaling unit = one (N/2)^3 fine patch per GPU, like the single-level headline."""
    import torch
    import torch.distributed as dist
    from bench import burns_christon_fields
    from uintah_b200 import RMCRT, LevelData, partition
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if own_pg:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    pg = partition.patch_grid(world)
    P = N // 2
    shape = tuple(P * g for g in pg)                       # fine interior: 8 ranks -> (N, N, N)
    lo, hi = (-1,) * 3, tuple(s + 1 for s in shape)
    dxf = tuple(1.0 / s for s in shape)
    plo, phi = partition.rank_patch(rank, pg, P)
    blo, bhi = partition.owner_box(plo, phi, shape)
    ab, s4, ct = burns_christon_fields(shape, blo, bhi, np.float32)
    d_ab, d_s4, d_ct = (torch.from_numpy(x).to(dev) for x in (ab, s4, ct))
    boxes_lo, boxes_hi = partition.all_owner_boxes(world, pg, P, shape)
    rec_bytes = 4 if half else 8
    slot = partition.slot_bytes(boxes_lo, boxes_hi, rec_bytes)
    send = torch.zeros(slot, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(slot * world, dtype=torch.uint8, device=dev)
    kw = {} if coherence < 0 else {"rayCoherence": coherence, "cellsPerLane": cells_per_lane}
    g = RMCRT(device=local, nDivQRays=args.rays, threshold=0.05, haloCells=(4, 4, 4), seed=99, half_records=bool(half), **kw)
    fine = LevelData(None, None, None, lo, hi, dxf, rr=(4, 4, 4), interior_lo=(0,) * 3, interior_hi=shape)
    cshape = tuple(s // 4 for s in shape)
    coarse = LevelData(None, None, None, lo, tuple(c + 1 for c in cshape), tuple(4.0 / s for s in shape), interior_lo=(0,) * 3, interior_hi=cshape)
    g.define_level(0, coarse.desc()); g.levels[0] = coarse
    g.define_level(1, fine.desc()); g.levels[1] = fine
    oshape = (bhi[2] - blo[2], bhi[1] - blo[1], bhi[0] - blo[0])
    d_divQ = torch.empty(oshape, dtype=torch.float64, device=dev)
    d_volq = torch.empty_like(d_divQ)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def solve(ts):
        g.pack_box_device(1, d_ab, d_s4, d_ct, blo, bhi, blo, bhi, send, stream=stream)
        dist.all_gather_into_tensor(gathered, send)
        g.assemble_level(1, gathered, slot, boxes_lo, boxes_hi, stream=stream)
        g.coarsen_level_device(1, 0, stream=stream)
        g.set_level_extra_cells(0, [1.0] * 6, [0.0] * 6, [8] * 6, stream=stream)
        g.finalize_level_async(1, stream)
        g.finalize_level_async(0, stream)
        g.trace_dataonion_device(2, plo, phi, d_divQ, d_volq, None, blo, bhi, timestep=ts, zero_init=True, stream=stream)

    for w in range(3):
        solve(w + 1)
    dist.barrier(); torch.cuda.synchronize()
    g.reset_counters(stream)
    ev = []
    for k in range(args.steps):
        flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); solve(100 + k); e1.record(stream)
        ev.append((e0, e1))
    dist.barrier(); torch.cuda.synchronize()
    c = g.counters(stream)
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) * 1e-3], dtype=torch.float64, device=dev)
    wk = torch.tensor([float(c.nRaySteps), float(c.nDivQCells)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(wk, op=dist.ReduceOp.SUM)
    rec = None
    if rank == 0:
        sec = float(t.item())
        rec = ({"config": "C4", "n_gpus": world, "ray_coherence": int(g.params.rayCoherence),
                          "workload": f"Data Onion fine {shape[0]}x{shape[1]}x{shape[2]} ({world} patches of {P}^3, one per GPU) / coarse {cshape[0]}x{cshape[1]}x{cshape[2]}, RR 4, "
                                      f"halo 4, {args.rays} rays/cell; packed fine level all-gathered per solve, coarse level derived on every GPU",
                          "ray_cell_steps_per_s": float(wk[0]) / sec, "divq_cells_per_s": float(wk[1]) / sec, "ms_per_solve": 1e3 * sec / args.steps,
                          "steps_per_ray": float(wk[0]) / (float(wk[1]) * args.rays), "record_bytes": rec_bytes, "solves": args.steps,
                          "timing": "CUDA events per rank, max over ranks"})
    g.close()
    dist.barrier()
    if own_pg:
        if rec is not None:
            print(json.dumps(rec))
        dist.destroy_process_group()
    return rec


def _timed(torch, stream, flush, solve, g, steps, warm=2):
    for w in range(warm):
        solve(w + 1)
    torch.cuda.synchronize()
    g.reset_counters(stream)
    ev = []
    for k in range(steps):
        flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solve(100 + k)
        e1.record(stream)
        ev.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    return sum(ms) * 1e-3, g.counters(stream)


def _half_policy(N, half):
    return half if half >= 0 else (1 if (N + 2) ** 3 * 8 > 64e6 else 0)


def c4_single_gpu(N=256, rays=100, steps=2, half=-1, coherence=-1, cells_per_lane=4, warm=2, flush=None):
    """BASELINE configs[3] on ONE B200: all eight (N/2)^3 fine patches, coarse level derived on the device every solve"""
    import torch
    from bench import burns_christon_fields
    from uintah_b200 import RMCRT, LevelData
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream()
    half = _half_policy(N, half)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush is None else flush
    lo, hi = (-1,) * 3, (N + 1,) * 3
    P = N // 2
    patches = [((a * P, b * P, c * P), ((a + 1) * P, (b + 1) * P, (c + 1) * P)) for c in range(2) for b in range(2) for a in range(2)]
    oshape = (N + 2, N + 2, N + 2)
    d_divQ = torch.empty(oshape, dtype=torch.float64, device=dev)
    d_volq = torch.empty_like(d_divQ)
    ab, s4, ct = burns_christon_fields((N,) * 3, lo, hi, np.float32)
    d_ab, d_s4, d_ct = (torch.from_numpy(x).to(dev) for x in (ab, s4, ct))
    kw = {} if coherence < 0 else {"rayCoherence": coherence, "cellsPerLane": cells_per_lane}
    g = RMCRT(device=dev.index, nDivQRays=rays, threshold=0.05, haloCells=(4, 4, 4), seed=99, half_records=bool(half), **kw)
    fine = LevelData(None, None, None, lo, hi, (1.0 / N,) * 3, rr=(4, 4, 4), interior_lo=(0,) * 3, interior_hi=(N,) * 3)
    C = N // 4
    coarse = LevelData(None, None, None, lo, (C + 1,) * 3, (4.0 / N,) * 3, interior_lo=(0,) * 3, interior_hi=(C,) * 3)
    g.define_level(0, coarse.desc()); g.levels[0] = coarse
    g.define_level(1, fine.desc()); g.levels[1] = fine

    def solve(ts):
        g.pack_level_device(1, d_ab, d_s4, d_ct, lo, hi, stream=stream)
        g.coarsen_level_device(1, 0, stream=stream)
        g.set_level_extra_cells(0, [1.0] * 6, [0.0] * 6, [8] * 6, stream=stream)
        g.finalize_level_async(1, stream)
        g.finalize_level_async(0, stream)
        for i, (plo, phi) in enumerate(patches):
            g.trace_dataonion_device(2, plo, phi, d_divQ, d_volq, None, lo, hi, timestep=ts, zero_init=(i == 0), stream=stream)
    sec, c = _timed(torch, stream, flush, solve, g, steps, warm)
    rec = {"config": "C4", "workload": f"Data Onion fine {N}^3 (8 patches of {P}^3) / coarse {C}^3, RR 4, halo 4, {rays} rays/cell, 1 B200",
           "ray_coherence": int(g.params.rayCoherence),
           "ray_cell_steps_per_s": c.nRaySteps / sec, "divq_cells_per_s": c.nDivQCells / sec, "ms_per_solve": 1e3 * sec / steps,
           "steps_per_ray": c.nRaySteps / (c.nDivQCells * rays), "record_bytes": 4 if half else 8, "solves": steps,
           "kernels_per_solve": 5 + 8 * 3}
    g.close()
    return rec


def c5_single_gpu(N=256, rays=100, steps=2, half=-1, coherence=-1, cells_per_lane=4, warm=2, flush=None):
    """BASELINE configs[4] on ONE B200"""
    import torch
    from uintah_b200 import RMCRT, LevelData
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream()
    half = _half_policy(N, half)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush is None else flush
    lo, hi = (-1,) * 3, (N + 1,) * 3
    oshape = (N + 2, N + 2, N + 2)
    d_divQ = torch.empty(oshape, dtype=torch.float64, device=dev)
    d_volq = torch.empty_like(d_divQ)
    ab, s4, ct = plates_fields(N)
    d_ab, d_s4, d_ct = (torch.from_numpy(x).to(dev) for x in (ab, s4, ct))
    kw = {} if coherence < 0 else {"rayCoherence": coherence, "cellsPerLane": cells_per_lane}
    g = RMCRT(device=dev.index, nDivQRays=rays, nFluxRays=rays, threshold=0.001, solveBoundaryFlux=1, allowReflect=1, seed=99,
              half_records=bool(half), **kw)
    lvl = LevelData(None, None, None, lo, hi, (1.0 / N,) * 3, interior_lo=(0,) * 3, interior_hi=(N,) * 3)
    g.define_level(0, lvl.desc()); g.levels[0] = lvl
    d_flux = torch.empty(oshape + (7,), dtype=torch.float64, device=dev)

    def solve(ts):
        g.pack_level_device(0, d_ab, d_s4, d_ct, lo, hi, stream=stream)
        g.finalize_level_async(0, stream)
        g.trace_divq_device(0, (0,) * 3, (N,) * 3, d_divQ, d_volq, d_flux, lo, hi, timestep=ts, zero_init=True, stream=stream)
    sec, c = _timed(torch, stream, flush, solve, g, steps, warm)
    tot = c.nRaySteps + c.nFluxRaySteps
    rec = {"config": "C5", "workload": f"single level {N}^3, benchmark 2, walls 1000 K with x-/y- mirrors, solveBoundaryFlux + allowReflect, "
                                       f"{rays} divQ + {rays} flux rays, Threshold 0.001, 1 B200",
           "ray_coherence": int(g.params.rayCoherence),
           "ray_cell_steps_per_s": tot / sec, "divq_cells_per_s": c.nDivQCells / sec, "ms_per_solve": 1e3 * sec / steps,
           "divq_steps_per_ray": c.nRaySteps / (c.nDivQCells * rays), "flux_cells": c.nFluxCells / steps,
           "flux_steps_share": c.nFluxRaySteps / tot, "record_bytes": 4 if half else 8, "solves": steps}
    g.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--rays", type=int, default=100)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--config", nargs="+", default=["C4", "C5"])
    ap.add_argument("--half-records", type=int, default=-1)
    ap.add_argument("--coherence", type=int, default=-1, help="rayCoherence (-1: library default = 0, the reference's sampling)")
    ap.add_argument("--cells-per-lane", type=int, default=4)
    args = ap.parse_args()

    import torch
    assert torch.cuda.is_available(), "needs a GPU (no CPU fallback)"
    N = args.n
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if "C4" in args.config and world > 1:
        c4_multi_gpu(args, N, _half_policy(N, args.half_records), world, coherence=args.coherence, cells_per_lane=args.cells_per_lane)
        return
    if "C4" in args.config:
        print(json.dumps(c4_single_gpu(N, args.rays, args.steps, args.half_records, args.coherence, args.cells_per_lane)))
    if "C5" in args.config:
        print(json.dumps(c5_single_gpu(N, args.rays, args.steps, args.half_records, args.coherence, args.cells_per_lane)))


if __name__ == "__main__":
    main()
