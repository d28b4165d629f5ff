#!/usr/bin/env python
"""Throughput of the BASELINE configs that are not the bench.py headline, on ONE B200 (results go to profiles/):

  C4  RMCRT_bm1_DO pattern: Data Onion, fine 256^3 in 2x2x2 patches of 128^3, coarse 64^3 (RR 4), halo [4,4,4],
      patch_based ROI, 100 rays/cell -- the coarse level is derived on the device (coarsen_Q + computeCellType)
  C5  RMCRT_1L_perf pattern: single level 256^3, benchmark 2 (abskg 1, 1500 K), walls 1000 K, x-/y- mirrors,
      solveBoundaryFlux + allowReflect, nDivQRays = nFluxRays = 100, Threshold 0.001

    python tools/bench_configs.py [--n 256] [--rays 100] [--steps 2] [--config C4 C5]

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--rays", type=int, default=100)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--config", nargs="+", default=["C4", "C5"])
    ap.add_argument("--half-records", type=int, default=-1)
    args = ap.parse_args()

    import torch
    from bench import burns_christon_fields
    from uintah_b200 import RMCRT, LevelData
    assert torch.cuda.is_available(), "needs a GPU (no CPU fallback)"
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream()
    N = args.n
    half = args.half_records if args.half_records >= 0 else (1 if (N + 2) ** 3 * 8 > 64e6 else 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lo, hi = (-1,) * 3, (N + 1,) * 3
    P = N // 2
    patches = [((a * P, b * P, c * P), ((a + 1) * P, (b + 1) * P, (c + 1) * P)) for c in range(2) for b in range(2) for a in range(2)]

    def timed(solve, g):
        for w in range(2):
            solve(w + 1)
        torch.cuda.synchronize()
        g.reset_counters(stream)
        ev = []
        for k in range(args.steps):
            flush.fill_(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            solve(100 + k)
            e1.record(stream)
            ev.append((e0, e1))
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]
        return sum(ms) * 1e-3, g.counters(stream)

    out_lo, out_hi = lo, hi
    oshape = (N + 2, N + 2, N + 2)
    d_divQ = torch.empty(oshape, dtype=torch.float64, device=dev)
    d_volq = torch.empty_like(d_divQ)

    if "C4" in args.config:
        ab, s4, ct = burns_christon_fields((N,) * 3, lo, hi, np.float32)
        d_ab, d_s4, d_ct = (torch.from_numpy(x).to(dev) for x in (ab, s4, ct))
        g = RMCRT(device=0, nDivQRays=args.rays, threshold=0.05, haloCells=(4, 4, 4), seed=99, half_records=bool(half))
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
                g.trace_dataonion_device(2, plo, phi, d_divQ, d_volq, None, out_lo, out_hi, timestep=ts, zero_init=(i == 0), stream=stream)
        sec, c = timed(solve, g)
        print(json.dumps({"config": "C4", "workload": f"Data Onion fine {N}^3 (8 patches of {P}^3) / coarse {C}^3, RR 4, halo 4, {args.rays} rays/cell, 1 B200",
                          "ray_cell_steps_per_s": c.nRaySteps / sec, "divq_cells_per_s": c.nDivQCells / sec, "ms_per_solve": 1e3 * sec / args.steps,
                          "steps_per_ray": c.nRaySteps / (c.nDivQCells * args.rays), "record_bytes": 4 if half else 8, "solves": args.steps}))
        g.close()

    if "C5" in args.config:
        ab, s4, ct = plates_fields(N)
        d_ab, d_s4, d_ct = (torch.from_numpy(x).to(dev) for x in (ab, s4, ct))
        g = RMCRT(device=0, nDivQRays=args.rays, nFluxRays=args.rays, threshold=0.001, solveBoundaryFlux=1, allowReflect=1, seed=99,
                  half_records=bool(half))
        lvl = LevelData(None, None, None, lo, hi, (1.0 / N,) * 3, interior_lo=(0,) * 3, interior_hi=(N,) * 3)
        g.define_level(0, lvl.desc()); g.levels[0] = lvl
        d_flux = torch.empty(oshape + (7,), dtype=torch.float64, device=dev)

        def solve(ts):
            g.pack_level_device(0, d_ab, d_s4, d_ct, lo, hi, stream=stream)
            g.finalize_level_async(0, stream)
            g.trace_divq_device(0, (0,) * 3, (N,) * 3, d_divQ, d_volq, d_flux, out_lo, out_hi, timestep=ts, zero_init=True, stream=stream)
        sec, c = timed(solve, g)
        tot = c.nRaySteps + c.nFluxRaySteps
        print(json.dumps({"config": "C5", "workload": f"single level {N}^3, benchmark 2, walls 1000 K with x-/y- mirrors, solveBoundaryFlux + allowReflect, "
                                                      f"{args.rays} divQ + {args.rays} flux rays, Threshold 0.001, 1 B200",
                          "ray_cell_steps_per_s": tot / sec, "divq_cells_per_s": c.nDivQCells / sec, "ms_per_solve": 1e3 * sec / args.steps,
                          "divq_steps_per_ray": c.nRaySteps / (c.nDivQCells * args.rays), "flux_cells": c.nFluxCells / args.steps,
                          "flux_steps_share": c.nFluxRaySteps / tot, "record_bytes": 4 if half else 8, "solves": args.steps}))
        g.close()


if __name__ == "__main__":
    main()
