#!/usr/bin/env python
"""At-size Monte-Carlo fixtures for the BASELINE configurations (SURVEY 8c acceptance at the sizes BASELINE.json names).

For each case a z-slab of origin cells of the FULL-size level is traced K times with independent MT19937 seeds by the CPU side:
  * single level (C2 128^3, C3 256^3): the reference's OWN per-ray routines (oracle/_ref/libref_rmcrt.so: findRayDirection,
    ray_Origin, updateSumI<T> cut out of /root/reference at build time) under the Ray::rayTrace divQ cell loop (Ray.cc:742-788);
  * Data Onion (C4, fine 256^3 / coarse 64^3, RR 4, halo 4): oracle restatement of Ray::rayTrace_dataOnion (Ray.cc:943-1274), whose
    updateSumI_ML is pinned bit for bit on the reference's (tests/test_oracle_ref.py), with the ROI of the real 128^3 fine patch;
  * C5 (256^3 benchmark 2, x-/y- mirrors, boundary fluxes, Threshold 0.001): oracle restatement of Ray::rayTrace incl. the flux loop
    (Ray.cc:645-730), pattern of inputs/Examples/RMCRT_1L_perf.ups:24-126.
What is committed per slab (float32, npz): per-cell mean and standard deviation over the K runs, and the reference's own seed-to-seed
behaviour under the acceptance statistic (leave-one-out control: fraction of cells within 3 and 5 sigma, relative L2 distance from the
mean of the others), the centre-line L2 against the Burns & Christon exact solution per seed (RMCRT.m:22-38,368-413), executed steps.
tests/test_gpu_atsize.py compares the CUDA path with these on a B200.

    python tests/golden/make_atsize_golden.py [case ...]        # ~40 min on 8 cores for all cases

Burns & Christon runs use allowReflect=false: with cold black walls the reference's second, zero-weight leg adds exact zeros
(tests/test_oracle.py::test_zero_weight_second_leg_is_a_noop), so divQ is bit-identical and the CPU time halves.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po      # noqa: E402
from util import loo_control           # noqa: E402

K = 8
THREADS = os.cpu_count() or 8


def summarise(runs):
    a = np.stack(runs)
    mu, sd = a.mean(0), a.std(0, ddof=1)
    f3, f5, l2 = loo_control(runs)
    return mu, sd, dict(loo_f3=f3, loo_f5=f5, loo_l2=l2)


def centre_lines(d, N, z_index):
    """x and y centre lines through cell N/2 (RMCRT.m:383-398) of the plane z_index of a slab array [z][y][x]"""
    mid = N // 2
    x = (np.arange(N) + 0.5) / N
    ex = po.burns_christon_exact(x)
    lx, ly = d[z_index, mid, :], d[z_index, :, mid]
    return [float(np.sqrt(np.mean((lx - ex) ** 2))), float(np.sqrt(np.mean((ly - ex) ** 2)))]


def single_level_case(name, N, slabs, level, params, tracer, centre=None, ny=None):
    """ny: only rows y < ny of every plane are traced and kept (the statistics must not see untraced cells)"""
    out, meta = {}, dict(name=name, N=N, K=K, slabs=[list(s) for s in slabs], seeds=[], steps=[], centreline_L2=[])
    for si, (z0, z1) in enumerate(slabs):
        runs, fruns = [], []
        for k in range(K):
            seed = 1000 * (k + 1) + 17 * si
            t0 = time.time()
            r = tracer(params, level, (0, 0, z0), (N, N, z1), seed=seed, patches=(1, THREADS, 1), threads=THREADS)
            d = r["divQ"][z0 + 1:z1 + 1, 1:(ny or N) + 1, 1:N + 1].copy()
            runs.append(d)
            if params.solveBoundaryFlux:
                fruns.append(r["boundFlux"][z0 + 1:z1 + 1, 1:(ny or N) + 1, 1:N + 1, :].copy())
            meta["seeds"].append(seed); meta["steps"].append(int(r["steps"]))
            if centre is not None and z0 <= centre < z1:
                meta["centreline_L2"].append(centre_lines(d, N, centre - z0))
            print(f"  {name} slab [{z0},{z1}) seed {seed}: {time.time() - t0:.1f} s, {r['steps']:.3e} steps", flush=True)
        mu, sd, ctl = summarise(runs)
        out[f"mu{si}"] = mu.astype(np.float32); out[f"sd{si}"] = sd.astype(np.float32)
        meta[f"control{si}"] = ctl
        if fruns:
            fa = np.stack(fruns)
            out[f"fmu{si}"] = fa.mean(0).astype(np.float32); out[f"fsd{si}"] = fa.std(0, ddof=1).astype(np.float32)
    return out, meta


def case_c2():
    lv = po.burns_christon(128, dtype=np.float32)
    p = po.default_params(nDivQRays=100, threshold=0.05, isSeedRandom=1, allowReflect=0)
    return single_level_case("c2_128", 128, [(64, 66), (0, 2)], lv, p, po.ref_ray_trace, centre=64)


def case_c3():
    lv = po.burns_christon(256, dtype=np.float32)
    p = po.default_params(nDivQRays=100, threshold=0.05, isSeedRandom=1, allowReflect=0)
    return single_level_case("c3_256", 256, [(128, 130), (0, 2)], lv, p, po.ref_ray_trace, centre=128)


def case_c5():
    N = 256
    lv = po.parallel_plates(N, dtype=np.float32, mirror_faces=("x-", "y-"))
    p = po.default_params(nDivQRays=100, nFluxRays=100, threshold=0.001, isSeedRandom=1, allowReflect=1, solveBoundaryFlux=1)

    def tracer(p_, lv_, lo, hi, seed, patches, threads):     # half planes: the mirror corner (x-, y-) and the z- wall
        return po.ray_trace(p_, lv_, lo, (hi[0], N // 2, hi[2]), seed=seed, patches=patches, threads=threads)
    out, meta = single_level_case("c5_256", N, [(0, 2), (128, 130)], lv, p, tracer, ny=N // 2)
    meta["y_range"] = [0, N // 2]
    return out, meta


def case_c4():
    N = 256
    fine = po.burns_christon(N, dtype=np.float32)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    p = po.default_params(nDivQRays=100, threshold=0.05, isSeedRandom=1, allowReflect=0)
    plo, phi, halo = (0, 0, 0), (128, 128, 128), (4, 4, 4)
    slabs = [(126, 128), (64, 66)]
    out, meta = {}, dict(name="c4_256_64", N=N, K=K, slabs=[list(s) for s in slabs], patch=[list(plo), list(phi)], halo=list(halo),
                         seeds=[], steps=[])
    for si, (z0, z1) in enumerate(slabs):
        runs = []
        for k in range(K):
            seed = 2000 * (k + 1) + 13 * si
            t0 = time.time()
            r = po.ray_trace_data_onion_sub(p, [coarse, fine], halo, plo, phi, (0, 0, z0), (128, 128, z1), seed=seed,
                                            parts=(1, THREADS, 1), threads=THREADS)
            runs.append(r["divQ"][z0 + 1:z1 + 1, 1:129, 1:129].copy())
            meta["seeds"].append(seed); meta["steps"].append(int(r["steps"]))
            print(f"  c4 slab [{z0},{z1}) seed {seed}: {time.time() - t0:.1f} s, {r['steps']:.3e} steps", flush=True)
        mu, sd, ctl = summarise(runs)
        out[f"mu{si}"] = mu.astype(np.float32); out[f"sd{si}"] = sd.astype(np.float32)
        meta[f"control{si}"] = ctl
    return out, meta


CASES = {"c2": case_c2, "c3": case_c3, "c4": case_c4, "c5": case_c5}

if __name__ == "__main__":
    po.build()
    assert po.ref_lib() is not None, "oracle/_ref/libref_rmcrt.so missing: run `make -C oracle ref` where /root/reference exists"
    for c in (sys.argv[1:] or list(CASES)):
        t0 = time.time()
        out, meta = CASES[c]()
        meta["generated_by"] = "tests/golden/make_atsize_golden.py"
        meta["cpu_seconds_wall"] = time.time() - t0
        np.savez_compressed(os.path.join(HERE, f"atsize_{c}.npz"), meta=json.dumps(meta), **out)
        print(c, "done in", round(time.time() - t0), "s")
