"""GPU check helper for the four-cells-per-lane kernel (not collected by pytest; the assertions live in tests/test_gpu_parity.py).
With rayCoherence = 2 a cell's rays are keyed on the first cell of its tile only, so the cells (j - lo, k - lo both even) of the 1-, 2- and
4-cells-per-lane kernels trace IDENTICAL rays: their divQ must agree bit for bit (same per-ray arithmetic, same accumulation order)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    from oracle import pyoracle as po
    from util import add_random_walls
    from uintah_b200 import RMCRT
    po.build()
    rng = np.random.default_rng(11)
    lv = po.burns_christon(37, dtype=np.float32)
    add_random_walls(lv, 0.02, rng, abskg=1.0, sigmaT4=float(lv.sigmaT4[5, 5, 5]))
    lo, hi = (3, 5, 2), (35, 30, 13)      # 32 cells in x: a tile of the one-cell kernel is a row
    out = {}
    for half in (False, True):
        for cpl in (1, 2, 4):
            g = RMCRT(device=0, nDivQRays=32, threshold=0.05, seed=3, rayCoherence=2, warpsPerTile=4, cellsPerLane=cpl, half_records=half)
            g.upload_level(0, lv)
            d = g.trace_divq(0, patch_lo=lo, patch_hi=hi, timestep=3)["divQ"]
            c = g.counters()
            out[(half, cpl)] = (d, int(c.nRaySteps), int(c.nDivQCells))
        d1, d2, d4 = out[(half, 1)][0], out[(half, 2)][0], out[(half, 4)][0]
        # arrays are [k, j, i] over the level incl. extra cells (index + 1)
        A = d1[lo[2] + 1:hi[2] + 1:2, lo[1] + 1:hi[1] + 1:2, lo[0] + 1:hi[0] + 1]
        B = d4[lo[2] + 1:hi[2] + 1:2, lo[1] + 1:hi[1] + 1:2, lo[0] + 1:hi[0] + 1]
        C2 = d2[lo[2] + 1:hi[2] + 1:2, lo[1] + 1:hi[1] + 1:2, lo[0] + 1:hi[0] + 1]
        print("half", half, "cells", [out[(half, c)][2] for c in (1, 2, 4)], "steps", [out[(half, c)][1] for c in (1, 2, 4)])
        print("  A-rows cpl1 vs cpl2 identical:", np.array_equal(A, C2), " cpl1 vs cpl4 identical:", np.array_equal(A, B),
              " nonzero", int((A != 0).sum()), "maxdiff", float(np.abs(A - B).max()))
        print("  written cells identical:", np.array_equal(d1 != 0, d4 != 0),
              " rel L2 all cells cpl4 vs cpl1: %.4f  cpl2 vs cpl1: %.4f" % (np.linalg.norm(d4 - d1) / np.linalg.norm(d1), np.linalg.norm(d2 - d1) / np.linalg.norm(d1)))
    # isothermal black enclosure: divQ == 0 to round-off for every cell of the 2 x 2 blocks
    M = 27
    ab = rng.uniform(0.1, 2.0, (M, M, M)).astype(np.float32); s4 = np.full((M, M, M), 0.6, np.float32); ct = np.full((M, M, M), -1, np.int32)
    for ax in range(3):
        for idx in (0, M - 1):
            sl = [slice(None)] * 3; sl[ax] = idx; ct[tuple(sl)] = 8; ab[tuple(sl)] = 1.0
    walls = (rng.uniform(size=(M, M, M)) < 0.03) & (ct == -1); ct[walls] = 8; ab[walls] = 1.0
    iso = po.Level(ab, s4, ct, (-1,) * 3, (M - 1,) * 3, (1.0 / 25,) * 3, interior_lo=(0,) * 3, interior_hi=(25,) * 3)
    for coh in (1, 2):
        g = RMCRT(device=0, nDivQRays=8, threshold=0.001, seed=5, allowReflect=0, cellsPerLane=4, rayCoherence=coh)
        g.upload_level(0, iso)
        r = g.trace_divq(0)["divQ"]
        print("isothermal coh", coh, "max |divQ| / (4 pi k s):", float(np.abs(r).max() / (4 * np.pi * 2.0 * 0.6)), "cells", int(g.counters().nDivQCells), int((ct == -1).sum()))


if __name__ == "__main__":
    main()
