"""A/B helper (not a test): run a fixed set of solves with the library RMCRT_B200_LIB points at and save the results, so that two
builds of the same ABI (e.g. -DRMCRT_PAIR_DEPTH=1 vs 2) can be compared BIT FOR BIT:

    RMCRT_B200_LIB=uintah_b200/lib/librmcrt_b200_d1.so python tests/ab_dump.py gpurun_out/ab_d1.npz
    python tests/ab_dump.py gpurun_out/ab_d2.npz
    python tests/ab_dump.py --compare gpurun_out/ab_d1.npz gpurun_out/ab_d2.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def compare(a, b):
    A, B = np.load(a), np.load(b)
    bad = 0
    for k in A.files:
        same = np.array_equal(A[k], B[k])
        print("%-28s %s" % (k, "identical" if same else "DIFFERENT (max abs %.3e)" % np.abs(A[k].astype(np.float64) - B[k].astype(np.float64)).max()))
        bad += not same
    return bad


def main(out):
    from oracle import pyoracle as po
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from util import add_random_walls
    from uintah_b200 import RMCRT
    po.build()
    res = {}

    def run(tag, lv, patch=None, **kw):
        g = RMCRT(device=0, **kw)
        g.upload_level(0, lv)
        r = g.trace_divq(0, timestep=3, **(patch or {}))
        c = g.counters()
        res[tag + "_divQ"] = r["divQ"]
        res[tag + "_volq"] = r["radiationVolq"]
        res[tag + "_n"] = np.array([int(c.nRaySteps), int(c.nDivQCells)])

    rng = np.random.default_rng(11)
    lv = po.burns_christon(37, dtype=np.float32)
    add_random_walls(lv, 0.04, rng, abskg=1.0, sigmaT4=float(lv.sigmaT4[5, 5, 5]))
    for coh in (1, 2):
        run("bc37_walls_c%d" % coh, lv, dict(patch_lo=(3, 5, 2), patch_hi=(36, 30, 11)), nDivQRays=64, threshold=0.05, seed=3, rayCoherence=coh)
    run("bc37_walls_cpl1", lv, None, nDivQRays=16, threshold=0.05, seed=3, cellsPerLane=1)
    lv = po.burns_christon(64, dtype=np.float32)
    run("bc64", lv, None, nDivQRays=32, threshold=0.05, seed=4)
    run("bc64_half", lv, None, nDivQRays=32, threshold=0.05, seed=4, half_records=True)
    run("bc64_refsem", lv, None, nDivQRays=16, threshold=0.05, seed=4, skipZeroWeightLegs=0)
    pl = po.parallel_plates(24, dtype=np.float32, mirror_faces=("x-", "y-"))
    add_random_walls(pl, 0.02, rng, abskg=0.5, sigmaT4=0.2)
    run("plates_noreflect", pl, None, nDivQRays=32, threshold=0.01, seed=6, allowReflect=0)
    run("plates_reflect_cpl2", pl, None, nDivQRays=16, threshold=0.01, seed=6, cellsPerLane=2)
    run("plates_reflect_default", pl, None, nDivQRays=16, threshold=0.001, seed=7)
    run("bc64_cpl2", lv, None, nDivQRays=16, threshold=0.05, seed=4, cellsPerLane=2)
    np.savez_compressed(out, **res)
    print("wrote", out, {k: v.tolist() for k, v in res.items() if k.endswith("_n")})


if __name__ == "__main__":
    if sys.argv[1] == "--compare":
        sys.exit(compare(sys.argv[2], sys.argv[3]))
    main(sys.argv[1])
