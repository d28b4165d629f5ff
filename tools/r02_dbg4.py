import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import pyoracle as po
from util import add_random_walls
from uintah_b200 import RMCRT
rng = np.random.default_rng(31)
fine = po.burns_christon(32, dtype=np.float32)
add_random_walls(fine, 0.02, rng, abskg=0.7, sigmaT4=0.3)
fine.rr = (4, 4, 4)
coarse = po.coarsen_level(fine, (4, 4, 4))
patch = ((8, 8, 8), (24, 24, 24))
p = po.default_params(nDivQRays=20, threshold=0.05, isSeedRandom=1, allowReflect=1)
runs = [po.ray_trace_data_onion(p, [coarse, fine], (4, 4, 4), patch[0], patch[1], seed=7 + s, threads=8) for s in range(3)]
sl = tuple(slice(patch[0][2 - d] + 1, patch[1][2 - d] + 1) for d in range(3))
for r in runs:
    print("oracle  volq mean", float(r["radiationVolq"][sl].mean()), "steps", r["steps"] if "steps" in r else None)
for name, kw in (("sorted", dict(blockedMarch=3)), ("roi-copy", dict(blockedMarch=0)), ("generic", dict(blockedMarch=0, genericMarch=1)),
                 ("generic-float-recs", dict(blockedMarch=0, genericMarch=1, half_records=False)), ("fp64", dict(marchPrecision=1, half_records=False)),
                 ("sorted-refsem", dict(blockedMarch=3, skipZeroWeightLegs=0)), ("generic-refsem", dict(blockedMarch=0, genericMarch=1, skipZeroWeightLegs=0))):
    kw = dict(dict(half_records=True), **kw)
    g = RMCRT(device=0, nDivQRays=20, haloCells=(4, 4, 4), seed=17, threshold=0.05, **kw)
    g.upload_level(0, coarse); g.upload_level(1, fine)
    out = g.trace_dataonion(2, patch[0], patch[1], timestep=4)
    c = g.counters()
    print(name, "volq mean", float(out["radiationVolq"][sl].mean()), "steps", int(c.nRaySteps))
    g.close()
