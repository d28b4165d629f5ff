import sys, os, numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
from uintah_b200 import RMCRT
lv = po.burns_christon(21)
def run(tag, pinned, **kw):
    g = RMCRT(device=0, marchPrecision=1, nDivQRays=3, threshold=0.05, isSeedRandom=0, skipZeroWeightLegs=0, **kw)
    g.upload_level(0, lv)
    shape = lv.abskg.shape
    if pinned:
        d = torch.zeros(shape, dtype=torch.float64).pin_memory(); v = torch.zeros(shape, dtype=torch.float64).pin_memory()
        out = g.trace_divq(0, outputs=(d.numpy(), v.numpy(), None))
    else:
        out = g.trace_divq(0)
    print(tag, "pinned" if pinned else "pageable", "max|divQ|", float(np.abs(out["divQ"]).max()), "steps", g.counters().nRaySteps, flush=True)
for pinned in (False, True):
    run("plain", pinned)
    run("bounded", pinned, maxRayLength=0.25)
    run("lhc", pinned, rayDirSampleAlgo=1)
    run("bounded+lhc", pinned, rayDirSampleAlgo=1, maxRayLength=0.25)
