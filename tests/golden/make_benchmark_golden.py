"""Regenerate tests/golden/rmcrt_benchmarks.json from the REFERENCE's own exact-solution tables.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_benchmark_golden.py

Sources (src/parametricStudies/postProcessTools/, read by RMCRT.m to compute the L2 error of a divQ centre line):
  bench3Exact.txt      RMCRT.m:100-135  benchmark 3 (RMCRT_Test.cc:653-691: Burns & Christon abskg, T = 1000*abskg, cold black
                       walls; deck inputs/Examples/RMCRT_bm3.ups), 243 cell-centred values along the centre line.  The table is
                       stored as emission - absorption; Ray.cc:775 computes the negative of that.
  kexpDivQ{x,y,z}.dat  RMCRT.m:202-224  RMCRT_Test benchmark 7 (RMCRT_Test.cc:736-756): abskg = 5 exp(-2x-0.1),
                       T^4 = (2x+1)*1000^4 in the table's own header comment, while the code initialises T^4 = 1e7 (2x+1):
                       divQ is linear in sigmaT4 for cold walls, so the table is 1e5 x the code's problem.
  cloudsDivQ{x,y}.dat  RMCRT.m:255-277  RMCRT_Test benchmark 8 (RMCRT_Test.cc:759-779), no rescaling.
Not usable: sinDivQ*.dat (benchmark 6 as coded does not reproduce the table with any wall model tried),
hotspotDivQ*.dat (no in-tree initialisation), b4divQCenterExact.txt (isotropic scattering, RAY_SCATTER builds only).
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/parametricStudies/postProcessTools"

out = {"source": "src/parametricStudies/postProcessTools (see make_benchmark_golden.py)"}
b3 = [float(x) for x in open(os.path.join(SRC, "bench3Exact.txt")).read().split()]
assert len(b3) == 243
out["bench3"] = {"n": 243, "stored_sign": "emission-absorption (negate for Ray.cc:775)", "values": b3}
for name, axes in (("kexp", "xyz"), ("clouds", "xy")):
    out[name] = {}
    for ax in axes:
        d = np.loadtxt(os.path.join(SRC, f"{name}DivQ{ax}.dat"))
        out[name][ax] = {"x": [float(v) for v in d[:, 0]], "divQ": [float(v) for v in d[:, 1]]}
out["kexp"]["scale_to_code_problem"] = 1e-5
with open(os.path.join(HERE, "rmcrt_benchmarks.json"), "w") as f:
    json.dump(out, f)
print("wrote", os.path.join(HERE, "rmcrt_benchmarks.json"))
