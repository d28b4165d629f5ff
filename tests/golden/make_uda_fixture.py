"""Regenerate tests/golden/restartUda_RT_min.uda from the REFERENCE's own RMCRT fixture archive.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_uda_fixture.py

Source: src/StandAlone/inputs/Examples/restartUda_RT.uda (a 50^3 Arches state in 12 patches, time step 1000; used by
inputs/Examples/RMCRT_udaInit.ups through RMCRT_Test::initializeWithUda, Examples/RMCRT_Test.cc:824-940).  Only the three variables that
deck reads -- abskg, temperature, cellType -- are kept (the archive holds 26), re-encoded in the same on-disk format (index.xml,
t01000/timestep.xml, t01000/l0/pNNNNN.{xml,data}, gzip blocks) with the same 3x2x2 patch layout, so the fixture is a real UDA that the
C++ reader (uintah_b200/host/DataArchive.h) opens on the GPU box where /root/reference does not exist.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import uda_util as U  # noqa: E402

SRC = "/root/reference/src/StandAlone/inputs/Examples/restartUda_RT.uda"
DST = os.path.join(HERE, "restartUda_RT_min.uda")

fields = {}
for name in ("abskg", "temperature", "cellType"):
    a, lo, hi = U.read_level_var(SRC, name)
    fields[name] = np.ascontiguousarray(a)
lev = U.read_grid(SRC, "t01000")[0]
shutil.rmtree(DST, ignore_errors=True)
plist = U.write_uda(DST, fields, lo, hi, (0, 0, 0), (50, 50, 50), (3, 2, 2), lev["anchor"], lev["dx"], lev["extra"], timestep=1000,
                    time=0.0011495386101984409)
assert [(p[1], p[2], p[3], p[4]) for p in plist] == [(p[1], p[2], p[3], p[4]) for p in lev["patches"]], "patch layout differs from the source archive"
for name in fields:     # round trip
    b, _, _ = U.read_level_var(DST, name)
    assert np.array_equal(b, fields[name])
print("wrote", DST, sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(DST) for f in fs), "bytes")
