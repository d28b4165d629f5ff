"""The on-disk format either side of the path: uintah_b200/host/DataArchive.h (UDA reader + writer), RMCRT_Test's
<initalizeUsingUda> (Examples/RMCRT_Test.cc:211-310, 824-940) and the lnorm_udas comparison tool (StandAlone/tools/compute_Lnorm_udas.cc).

The fixture tests/golden/restartUda_RT_min.uda is the reference's own inputs/Examples/restartUda_RT.uda cut down to the three variables
RMCRT_udaInit.ups reads (made by tests/golden/make_uda_fixture.py); tests/uda_util.py is an independent python implementation of the format.
"""
import os
import re
import subprocess

import numpy as np
import pytest

import uda_util as U
from test_host_cpp import DECKS, HOST, edit_deck, load_dump

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "restartUda_RT_min.uda")
REF_UDA = "/root/reference/src/StandAlone/inputs/Examples/restartUda_RT.uda"
LNORM = os.path.join(HOST, "lnorm_udas")
EXE = os.path.join(HOST, "rmcrt_test")


@pytest.fixture(scope="module")
def tools():
    from uintah_b200 import build
    build.build()
    subprocess.run(["make", "-C", HOST, "-j4"], check=True, capture_output=True)
    return EXE, LNORM


def lnorm(a, b, *extra):
    r = subprocess.run([LNORM, a, b, *extra], capture_output=True, text=True)
    rows = {}
    for line in r.stdout.splitlines():
        m = re.match(r"(\S+) L-(\d+) time \S+\s+norms: L1 (\S+) L2: (\S+) Linf: (\S+) n_cells: (\d+)", line)
        if m:
            rows[(m.group(1), int(m.group(2)))] = (float(m.group(3)), float(m.group(4)), float(m.group(5)), int(m.group(6)))
    return r, rows


def test_fixture_is_the_reference_archive(tools):
    """pins the fixture (and the C++ reader, which here opens the reference's own 26-variable archive) on the reference tree"""
    if not os.path.isdir(REF_UDA):
        pytest.skip("reference tree not present (GPU box)")
    r, rows = lnorm(FIXTURE, REF_UDA)
    assert r.returncode == 0, r.stderr
    assert set(rows) == {("abskg", 0), ("temperature", 0), ("cellType", 0)}
    assert all(v == (0.0, 0.0, 0.0, 125000) for v in rows.values())


def test_cpp_reader_agrees_with_python_across_layouts(tools, tmp_path):
    """same fields re-written by the python writer with another patch layout and without compression: all norms are zero"""
    fields = {n: np.ascontiguousarray(U.read_level_var(FIXTURE, n)[0]) for n in ("abskg", "temperature", "cellType")}
    assert fields["abskg"].shape == (52, 52, 52) and fields["abskg"].dtype == np.float32 and fields["cellType"].dtype == np.int32
    lev = U.read_grid(FIXTURE, "t01000")[0]
    other = str(tmp_path / "relayout.uda")
    U.write_uda(other, fields, (-1,) * 3, (51,) * 3, (0,) * 3, (50,) * 3, (2, 5, 1), lev["anchor"], lev["dx"], timestep=1000,
                time=0.0011495386101984409, gzip=False)
    r, rows = lnorm(FIXTURE, other)
    assert r.returncode == 0, r.stderr
    assert all(v == (0.0, 0.0, 0.0, 125000) for v in rows.values()) and len(rows) == 3


@pytest.mark.parametrize("nprocs,gridxml", [(1, False), (3, False), (4, True)])
def test_files_are_named_after_the_writing_rank(tools, tmp_path, nprocs, gridxml):
    """lL/p<rank>.{xml,data} hold every patch a rank owned (DataArchive.cc:2310-2314): archives written by fewer ranks than patches --
    1 rank / 10 patches (everything in p00000), 3 ranks, 4 ranks with <Grid> in grid.xml (newer archives, DataArchive.cc:431-457)"""
    fields = {n: np.ascontiguousarray(U.read_level_var(FIXTURE, n)[0]) for n in ("abskg", "temperature", "cellType")}
    lev = U.read_grid(FIXTURE, "t01000")[0]
    other = str(tmp_path / f"ranks{nprocs}.uda")
    U.write_uda(other, fields, (-1,) * 3, (51,) * 3, (0,) * 3, (50,) * 3, (2, 5, 1), lev["anchor"], lev["dx"], timestep=1000,
                time=0.0011495386101984409, gzip=True, nprocs=nprocs)
    assert sorted(os.listdir(os.path.join(other, "t01000", "l0"))) == sorted(f"p{r:05d}.{e}" for r in range(nprocs) for e in ("xml", "data"))
    if gridxml:       # move <Grid> out of timestep.xml
        tsx = os.path.join(other, "t01000", "timestep.xml")
        x = open(tsx).read()
        a, b = x.index("  <Grid>"), x.index("</Grid>") + len("</Grid>\n")
        open(os.path.join(other, "t01000", "grid.xml"), "w").write('<?xml version="1.0" encoding="UTF-8"?>\n<Grid>' + x[a + len("  <Grid>"):b])
        open(tsx, "w").write(x[:a] + x[b:])
    r, rows = lnorm(FIXTURE, other)
    assert r.returncode == 0, r.stderr
    assert all(v == (0.0, 0.0, 0.0, 125000) for v in rows.values()) and len(rows) == 3


def test_lnorm_known_answer(tools, tmp_path):
    """L1 = sum|d|/n, L2 = sqrt(sum d^2/n), Linf = max|d| over interior cells (compute_Lnorm_udas.cc:208-214, 332-335)"""
    rng = np.random.default_rng(5)
    N = 6
    a = rng.normal(size=(N + 2,) * 3)
    d = np.zeros_like(a)
    d[1:-1, 1:-1, 1:-1] = rng.normal(size=(N,) * 3)
    d[0, 0, 0] = 100.0                        # extra cells are not compared
    s7a = rng.normal(size=(N + 2,) * 3 + (7,))   # Stencil7 is not writable by the python helper: use double fields only
    del s7a
    geo = dict(lo=(-1,) * 3, hi=(N + 1,) * 3, interior_lo=(0,) * 3, interior_hi=(N,) * 3, anchor=(0, 0, 0), dx=(0.1,) * 3)
    U.write_uda(str(tmp_path / "a.uda"), {"divQ": a, "n": np.arange(a.size, dtype=np.int32).reshape(a.shape)}, patches=(1, 2, 1), **geo)
    U.write_uda(str(tmp_path / "b.uda"), {"divQ": a + d, "n": np.arange(a.size, dtype=np.int32).reshape(a.shape) + 2}, patches=(3, 1, 2), **geo)
    r, rows = lnorm(str(tmp_path / "a.uda"), str(tmp_path / "b.uda"), "--out", str(tmp_path / "Lnorm"))
    assert r.returncode == 0, r.stderr
    di = np.abs(d[1:-1, 1:-1, 1:-1])
    L1, L2, Linf, n = rows[("divQ", 0)]
    assert n == N ** 3
    np.testing.assert_allclose([L1, L2, Linf], [di.mean(), np.sqrt((di ** 2).mean()), di.max()], rtol=1e-9)
    assert rows[("n", 0)] == (2.0, 2.0, 2.0, N ** 3)
    out = open(tmp_path / "Lnorm" / "L-0" / "divQ_0").read().split()
    np.testing.assert_allclose([float(x) for x in out[1:]], [L1, L2, Linf], rtol=1e-9)
    # different output times cannot be compared
    U.write_uda(str(tmp_path / "c.uda"), {"divQ": a}, patches=(1, 1, 1), time=0.5, **geo)
    r, _ = lnorm(str(tmp_path / "a.uda"), str(tmp_path / "c.uda"))
    assert r.returncode == 5 and "differs from the output time" in r.stderr


@pytest.mark.parametrize("edit,needle", [
    (("<timestep>            1000        </timestep>", "<timestep> 7 </timestep>"), "The timestep (7) was not found in the uda"),
    (("<abskg_varName>       abskg       </abskg_varName>", "<abskg_varName> absk_gas </abskg_varName>"), "was not found in the uda"),
    (("<patches>    [3,2,2] </patches>", "<patches> [2,2,2] </patches>"), "is not equal to the initialization uda's grid"),
    (("<resolution> [50,50,50] </resolution>", "<resolution> [48,50,50] </resolution>"), "is not equal to the initialization uda's grid"),
])
def test_uda_init_input_errors(tools, tmp_path, edit, needle):
    r = subprocess.run([EXE, edit_deck(tmp_path, "udaInit_1L.ups", edit), "--init-uda", FIXTURE, "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "ProblemSetupException" in r.stderr and needle in r.stderr, r.stderr


def test_uda_init_schedule(tools):
    r = subprocess.run([EXE, os.path.join(DECKS, "udaInit_1L.ups"), "--init-uda", FIXTURE, "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "TASK RMCRT_Test::initializeWithUda level=0 patches=12" in r.stdout and "RMCRT_Test::initialize " not in r.stdout


@pytest.mark.gpu
def test_uda_init_solve_and_output(tools, po, tmp_path, gpu_lib):
    """RMCRT_udaInit.ups pattern end to end: the reference's 50^3 Arches state (12 patches) -> initializeWithUda -> sigmaT4 ->
    setBoundaryConditions -> Ray::rayTrace with boundary fluxes -> <DataArchiver> output, against the oracle on the same fields"""
    out_uda = str(tmp_path / "out.uda")
    r = subprocess.run([EXE, os.path.join(DECKS, "udaInit_1L.ups"), "--init-uda", FIXTURE, "--dump", str(tmp_path / "d"), "--uda", out_uda,
                        "--knob", "skipZeroWeightLegs=0"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = load_dump(str(tmp_path / "d"))
    fx = {n: U.read_level_var(FIXTURE, n)[0] for n in ("abskg", "temperature", "cellType")}
    inner = (slice(1, -1),) * 3
    assert np.array_equal(out["cellType"], fx["cellType"])                                     # copied incl. extra cells (850-853)
    assert np.array_equal(out["abskg"][inner], fx["abskg"][inner].astype(np.float64))          # float archive -> double variable (914-925)
    assert (out["abskg"][0] == 1.0).all() and (out["abskg"][:, :, -1] == 1.0).all()            # Dirichlet abskg on the extra cells
    T = fx["temperature"].astype(np.float64)
    sop = po.SIGMA / np.pi
    np.testing.assert_allclose(out["sigmaT4"][inner], sop * T[inner] ** 4, rtol=1e-14)
    np.testing.assert_allclose(out["sigmaT4"][0, 1:-1, 1:-1], sop * T[1, 1:-1, 1:-1] ** 4, rtol=1e-12)   # Neumann 0: wall T = adjacent cell
    N = 50
    lv = po.Level(out["abskg"], out["sigmaT4"], out["cellType"], (-1,) * 3, (N + 1,) * 3, (0.002,) * 3, anchor=(-0.05, -0.05, 0.0),
                  interior_lo=(0,) * 3, interior_hi=(N,) * 3)
    p = po.default_params(nDivQRays=3, nFluxRays=3, threshold=0.05, isSeedRandom=0, solveBoundaryFlux=1, allowReflect=1)
    ref = po.ray_trace(p, lv, patches=(1, 1, 8), threads=8)
    assert out["report"]["nRaySteps"] == ref["steps"] and out["report"]["nFluxRaySteps"] == ref["fluxSteps"]
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref["divQ"]).max())
    S, R = out["RMCRTboundFlux"], ref["boundFlux"]
    assert np.array_equal(S[..., 6], R[..., 6]) and np.array_equal(S[..., :6] != 0, R[..., :6] != 0)
    np.testing.assert_allclose(S, R, rtol=1e-9, atol=0)
    # the archive the driver wrote holds exactly what it computed (read back by the independent python reader) ...
    for name, key in (("divQ", "divQ"), ("radiationVolq", "radiationVolq"), ("RMCRTboundFlux", "RMCRTboundFlux"), ("sigmaT4", "sigmaT4"),
                      ("abskg", "abskg")):
        a, lo, hi = U.read_level_var(out_uda, name)
        assert lo == (-1,) * 3 and hi == (51,) * 3 and np.array_equal(a, out[key]), name
    # ... and compares clean against itself with the C++ tool, Stencil7 component by component
    r2, rows = lnorm(out_uda, out_uda)
    assert r2.returncode == 0 and ("RMCRTboundFlux.p", 0) in rows and all(v[:3] == (0.0, 0.0, 0.0) for v in rows.values())
