"""The Arches caller of the plugin, RMCRT_Radiation (uintah_b200/host/ArchesRMCRT.{h,cc} + the arches_rmcrt_test driver): SURVEY 8f-2.

CPU part: the <src type="rmcrt_radiation"> block parses, the reference's tasks / labels / dependencies get registered, input errors
raise ProblemSetupException.  GPU part: the whole chain sigmaT4 (gas + particles) -> sumAbsk -> Ray::rayTrace with boundary fluxes ->
stencilToDBLs through the C++ tasks and the C ABI against the oracle's restatements, deterministic mode, round-off agreement.
"""
import os
import subprocess

import numpy as np
import pytest

from test_host_cpp import DECKS, HOST, edit_deck, load_dump, tasks_of

EXE = os.path.join(HOST, "arches_rmcrt_test")
DECK = "arches_rmcrt_1L.ups"


@pytest.fixture(scope="module")
def exe():
    from uintah_b200 import build
    build.build()
    subprocess.run(["make", "-C", HOST, "-j4"], check=True, capture_output=True)
    return EXE


def test_arches_schedule_matches_reference_surface(exe):
    """RMCRT_Radiation::sched_computeSource, singleLevel branch (Arches/SourceTerms/RMCRT.cc:513-533) + sched_initialize (574-660)"""
    r = subprocess.run([exe, os.path.join(DECKS, DECK), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "sumAbsk = (gas_abskg + abskp_0 + abskp_1)" in r.stdout
    t = tasks_of(r.stdout)
    names = [x["name"] for x in t]
    tail = ["RMCRTCommon::carryForward_VarLabels", "RMCRT_Radiation::sigmaT4", "RMCRT_Radiation::sumAbsk", "Ray::rayTrace",
            "RMCRT_Radiation::stencilToDBLs"]
    assert names[-5:] == tail
    assert "RMCRT_Radiation::sched_initialize_L-0" in names and "RMCRT_Radiation::fluxInit" in names
    byname = {x["name"]: x for x in t}
    s4 = byname["RMCRT_Radiation::sigmaT4"]
    assert s4["tg"] == "1" and [d[1] for d in s4["deps"] if d[0] == "requires"] == \
        ["volFraction", "gas_abskg", "radiation_temperature", "abskp_0", "temperature_0", "abskp_1", "temperature_1"]
    assert all(d[2] == "dw=OldDW" for d in s4["deps"] if d[0] == "requires")
    sa = byname["RMCRT_Radiation::sumAbsk"]
    assert [d[1] for d in sa["deps"]] == ["volFraction", "gas_abskg", "abskp_0", "abskp_1", "RMCRT_sumAbsk"]
    rt = byname["Ray::rayTrace"]       # abskg label == the component's RMCRT_sumAbsk (registerVariables, RMCRTCommon.cc:157-203)
    assert [d[1] for d in rt["deps"] if d[0] == "requires"] == ["RMCRT_sumAbsk", "sigmaT4", "cellType"]
    assert [d[1] for d in rt["deps"] if d[0] == "computes"] == ["divQ", "RMCRTboundFlux", "radiationVolq"]
    st = byname["RMCRT_Radiation::stencilToDBLs"]
    assert [d[1] for d in st["deps"] if d[0] == "computes"] == ["radiationFlux" + f for f in "EWNSTB"]
    cf = byname["RMCRTCommon::carryForward_VarLabels"]
    assert cf["tg"] == "0" and {d[1] for d in cf["deps"]} == {"divQ", "RMCRTboundFlux", "radiationVolq", "RMCRT_sumAbsk", "sigmaT4"}


@pytest.mark.parametrize("edit,needle", [
    (("<gas_absk> gas_abskg </gas_absk>", ""), "<gas_absk>"),
    (("<calc_frequency> 2 </calc_frequency>", "<calc_frequency> 2 </calc_frequency> <calc_on_all_RKsteps> true </calc_on_all_RKsteps>"), "calc_on_all_RKstes = false"),
    (("<DQMOM> <number_quad_nodes> 2 </number_quad_nodes> </DQMOM>", ""), "only works for DQMOM/CQMOM"),
    (("<calc_frequency> 2 </calc_frequency>", ""), "<calc_frequency>"),
    (('<algorithm type="singleLevel"/>', '<algorithm type="dataOnion"> <haloCells> [4,4,4] </haloCells> <ROI_extents type="patch_based"/> </algorithm>'),
     "boundary condition for (radiation_temperature)"),      # multi-level algorithms switch Ray's own BCs on (RMCRT.cc:155,161,167)
])
def test_arches_input_errors(exe, tmp_path, edit, needle):
    r = subprocess.run([exe, edit_deck(tmp_path, DECK, edit), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "ProblemSetupException" in r.stderr and needle in r.stderr, r.stderr


def arches_fields(N):
    """the synthetic Arches state of uintah_b200/host/Arches_Test.cc (ArchesFields), rebuilt bit for bit"""
    idx = np.arange(-1, N + 1)
    k, j, i = np.meshgrid(idx, idx, idx, indexing="ij")
    wall = (i < 0) | (j < 0) | (k < 0) | (i >= N) | (j >= N) | (k >= N)
    intr = (i >= N // 4) & (i < N // 2) & (j >= N // 4) & (j < N // 2) & (k >= N // 2) & (k < 3 * N // 4)
    solid = wall | intr
    dark = np.mod(i + j + k, 11) == 0
    Tg = 900.0 + 7.5 * np.mod(3 * i + 5 * j + 7 * k, 97)
    T0 = np.where(wall, 600.0, np.where(solid, 700.0, Tg))
    ag = np.where(solid | dark, 0.0, 0.1 + 0.004 * np.mod(i + 2 * j + 3 * k, 50))
    temps, absks = [T0], [ag]
    for q in range(2):
        temps.append(np.where(solid, T0, Tg + 50.0 * (q + 1)))
        absks.append(np.where(solid | dark, 0.0, 0.02 * (q + 1) * (1 + np.mod(i + j + k + q, 3))))
    return dict(vol=np.where(solid, 0.0, 1.0), cellType=np.where(solid, 8, -1).astype(np.int32), temps=temps, absks=absks, dark=dark & ~solid)


@pytest.mark.gpu
@pytest.mark.parametrize("flt", [False, True])
def test_arches_single_level_chain(exe, po, tmp_path, gpu_lib, flt):
    edits = [('<RMCRT type="double">', '<RMCRT type="float">')] if flt else []
    deck = edit_deck(tmp_path, DECK, *edits) if edits else os.path.join(DECKS, DECK)
    r = subprocess.run([exe, deck, "--dump", str(tmp_path / "d"), "--knob", "skipZeroWeightLegs=0", "--knob", "marchPrecision=1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "timestep 1: task graph TG_RMCRT" in r.stdout and "timestep 3: task graph TG_CARRY_FORWARD" in r.stdout
    out = load_dump(str(tmp_path / "d"))
    N = 16
    f = arches_fields(N)
    assert np.array_equal(out["cellType"], f["cellType"]) and np.array_equal(out["volFraction"], f["vol"])
    for name, a in (("radiation_temperature", f["temps"][0]), ("gas_abskg", f["absks"][0]), ("temperature_1", f["temps"][2]), ("abskp_0", f["absks"][1])):
        assert np.array_equal(out[name], a), name
    dt = np.float32 if flt else np.float64
    s4 = po.arches_sigmaT4(f["vol"], f["absks"], f["temps"], True, dt)
    ab = po.arches_sumAbsk(f["vol"], f["absks"], (-1,) * 3, (N + 1,) * 3, (0,) * 3, (N,) * 3, dt)
    assert np.array_equal(out["sigmaT4"], s4) and np.array_equal(out["RMCRT_sumAbsk"], ab)
    assert (s4[f["dark"]] == 0).all() and (ab[f["dark"]] == 0).all() and f["dark"].sum() > 100     # gas cells without any absorber
    lv = po.Level(ab, s4, f["cellType"], (-1,) * 3, (N + 1,) * 3, (1.0 / N,) * 3, interior_lo=(0,) * 3, interior_hi=(N,) * 3)
    p = po.default_params(nDivQRays=3, nFluxRays=3, threshold=0.05, isSeedRandom=0, solveBoundaryFlux=1, allowReflect=1)
    ref = po.ray_trace(p, lv)
    assert out["report"]["nRaySteps"] == ref["steps"] and out["report"]["nFluxRaySteps"] == ref["fluxSteps"]
    assert out["report"]["nDivQCells"] == int((f["cellType"] == -1).sum())
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref["divQ"]).max())
    S, R = out["RMCRTboundFlux"], ref["boundFlux"]
    assert np.array_equal(S[..., 6], R[..., 6]) and np.array_equal(S[..., :6] != 0, R[..., :6] != 0)
    np.testing.assert_allclose(S, R, rtol=1e-9, atol=0)
    six = po.arches_stencilToDBLs(S)
    for k in "EWNSTB":
        assert np.array_equal(out["radiationFlux" + k], six[k]), k
    assert np.count_nonzero(out["radiationFluxW"]) > 0 and np.count_nonzero(out["radiationFluxT"]) > 0
