"""The C++ host mirror of the reference plugin surface (uintah_b200/host: Ray, RMCRTCommon, RMCRT_Test driver).

CPU part (-m "not gpu"): the decks parse, the same tasks with the same labels/dependencies/task graphs as the reference's
Ray::sched_* get registered, input errors raise ProblemSetupException, and without a GPU the solve fails loudly.
GPU part: `rmcrt_test <deck> --dump` end to end (deck -> grid -> DataWarehouse -> tasks -> C ABI -> kernels) against the oracle.
"""
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HOST = os.path.join(ROOT, "uintah_b200", "host")
EXE = os.path.join(HOST, "rmcrt_test")
DECKS = os.path.join(HERE, "decks")


@pytest.fixture(scope="module")
def exe():
    from uintah_b200 import build
    build.build()
    subprocess.run(["make", "-C", HOST, "-j4"], check=True, capture_output=True)
    return EXE


def run(exe, deck, *args, check=True):
    r = subprocess.run([exe, os.path.join(DECKS, deck), *args], capture_output=True, text=True, timeout=600)
    if check:
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


def tasks_of(stdout):
    tasks, cur = [], None
    for line in stdout.splitlines():
        if line.startswith("TASK "):
            name, rest = line[5:].split(" level=", 1)            # task names may hold blanks ("RadiationSource: setup bndflux")
            cur = {"name": name, **{k: v for k, v in (x.split("=") for x in ("level=" + rest).split())}, "deps": []}
            tasks.append(cur)
        elif line.startswith("  ") and cur is not None:
            f = line.split()
            cur["deps"].append((f[0], f[1], *f[2:]))
    return tasks


# ------------------------------------------------------------------------------------------------ CPU: surface
def test_single_level_schedule_matches_reference_surface(exe):
    """Ray::sched_rayTrace (Ray.cc:422-506): task name, labels, SHRT_MAX ghosts, TG_RMCRT; carry-forward on TG_CARRY_FORWARD"""
    t = tasks_of(run(exe, "bm1_1L_double.ups", "--parse-only").stdout)
    names = [x["name"] for x in t]
    assert names == ["RMCRT_Test::initialize", "RMCRT_Test::initProperties", "RMCRTCommon::carryForward_VarLabels",
                     "RMCRTCommon::sigmaT4", "Ray::setBoundaryConditions", "Ray::rayTrace"]
    rt = t[-1]
    assert rt["tg"] == "1" and rt["device"] == "1"
    req = [d for d in rt["deps"] if d[0] == "requires"]
    assert [d[1] for d in req] == ["abskg", "sigmaT4", "cellType"]
    assert all(d[3] == "ghost=32767" and d[2] == "dw=NewDW" for d in req)
    assert [d[1] for d in rt["deps"] if d[0] == "computes"] == ["divQ", "RMCRTboundFlux", "radiationVolq"]
    cf = t[2]
    assert cf["tg"] == "0" and {d[1] for d in cf["deps"]} == {"divQ", "RMCRTboundFlux", "radiationVolq", "sigmaT4"}


def test_float_mode_creates_abskgRMCRT(exe):
    """<RMCRT type="float"> with a double abskg: DoubleToFloat -> "abskgRMCRT" (RMCRTCommon.cc:188-191, 211-237)"""
    t = tasks_of(run(exe, "bm1_1L_float_8patches.ups", "--parse-only").stdout)
    d2f = [x for x in t if x["name"] == "RMCRTCommon::DoubleToFloat"]
    assert len(d2f) == 1 and ("computes", "abskgRMCRT", "dw=NewDW", "ghost=0", "domain=normal", "offset=0") in d2f[0]["deps"]
    rt = [x for x in t if x["name"] == "Ray::rayTrace"][0]
    assert rt["patches"] == "8" and rt["deps"][0][1] == "abskgRMCRT"


def test_data_onion_schedule(exe):
    """Ray::sched_rayTrace_dataOnion (Ray.cc:824-938): finest level only, halo ghosts on the fine level, whole coarser levels"""
    t = tasks_of(run(exe, "bm1_DO.ups", "--parse-only").stdout)
    rt = [x for x in t if x["name"] == "Ray::rayTrace_dataOnion"]
    assert len(rt) == 1 and rt[0]["level"] == "1" and rt[0]["patches"] == "8" and rt[0]["tg"] == "1"
    fine = [d for d in rt[0]["deps"] if d[0] == "requires" and d[4] == "domain=normal"]
    coarse = [d for d in rt[0]["deps"] if d[0] == "requires" and d[4] == "domain=coarse"]
    assert [d[3] for d in fine] == ["ghost=4"] * 3
    assert [d[1] for d in coarse] == ["abskg", "sigmaT4", "cellType"] and all(d[3] == "ghost=32767" and d[5] == "offset=1" for d in coarse)
    names = [x["name"] for x in t]
    assert names.index("Ray::coarsenAll") < names.index("Ray::rayTrace_dataOnion")


@pytest.mark.parametrize("edit,needle", [
    (("<periodic> [0,0,0] </periodic>", "<periodic> [1,0,0] </periodic>"), "Periodic boundary conditions are not allowed"),
    (("<extraCells> [1,1,1]   </extraCells>", "<extraCells> [0,0,0] </extraCells>"), "one extraCell"),
    (('<BCType id="0" label="abskg"    var="Dirichlet"> <value> 1.0 </value> </BCType>', ""), "boundary condition for (abskg)"),
    (("<Temperature>    64.804 </Temperature>", ""), "<Temperature>"),
])
def test_input_errors_raise_problem_setup_exception(exe, tmp_path, edit, needle):
    src = open(os.path.join(DECKS, "bm1_1L_double.ups")).read()
    assert edit[0] in src, edit[0]
    deck = tmp_path / "bad.ups"
    deck.write_text(src.replace(edit[0], edit[1], 1))
    r = subprocess.run([exe, str(deck), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "ProblemSetupException" in r.stderr and needle in r.stderr, r.stderr


def test_dynamic_roi_schedules_roi_extents(exe, tmp_path):
    """<ROI_extents type="dynamic">: Ray::sched_ROI_Extents (Ray.cc:1947-1978) adds the Ray::ROI_Extents task on the finest level, before the
    trace, computing mag_grad_abskg / mag_grad_sigmaT4 / flaggedCells and the ROI_loCell / ROI_hiCell reduction (the reference's own
    device build refuses this ROI, Ray.cc:338-345; patch_based decks get no such task)"""
    src = open(os.path.join(DECKS, "bm1_DO.ups")).read()
    deck = tmp_path / "dyn.ups"
    deck.write_text(src.replace('<ROI_extents type="patch_based"/>', '<ROI_extents type="dynamic"> <abskgd_threshold> 0.1 </abskgd_threshold> </ROI_extents>'))
    r = subprocess.run([exe, str(deck), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t = tasks_of(r.stdout)
    names = [x["name"] for x in t]
    assert names.count("Ray::ROI_Extents") == 1 and names.index("Ray::ROI_Extents") < names.index("Ray::rayTrace_dataOnion")
    roi = [x for x in t if x["name"] == "Ray::ROI_Extents"][0]
    assert roi["level"] == "1" and [d[1] for d in roi["deps"] if d[0] == "computes"] == ["mag_grad_abskg", "mag_grad_sigmaT4", "flaggedCells", "ROI_loCell", "ROI_hiCell"]
    assert [d[3] for d in roi["deps"] if d[0] == "requires"] == ["ghost=1", "ghost=1"]
    assert "Ray::ROI_Extents" not in [x["name"] for x in tasks_of(run(exe, "bm1_DO.ups", "--parse-only").stdout)]


def test_fixed_roi_and_coarsen_extra_cells_are_accepted(exe, tmp_path):
    """<ROI_extents type="fixed"> (Ray.cc:265-271) and <coarsenExtraCells> (Ray.cc:256): parsed and scheduled; with coarsenExtraCells the
    coarse level gets no setBoundaryConditions task (RMCRT_Test.cc:441-443); a fixed ROI off the fine level is refused like the reference
    does in computeExtents (Ray.cc:1308-1313)"""
    src = open(os.path.join(DECKS, "bm1_DO.ups")).read()
    deck = tmp_path / "fixed.ups"
    deck.write_text(src.replace('<ROI_extents type="patch_based"/>', '<ROI_extents type="fixed"> <min> [0.25,0.25,0.25] </min> <max> [0.75,0.75,0.75] </max> </ROI_extents>\n<coarsenExtraCells> true </coarsenExtraCells>'))
    r = subprocess.run([exe, str(deck), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    t = tasks_of(r.stdout)
    assert [x["level"] for x in t if x["name"] == "Ray::setBoundaryConditions"] == ["1"]      # fine level only
    base = tasks_of(run(exe, "bm1_DO.ups", "--parse-only").stdout)
    assert sorted(x["level"] for x in base if x["name"] == "Ray::setBoundaryConditions") == ["0", "1"]


def test_no_gpu_fails_loudly(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(exe, "bm1_1L_double.ups", check=False)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


# ------------------------------------------------------------------------------------------------ GPU: end to end
def load_dump(d):
    meta = json.load(open(os.path.join(d, "meta.json")))
    out = {"report": meta["report"]}
    for k, m in meta.items():
        if k == "report":
            continue
        a = np.fromfile(os.path.join(d, m["file"]), dtype=np.dtype("<" + m["dtype"]))
        shape = tuple(m["hi"][i] - m["lo"][i] for i in (2, 1, 0))
        out[k] = a.reshape(shape + ((m["components"],) if m["components"] > 1 else ()))
    return out


@pytest.mark.gpu
def test_bm1_single_level_double_deterministic(exe, po, tmp_path, gpu_lib):
    """RMCRT_bm1_1L pattern: <randomSeed> false, type double -> FP64 march, agreement with the oracle to round-off"""
    out = load_dump(run(exe, "bm1_1L_double.ups", "--dump", str(tmp_path), "--knob", "skipZeroWeightLegs=0").args[3])
    lv = po.burns_christon(21)
    # the driver's own initialisation + sigmaT4 + setBoundaryConditions reproduce the oracle's input fields bit for bit
    assert np.array_equal(out["abskg"], lv.abskg) and np.array_equal(out["cellType"], lv.cellType)
    np.testing.assert_allclose(out["sigmaT4"], lv.sigmaT4, rtol=1e-15, atol=0)
    ref = po.ray_trace(po.default_params(nDivQRays=3, threshold=0.05, isSeedRandom=0), lv)
    assert out["report"]["nRaySteps"] == ref["steps"] and out["report"]["nDivQCells"] == 21 ** 3
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref["divQ"]).max())
    np.testing.assert_allclose(out["radiationVolq"], ref["radiationVolq"], rtol=1e-9, atol=1e-12)
    assert not out["RMCRTboundFlux"].any()          # solveBoundaryFlux false: zero-initialised Stencil7 (Ray.cc:583-587)


@pytest.mark.gpu
def test_bm1_float_eight_patches_equals_whole_level_solve(exe, po, tmp_path, gpu_lib):
    """8 patches through the C++ task == one whole-level solve through the python harness (Philox keyed on the cell,
    rayCoherence 0, so the result is independent of the patch layout), and within Monte-Carlo error of the oracle"""
    from uintah_b200 import RMCRT
    out = load_dump(run(exe, "bm1_1L_float_8patches.ups", "--dump", str(tmp_path), "--knob", "rayCoherence=0").args[3])
    lv = po.burns_christon(32, dtype=np.float32)
    assert out["report"]["nDivQCells"] == 32 ** 3
    g = RMCRT(device=0, nDivQRays=32, threshold=0.05, rayCoherence=0)
    g.upload_level(0, lv)
    whole = g.trace_divq(0, timestep=2)["divQ"]      # the deck runs 2 time steps, both on TG_RMCRT (calc_frequency 2)
    np.testing.assert_allclose(out["divQ"], whole, rtol=1e-12, atol=1e-14)
    ref = po.ray_trace(po.default_params(nDivQRays=32, threshold=0.05), lv, seed=3, patches=(1, 1, 8), threads=8)["divQ"]
    assert np.linalg.norm(out["divQ"] - ref) / np.linalg.norm(ref) < 0.04


@pytest.mark.gpu
def test_bm1_data_onion_deterministic(exe, po, tmp_path, gpu_lib):
    """RMCRT_bm1_DO pattern (fine 32^3 in 8 patches, coarse 8^3, RR 4, halo 4): device coarsening + coarse BCs + DO trace"""
    out = load_dump(run(exe, "bm1_DO.ups", "--dump", str(tmp_path), "--knob", "skipZeroWeightLegs=0").args[3])
    fine = po.burns_christon(32)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    ref = po.ray_trace_data_onion(po.default_params(nDivQRays=3, threshold=0.05, isSeedRandom=0), [coarse, fine], (4, 4, 4),
                                  patches=(2, 2, 2), threads=8)
    assert out["report"]["nRaySteps"] == ref["steps"]
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref["divQ"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("roi,box", [
    ('<ROI_extents type="dynamic"> <abskgd_threshold> 0.5 </abskgd_threshold> </ROI_extents>', ((-1,) * 3, (33,) * 3)),          # every patch flagged
    ('<ROI_extents type="dynamic"> <abskgd_threshold> 1e9 </abskgd_threshold> </ROI_extents>', ((32767,) * 3, (-32767,) * 3)),   # none: empty ROI
    ('<ROI_extents type="fixed"> <min> [0.25,0.2,0.3] </min> <max> [0.8,0.75,0.7] </max> </ROI_extents>', ((8, 6, 9), (25, 24, 22))),
])
def test_data_onion_dynamic_and_fixed_roi_decks(exe, po, tmp_path, gpu_lib, roi, box):
    """bm1_DO.ups with the two other ROI types end to end through the plugin: Ray::ROI_Extents on the device (dynamic), getCellIndex of
    the two points (fixed), the same box for every patch in the trace; deterministic regression mode against the oracle"""
    deck = edit_deck(tmp_path, "bm1_DO.ups", ('<ROI_extents type="patch_based"/>', roi))
    out = load_dump(run(exe, deck, "--dump", str(tmp_path / "d"), "--knob", "skipZeroWeightLegs=0").args[3])
    fine = po.burns_christon(32)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    p = po.default_params(nDivQRays=3, threshold=0.05, isSeedRandom=0, roiFixed=1, roiFixedLo=box[0], roiFixedHi=box[1])
    ref = po.ray_trace_data_onion(p, [coarse, fine], (4, 4, 4), patches=(2, 2, 2), threads=8)
    assert out["report"]["nRaySteps"] == ref["steps"]
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref["divQ"]).max())


@pytest.mark.gpu
def test_bm2_boundary_flux_and_reflections(exe, po, tmp_path, gpu_lib):
    """RMCRT_1L_perf pattern: hot walls, x-/y- mirrors, solveBoundaryFlux, 2 patches: masks bit-exact, fluxes to round-off"""
    out = load_dump(run(exe, "bm2_1L_flux_reflect.ups", "--dump", str(tmp_path), "--knob", "skipZeroWeightLegs=0").args[3])
    lv = po.parallel_plates(16, mirror_faces=("x-", "y-"))
    assert np.array_equal(out["abskg"], lv.abskg) and np.array_equal(out["cellType"], lv.cellType)
    np.testing.assert_allclose(out["sigmaT4"], lv.sigmaT4, rtol=1e-14, atol=0)
    p = po.default_params(nDivQRays=4, nFluxRays=4, threshold=0.001, isSeedRandom=0, solveBoundaryFlux=1, allowReflect=1)
    ref = po.ray_trace(p, lv)
    S, R = out["RMCRTboundFlux"], ref["boundFlux"]
    assert np.array_equal(S[..., 6], R[..., 6])                        # boundFlux.p
    assert np.array_equal(S[..., :6] != 0, R[..., :6] != 0)            # face selection
    assert out["report"]["nRaySteps"] == ref["steps"] and out["report"]["nFluxRaySteps"] == ref["fluxSteps"]
    np.testing.assert_allclose(S, R, rtol=1e-9, atol=0)
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-9 * np.abs(ref["divQ"]).max())


# ------------------------------------------------------------------------------------------------ SURVEY 8f-4 features through the plugin
def edit_deck(tmp_path, deck, *edits):
    src = open(os.path.join(DECKS, deck)).read()
    for old, new in edits:
        assert old in src, old
        src = src.replace(old, new, 1)
    out = tmp_path / "edited.ups"
    out.write_text(src)
    return str(out)


BOUNDED = ("<CCRays>             false      </CCRays>",
           "<CCRays> false </CCRays>\n<algorithm type='singleLevel'> <ROI_extents type='boundedRayLength'> <length> 0.25 </length> </ROI_extents> </algorithm>")
LHC = ("<rayDirSampleAlgo>   naive      </rayDirSampleAlgo>", "<rayDirSampleAlgo> LatinHyperCube </rayDirSampleAlgo>")


def test_radiometer_schedule_matches_reference_surface(exe):
    """Radiometer::sched_radiometer (Radiometer.cc:346-398): initialize_VRFlux on every patch, the spatial radiometerTask only
    on the patches that hold radiometer cells, whole-level requires, modifies VRFlux / VRIntensity; RMCRT_Test's
    radiometerOnly graph (RMCRT_Test.cc:559-583) computes no divQ"""
    r = run(exe, "radiometer_only.ups", "--parse-only")
    t = tasks_of(r.stdout)
    names = [x["name"] for x in t]
    assert names == ["RMCRT_Test::initialize", "Radiometer::initialize_VRFlux", "RMCRT_Test::initProperties",
                     "RMCRTCommon::carryForward_VarLabels", "RMCRTCommon::sigmaT4", "Ray::setBoundaryConditions",
                     "Radiometer::initialize_VRFlux", "Radiometer::radiometerTask"]
    rt = t[-1]
    assert rt["patches"] == "3" and rt["tg"] == "1" and rt["device"] == "1"      # 4 patches, one holds no radiometer cell
    assert [(d[1], d[3]) for d in rt["deps"] if d[0] == "requires"] == [("abskg", "ghost=32767"), ("sigmaT4", "ghost=32767"), ("cellType", "ghost=32767")]
    assert [d[1] for d in rt["deps"] if d[0] == "modifies"] == ["VRFlux", "VRIntensity"]
    assert {d[1] for d in t[3]["deps"]} == {"sigmaT4", "VRFlux", "VRIntensity"} and t[3]["tg"] == "0"
    assert "lower cell [7,8,14] upper cell [9,9,16]" in r.stdout


@pytest.mark.parametrize("edit,needle", [
    (("<locationsMax>      [0.13,0.07,0.07]           </locationsMax>", "<locationsMax> [1.3,0.07,0.07] </locationsMax>"), "outside of the computational domain"),
    (("<locationsMax>      [0.13,0.07,0.07]           </locationsMax>", "<locationsMax> [0.13,0.0,0.07] </locationsMax>"), "same starting and ending points"),
    (("<viewAngle>         30                         </viewAngle>", "<viewAngle> 400 </viewAngle>"), "exceeds the maximum acceptable value of 360"),
    (("<unitNormalVector>  [0.57735,0.57735,0.57735]  </unitNormalVector>", ""), "<unitNormalVector> or <normalVector>"),
    (("<nRays>             40                         </nRays>", ""), "<nRays>"),
])
def test_radiometer_input_errors(exe, tmp_path, edit, needle):
    r = subprocess.run([exe, edit_deck(tmp_path, "radiometer_only.ups", edit), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "ProblemSetupException" in r.stderr and needle in r.stderr, r.stderr


def test_bounded_ray_length_and_lhc_are_accepted(exe, tmp_path):
    """Ray.cc:455-475: ghost cells = ceil(|length/Dx|) for boundedRayLength; Ray.cc:175-180 LatinHyperCube"""
    r = subprocess.run([exe, edit_deck(tmp_path, "bm1_1L_double.ups", BOUNDED, LHC), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Using Latin Hyper Cube method" in r.stdout and "bounded ray length has been set to: (0.25)" in r.stdout
    rt = [x for x in tasks_of(r.stdout) if x["name"] == "Ray::rayTrace"][0]
    n = int(np.ceil(np.sqrt(3) * 0.25 * 21))
    assert [d[3] for d in rt["deps"] if d[0] == "requires"] == [f"ghost={n}"] * 3


@pytest.mark.gpu
def test_radiometer_only_deck(exe, po, tmp_path, gpu_lib):
    """radiometerOnly through the plugin: VRFlux / VRIntensity of both radiometers against Radiometer::radiometerFlux's restatement"""
    out = load_dump(run(exe, "radiometer_only.ups", "--dump", str(tmp_path), "--knob", "skipZeroWeightLegs=0").args[3])
    assert "divQ" not in out
    lv = po.burns_christon(16)
    p = po.default_params(threshold=0.05, isSeedRandom=0, allowReflect=0)
    rads = [po.radiometer(30.0, (0.57735, 0.57735, 0.57735), (0, 0, 0), (2, 1, 1), 40),
            po.radiometer(90.0, (0.5, 0.5, -0.5), (7, 8, 14), (9, 9, 16), 25, is_normal_vector=True)]
    fo, io, _ = po.radiometer_flux(p, lv, rads)
    assert np.count_nonzero(fo) == 2 + 4
    assert np.array_equal(out["VRFlux"] != 0, fo != 0)
    np.testing.assert_allclose(out["VRFlux"], fo, rtol=1e-10)
    np.testing.assert_allclose(out["VRIntensity"], io, rtol=1e-10)


@pytest.mark.gpu
def test_single_level_with_radiometer_bounded_lhc(exe, po, tmp_path, gpu_lib):
    """singleLevel + <Radiometer> + boundedRayLength + LatinHyperCube in one deck (RMCRT_bm1_1L_bounded / RMCRT_VR patterns)"""
    rad = ("</RMCRT>", "<Radiometer> <viewAngle> 60 </viewAngle> <normalVector> [0,0,2] </normalVector> <locationsMin> [0.4,0.4,0.0] </locationsMin>"
                       " <locationsMax> [0.5,0.5,0.06] </locationsMax> <nRays> 30 </nRays> </Radiometer>\n</RMCRT>")
    deck = edit_deck(tmp_path, "bm1_1L_double.ups", BOUNDED, LHC, rad)
    r = subprocess.run([exe, deck, "--dump", str(tmp_path / "d"), "--knob", "skipZeroWeightLegs=0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = load_dump(str(tmp_path / "d"))
    lv = po.burns_christon(21)
    kw = dict(nDivQRays=3, threshold=0.05, isSeedRandom=0, rayDirSampleAlgo=1, maxRayLength=0.25)
    ref = po.ray_trace(po.default_params(**kw), lv)
    assert out["report"]["nRaySteps"] == ref["steps"]
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref["divQ"]).max())
    lo = (int(np.floor(0.4 * 21)), int(np.floor(0.4 * 21)), 0)
    hi = (int(np.floor(0.5 * 21)), int(np.floor(0.5 * 21)), int(np.floor(0.06 * 21)))
    rads = [po.radiometer(60.0, (0, 0, 2), lo, hi, 30, is_normal_vector=True)]
    fo, io, _ = po.radiometer_flux(po.default_params(**kw), lv, rads)
    np.testing.assert_allclose(out["VRFlux"], fo, rtol=1e-10)


# ------------------------------------------------------------------------------------------------ RMCRT_Test benchmarks 6-8 and <Intrusion>
BM7 = ("<benchmark>      1      </benchmark>", '<benchmark dir="1"> 7 </benchmark>')
CYL = ("</Uintah_specification>", "<Intrusion> <geom_object> <cylinder> <bottom>[0,.5,.5]</bottom> <top>[1,.5,.5]</top> <radius>.1</radius>"
                                  " </cylinder> </geom_object> </Intrusion>\n</Uintah_specification>")


def test_benchmark_direction_is_required(exe, tmp_path):
    """RMCRT_Test.cc:146-151"""
    deck = edit_deck(tmp_path, "bm1_1L_double.ups", ("<benchmark>      1      </benchmark>", "<benchmark> 6 </benchmark>"))
    r = subprocess.run([exe, deck, "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "requires the attribute 'dir'" in r.stderr
    deck = edit_deck(tmp_path, "bm1_1L_double.ups", ("<benchmark>      1      </benchmark>", "<benchmark> 4 </benchmark>"))
    r = subprocess.run([exe, deck, "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "isotropic scattering" in r.stderr


@pytest.mark.gpu
def test_benchmark7_with_cylinder_intrusion(exe, po, tmp_path, gpu_lib):
    """RMCRT_Test benchmark 7 along y (RMCRT_Test.cc:736-756) with the cylinder intrusion of RMCRT_test_1L_wIntrusion.ups:115-123:
    the driver's fields equal the oracle's restatement and the deterministic solve agrees to round-off (rays stop at the intrusion)"""
    deck = edit_deck(tmp_path, "bm1_1L_double.ups", BM7, CYL)
    r = subprocess.run([exe, deck, "--dump", str(tmp_path / "d"), "--knob", "skipZeroWeightLegs=0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = load_dump(str(tmp_path / "d"))
    N = 21
    lv = po.benchmark_level(7, N, pDir=1)
    c = (np.arange(-1, N + 1) + 0.5) / N
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    inside = (x >= 0) & (x <= 1) & ((y - 0.5) ** 2 + (z - 0.5) ** 2 <= 0.01)
    lv.cellType[inside] = 8
    assert inside[1:-1, 1:-1, 1:-1].sum() > 100
    assert np.array_equal(out["cellType"], lv.cellType)
    np.testing.assert_allclose(out["abskg"], lv.abskg, rtol=1e-15)
    np.testing.assert_allclose(out["sigmaT4"], lv.sigmaT4, rtol=1e-13)
    lv.abskg[...] = out["abskg"]; lv.sigmaT4[...] = out["sigmaT4"]
    ref = po.ray_trace(po.default_params(nDivQRays=3, threshold=0.05, isSeedRandom=0), lv)
    assert out["report"]["nRaySteps"] == ref["steps"] and out["report"]["nDivQCells"] == int((lv.cellType == -1).sum())
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref["divQ"]).max())
    assert not out["divQ"][lv.cellType != -1].any()


# ------------------------------------------------------------------------------------------------ <algorithm type="RMCRT_coarseLevel">
def coarse_level_deck(tmp_path, order=1):
    return edit_deck(tmp_path, "bm1_DO.ups", ('''<algorithm type="dataOnion">
      <haloCells>        [4,4,4]    </haloCells>
      <ROI_extents type="patch_based"/>
    </algorithm>''', f'<algorithm type="RMCRT_coarseLevel"> <orderOfInterpolation> {order} </orderOfInterpolation> </algorithm>'))


def test_coarse_level_schedule_matches_reference_surface(exe, tmp_path):
    """RMCRT_Test.cc:464-519: CoarsenAll + BCs on the coarse level, Ray::rayTrace on the COARSE level, Ray::refine_Q on the fine one
    with the dependencies of Ray::sched_Refine_Q (Ray.cc:1841-1867)"""
    t = tasks_of(run(exe, coarse_level_deck(tmp_path), "--parse-only").stdout)
    names = [(x["name"], x["level"]) for x in t]
    assert ("Ray::rayTrace", "0") in names and ("Ray::rayTrace", "1") not in names
    assert names.index(("Ray::coarsenAll", "0")) < names.index(("Ray::setBoundaryConditions", "0")) < names.index(("Ray::rayTrace", "0")) \
        < names.index(("Ray::refine_Q", "1"))
    rq = [x for x in t if x["name"] == "Ray::refine_Q"][0]
    assert rq["patches"] == "8" and rq["tg"] == "1"
    assert [(d[1], d[3], d[4], d[5]) for d in rq["deps"] if d[0] == "requires"] == \
        [(n, "ghost=1", "domain=coarse", "offset=1") for n in ("divQ", "RMCRTboundFlux", "radiationVolq")]
    assert [d[1] for d in rq["deps"] if d[0] == "computes"] == ["divQ", "RMCRTboundFlux", "radiationVolq"]
    r = subprocess.run([exe, edit_deck(tmp_path, "bm1_DO.ups", ('<haloCells>        [4,4,4]    </haloCells>', ''),
                                       ('type="dataOnion"', 'type="RMCRT_coarseLevel"')), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "orderOfInterpolation" in r.stderr      # alg_ps->require (Ray.cc:314)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 2])
def test_coarse_level_deck(exe, po, tmp_path, gpu_lib, order):
    """RMCRT_ML.ups pattern through the plugin: coarse 8^3 solve == the oracle's Ray::rayTrace on the coarsened level, fine 32^3
    divQ == the oracle's refine_Q of it, bit for bit"""
    r = subprocess.run([exe, coarse_level_deck(tmp_path, order), "--dump", str(tmp_path / "d"), "--knob", "skipZeroWeightLegs=0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = load_dump(str(tmp_path / "d"))
    fine = po.burns_christon(32)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    ref_c = po.ray_trace(po.default_params(nDivQRays=3, threshold=0.05, isSeedRandom=0), coarse)
    assert out["report"]["nRaySteps"] == ref_c["steps"] and out["report"]["nDivQCells"] == 8 ** 3
    np.testing.assert_allclose(out["divQ_L0"], ref_c["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref_c["divQ"]).max())
    ref_f = po.refine_Q(out["divQ_L0"], coarse.lo, coarse.hi, coarse.lo, coarse.hi, (4, 4, 4), order, (0,) * 3, (32,) * 3, fine.lo, fine.hi)
    assert np.array_equal(out["divQ"], ref_f)
    assert not out["RMCRTboundFlux"].any()


# ------------------------------------------------------------------------------------------------ the Wasatch consumer (modifies_divQ = true)
def test_wasatch_schedule_matches_reference_surface(exe):
    """RadiationSource::schedule_ray_tracing (Wasatch/Expressions/RadiationSource.cc:141-172): setup bndflux, sigmaT4, BCs from the
    temperature field (backoutTemp false), Ray::rayTrace scheduled with MODIFIES on divQ / RMCRTboundFlux / radiationVolq"""
    t = tasks_of(run(exe, "bm2_1L_flux_reflect.ups", "--caller", "wasatch", "--parse-only").stdout)
    names = [x["name"] for x in t]
    assert names.index("RMCRTCommon::carryForward_VarLabels") < names.index("RMCRTCommon::sigmaT4") < names.index("Ray::setBoundaryConditions") \
        < names.index("Ray::rayTrace")
    rt = t[-1]
    assert rt["name"] == "Ray::rayTrace" and [d[1] for d in rt["deps"] if d[0] == "modifies"] == ["divQ", "RMCRTboundFlux", "radiationVolq"]
    assert not [d for d in rt["deps"] if d[0] == "computes"]
    bc = [x for x in t if x["name"] == "Ray::setBoundaryConditions"][0]
    assert ("requires", "color", "dw=NewDW", "ghost=0", "domain=normal", "offset=0") in bc["deps"]      # backoutTemp == false (Ray.cc:1604-1606)


@pytest.mark.gpu
def test_wasatch_caller_modifies_divQ(exe, po, tmp_path, gpu_lib):
    """modifies_divQ = true end to end: cells the trace does not compute keep what the caller put there (the expression tree's field),
    everything else equals the oracle; the boundary temperature comes from the temperature field, not from sigmaT4"""
    out = load_dump(run(exe, "bm2_1L_flux_reflect.ups", "--caller", "wasatch", "--dump", str(tmp_path), "--knob", "skipZeroWeightLegs=0").args[5])
    lv = po.parallel_plates(16, mirror_faces=("x-", "y-"))
    np.testing.assert_allclose(out["sigmaT4"], lv.sigmaT4, rtol=1e-14, atol=0)
    p = po.default_params(nDivQRays=4, nFluxRays=4, threshold=0.001, isSeedRandom=0, solveBoundaryFlux=1, allowReflect=1)
    ref = po.ray_trace(p, lv)
    flow = lv.cellType == -1
    np.testing.assert_allclose(out["divQ"][flow], ref["divQ"][flow], rtol=1e-9, atol=1e-9 * np.abs(ref["divQ"]).max())
    assert (out["divQ"][~flow] == 123.0).all() and (out["radiationVolq"][~flow] == -7.0).all()      # untouched (Ray.cc:573-580)
    np.testing.assert_allclose(out["RMCRTboundFlux"], ref["boundFlux"], rtol=1e-9, atol=0)
    assert out["report"]["nRaySteps"] == ref["steps"] and out["report"]["nFluxRaySteps"] == ref["fluxSteps"]


@pytest.mark.gpu
def test_isotropic_scattering_deck(exe, po, tmp_path, gpu_lib):
    """RMCRT_isoScat.ups pattern through the plugin: <sigmaScat> 1, boundary fluxes, <randomSeed> false.  The stock reference needs a
    --enable-ray-scatter build for this deck; here the scattering march is selected at run time"""
    deck = edit_deck(tmp_path, "bm1_1L_double.ups",
                     ("<nDivQRays>          3          </nDivQRays>", "<nDivQRays> 3 </nDivQRays> <sigmaScat> 1 </sigmaScat> <nFluxRays> 5 </nFluxRays>"),
                     ("<solveBoundaryFlux>  false      </solveBoundaryFlux>", "<solveBoundaryFlux> true </solveBoundaryFlux>"))
    r = subprocess.run([exe, deck, "--dump", str(tmp_path / "d"), "--knob", "skipZeroWeightLegs=0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Ray scattering is enabled" in r.stdout
    out = load_dump(str(tmp_path / "d"))
    lv = po.burns_christon(21)
    ref = po.ray_trace(po.default_params(nDivQRays=3, nFluxRays=5, threshold=0.05, isSeedRandom=0, solveBoundaryFlux=1, sigmaScat=1.0), lv)
    assert abs(out["report"]["nRaySteps"] - ref["steps"]) <= 2e-4 * ref["steps"]
    bad = np.abs(out["divQ"] - ref["divQ"]) > 1e-9 * np.abs(ref["divQ"]).max()
    assert bad.mean() < 2e-3
    assert np.array_equal(out["RMCRTboundFlux"][..., 6], ref["boundFlux"][..., 6])


# ------------------------------------------------------------------------------------------------ non-rectangular levels (RMCRT_+Domain*.ups)
def test_plus_domain_grid(exe):
    """a level made of seven <Box>es: 13 patches, extra cells only on patch faces that no other patch covers"""
    t = tasks_of(run(exe, "plus_domain_1L.ups", "--parse-only").stdout)
    rt = [x for x in t if x["name"] == "Ray::rayTrace"][0]
    assert rt["patches"] == "13" and rt["level"] == "0"


@pytest.mark.gpu
def test_plus_domain_deck(exe, po, tmp_path, gpu_lib):
    """RMCRT_+Domain.ups pattern: the whole-level gather covers the bounding box of the plus, the cells no patch owns are not flow cells
    (a ray meets the legs' extra cells first) -- fields and the deterministic solve against the oracle on the same arrays"""
    out = load_dump(run(exe, "plus_domain_1L.ups", "--dump", str(tmp_path), "--knob", "skipZeroWeightLegs=0").args[3])
    N = 40
    ct = out["cellType"]
    assert ct.shape == (N + 2,) * 3
    idx = np.arange(-1, N + 1)
    k, j, i = np.meshgrid(idx, idx, idx, indexing="ij")
    mid = lambda a: (a >= 16) & (a < 24)
    inr = lambda a: (a >= 0) & (a < N)
    flow = (mid(j) & mid(k) & inr(i)) | (mid(i) & mid(k) & inr(j)) | (mid(i) & mid(j) & inr(k))
    assert np.array_equal(ct == -1, flow) and flow.sum() == 8 * 8 * (3 * 40 - 2 * 8)
    assert set(np.unique(ct)) == {-1, 0, 8}                      # flow, not owned by any patch, wall
    # walls: exactly the cells one step outside the flow region along an axis (no extra cells on the internal box interfaces)
    sh = lambda a, ax, d: np.roll(a, d, axis=ax)
    touching = np.zeros_like(flow)
    for ax in range(3):
        for d in (1, -1):
            touching |= sh(flow, ax, d)
    assert (ct[touching & ~flow] == 8).all()
    ab = 0.9 * (1 - 2 * np.abs((i - (N - 1.0) / 2) / N)) * (1 - 2 * np.abs((j - (N - 1.0) / 2) / N)) * (1 - 2 * np.abs((k - (N - 1.0) / 2) / N)) + 0.1
    np.testing.assert_allclose(out["abskg"][flow], ab[flow], rtol=1e-13)
    lv = po.Level(out["abskg"], out["sigmaT4"], ct, (-1,) * 3, (N + 1,) * 3, (1.0 / N,) * 3, interior_lo=(0,) * 3, interior_hi=(N,) * 3)
    ref = po.ray_trace(po.default_params(nDivQRays=3, threshold=0.05, isSeedRandom=0), lv, patches=(1, 1, 8), threads=8)
    assert out["report"]["nRaySteps"] == ref["steps"] and out["report"]["nDivQCells"] == int(flow.sum())
    np.testing.assert_allclose(out["divQ"], ref["divQ"], rtol=1e-9, atol=1e-10 * np.abs(ref["divQ"]).max())
