"""GPU parity tests of the RMCRT features beside the plain divQ solve: Latin-Hyper-Cube sampling,
boundedRayLength ROI, the virtual radiometer.  Same bars as test_gpu_parity.py: integer outputs and step counts bit-exact,
deterministic mode (<randomSeed> false, FP64 march) to round-off, production mode within the Monte-Carlo tolerance."""
import numpy as np
import pytest

from util import random_rays

pytestmark = pytest.mark.gpu


def make(gpu_lib, **kw):
    from uintah_b200 import RMCRT
    return RMCRT(device=0, **kw)


# ------------------------------------------------------------------------------------------------ Latin Hyper Cube
@pytest.mark.parametrize("flux", [False, True])
def test_lhc_deterministic_matches_oracle(gpu_lib, po, flux):
    """<rayDirSampleAlgo> LatinHyperCube with <randomSeed> false: randVector's MT19937 Fisher-Yates permutation per
    (i+j+k) (RMCRTCommon.cc:998-1024) + per-ray reseed -- reproduced to round-off, identical step counts"""
    lv = po.parallel_plates(12) if flux else po.burns_christon(14)
    kw = dict(nDivQRays=12, nFluxRays=9, threshold=0.01, isSeedRandom=0, rayDirSampleAlgo=1, solveBoundaryFlux=1 if flux else 0)
    ref = po.ray_trace(po.default_params(**kw), lv)
    g = make(gpu_lib, marchPrecision=1, skipZeroWeightLegs=0, **kw)
    g.upload_level(0, lv)
    out = g.trace_divq(0)
    c = g.counters()
    assert c.nRaySteps == ref["steps"] and c.nFluxRaySteps == ref["fluxSteps"]
    scale = np.abs(ref["divQ"]).max()
    assert np.abs(out["divQ"] - ref["divQ"]).max() <= 1e-10 * scale
    if flux:
        assert np.array_equal(out["boundFlux"][..., 6], ref["boundFlux"][..., 6])
        assert np.array_equal(out["boundFlux"] != 0, ref["boundFlux"] != 0)
        np.testing.assert_allclose(out["boundFlux"], ref["boundFlux"], rtol=1e-10, atol=1e-10 * np.abs(ref["boundFlux"]).max())
    # differs from the naive sampler on the same deck
    kw0 = dict(kw, rayDirSampleAlgo=0)
    g0 = make(gpu_lib, marchPrecision=1, skipZeroWeightLegs=0, **kw0)
    g0.upload_level(0, lv)
    assert np.abs(g0.trace_divq(0)["divQ"] - out["divQ"]).max() > 1e-6 * scale


def test_lhc_fixed_random_matches_oracle(gpu_lib, po):
    """FIXED_RANDOM_NUM build: rand_int = 0.3*i in the shuffle (RMCRTCommon.cc:1017-1019) and every draw 0.3"""
    lv = po.burns_christon(10)
    kw = dict(nDivQRays=7, threshold=0.05, fixedRandom=1, rayDirSampleAlgo=1)
    ref = po.ray_trace(po.default_params(**kw), lv)
    g = make(gpu_lib, marchPrecision=1, skipZeroWeightLegs=0, **kw)
    g.upload_level(0, lv)
    out = g.trace_divq(0)
    assert g.counters().nRaySteps == ref["steps"]
    assert np.abs(out["divQ"] - ref["divQ"]).max() <= 1e-10 * np.abs(ref["divQ"]).max()


@pytest.mark.parametrize("coherence", [0, 1])
def test_lhc_production_is_unbiased_and_stratified(gpu_lib, po, coherence):
    """production LHC (Philox + keyed permutation): same mean as the oracle's LHC runs, and a smaller seed-to-seed
    spread than the production naive sampler at the same ray count"""
    lv = po.burns_christon(17)
    K = 6
    ref = np.mean([po.ray_trace(po.default_params(nDivQRays=64, threshold=0.05, rayDirSampleAlgo=1), lv, seed=50 + k)["divQ"]
                   for k in range(K)], 0)
    runs = {}
    for algo in (0, 1):
        g = make(gpu_lib, nDivQRays=64, threshold=0.05, rayDirSampleAlgo=algo, rayCoherence=coherence)
        g.upload_level(0, lv)
        runs[algo] = [g.trace_divq(0, timestep=10 + k)["divQ"] for k in range(K)]
    inner = (slice(1, -1),) * 3
    mu = np.mean(runs[1], 0)
    assert np.linalg.norm((mu - ref)[inner]) / np.linalg.norm(ref[inner]) < 0.02
    assert abs(mu[inner].mean() - ref[inner].mean()) < 0.005 * abs(ref[inner].mean())
    s_naive = np.std(runs[0], 0)[inner].mean()
    s_lhc = np.std(runs[1], 0)[inner].mean()
    assert s_lhc < 0.9 * s_naive


# ------------------------------------------------------------------------------------------------ bounded ray length
@pytest.mark.parametrize("max_len", [0.08, 0.3])
def test_bounded_ray_length_debug_rays(gpu_lib, po, max_len):
    lv = po.parallel_plates(16)
    rng = np.random.default_rng(21)
    cells, dirs, org = random_rays(lv, 300, rng)
    p = po.default_params(threshold=0.001, allowReflect=1, maxRayLength=max_len)
    ref = [po.update_sumI(p, lv, dirs[t], org[t], cells[t], max_path=3000) for t in range(len(cells))]
    ref_s = np.array([r[0] for r in ref]); ref_n = np.array([r[1] for r in ref])
    g = make(gpu_lib, threshold=0.001, allowReflect=1, marchPrecision=1, skipZeroWeightLegs=0, maxRayLength=max_len)
    g.upload_level(0, lv)
    s64, n64, p64 = g.debug_trace_rays(cells, dirs, org, max_path=3000)
    assert np.array_equal(n64, ref_n)
    for t in range(len(cells)):
        assert np.array_equal(p64[t, :ref_n[t], :3], ref[t][2]), t
    np.testing.assert_allclose(s64, ref_s, rtol=1e-11)
    g32 = make(gpu_lib, threshold=0.001, allowReflect=1, skipZeroWeightLegs=0, maxRayLength=max_len)
    g32.upload_level(0, lv)
    s32, n32, _ = g32.debug_trace_rays(cells, dirs, org)
    assert (n32 == ref_n).mean() > 0.97
    same = n32 == ref_n
    np.testing.assert_allclose(s32[same], ref_s[same], rtol=5e-4, atol=1e-6 * np.abs(ref_s).max())


@pytest.mark.parametrize("half", [False, True])
def test_bounded_ray_length_solve(gpu_lib, po, half):
    """RMCRT_bm1_1L_bounded.ups pattern: whole solve with <ROI_extents type="boundedRayLength">, deterministic mode to
    round-off (FP64 march) and the FP32 production kernels (float and half-packed records) on the same rays"""
    lv = po.burns_christon(16, dtype=np.float32)
    kw = dict(nDivQRays=6, nFluxRays=4, threshold=0.05, isSeedRandom=0, maxRayLength=0.25, solveBoundaryFlux=1)
    ref = po.ray_trace(po.default_params(**kw), lv)
    full = po.ray_trace(po.default_params(**dict(kw, maxRayLength=1e308)), lv)
    assert ref["steps"] < 0.8 * full["steps"]
    scale = np.abs(ref["divQ"]).max()
    if not half:
        g = make(gpu_lib, marchPrecision=1, skipZeroWeightLegs=0, **kw)
        g.upload_level(0, lv)
        out = g.trace_divq(0)
        c = g.counters()
        assert c.nRaySteps == ref["steps"] and c.nFluxRaySteps == ref["fluxSteps"]
        assert np.abs(out["divQ"] - ref["divQ"]).max() <= 1e-10 * scale
        np.testing.assert_allclose(out["boundFlux"], ref["boundFlux"], rtol=1e-10, atol=1e-10 * np.abs(ref["boundFlux"]).max())
    g32 = make(gpu_lib, skipZeroWeightLegs=0, half_records=half, **kw)
    g32.upload_level(0, lv)
    o32 = g32.trace_divq(0)
    tol = 5e-3 if half else 3e-4          # fp16 properties vs fp32
    assert np.abs(o32["divQ"] - ref["divQ"]).max() <= tol * scale
    assert abs(int(g32.counters().nRaySteps) - ref["steps"]) <= 2e-3 * ref["steps"]


def test_bounded_ray_length_refused_for_data_onion(gpu_lib, po):
    from uintah_b200 import RmcrtError
    fine = po.burns_christon(16)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    g = make(gpu_lib, nDivQRays=2, maxRayLength=0.2, haloCells=(2, 2, 2))
    g.upload_level(0, coarse)
    g.upload_level(1, fine)
    with pytest.raises(RmcrtError, match="single-level"):
        g.trace_dataonion(2)


# ------------------------------------------------------------------------------------------------ virtual radiometer
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_radiometer_deterministic_matches_oracle(gpu_lib, po, dtype):
    """Radiometer::radiometerFlux (Radiometer.cc:428-560), <randomSeed> false: per-ray reseed (i+j+k)*iRay+1, 2 draws"""
    lv = po.parallel_plates(14, dtype=dtype)
    lv.sigmaT4[:, :, -1] *= 0.25                      # an x+ wall that differs, so the orientation matters
    p = po.default_params(threshold=0.001, isSeedRandom=0)
    specs = [(30.0, (0, 0, 1), False, (6, 6, 6), (8, 7, 7), 50), (120.0, (1, -2, 0.5), True, (2, 3, 4), (3, 5, 5), 33),
             (60.0, (-1, 1, 0), False, (10, 10, 0), (12, 11, 1), 40)]
    g = make(gpu_lib, threshold=0.001, isSeedRandom=0, marchPrecision=1, skipZeroWeightLegs=0)
    g.upload_level(0, lv)
    rads_o = [po.radiometer(v, n, lo, hi, nr, is_normal_vector=nv) for v, n, nv, lo, hi, nr in specs]
    rads_g = [g.radiometer(v, n, lo, hi, nr, is_normal_vector=nv) for v, n, nv, lo, hi, nr in specs]
    for a, b in zip(rads_o, rads_g):                  # the angle derivation itself (Radiometer.cc:214-260)
        for f in ("solidAngle", "theta_viewAngle", "range", "phi_rotate", "theta_rotate", "xi_rotate"):
            assert getattr(a, f) == pytest.approx(getattr(b, f), rel=1e-15, abs=1e-300)
    fo, io, steps = po.radiometer_flux(p, lv, rads_o)
    fg, ig = g.trace_radiometer(rads_g, dtype=dtype)
    assert g.counters().nRadiometerSteps == steps
    assert np.array_equal(fg != 0, fo != 0) and np.count_nonzero(fo) == 2 + 2 + 2
    tol = 1e-10 if dtype == np.float64 else 3e-7
    np.testing.assert_allclose(fg, fo, rtol=tol)
    np.testing.assert_allclose(ig, io, rtol=tol)
    # patch restriction: only the cells of (radiometer box n patch) are written, the others keep their values
    keepF = np.full(fo.shape, 7.0, dtype); keepI = np.full(fo.shape, 9.0, dtype)
    g.trace_radiometer(rads_g, patch_lo=(0, 0, 0), patch_hi=(7, 14, 14), dtype=dtype, zero_init=False, outputs=(keepF, keepI))
    sel = np.zeros(fo.shape, bool)
    sel[7, 7, 7] = True            # cell (6,6,6) of radiometer 0
    sel[5:6, 4:6, 3] = True        # radiometer 1: x 2, y 3..4, z 4
    assert np.array_equal(keepF != 7.0, sel)
    np.testing.assert_allclose(keepF[sel], fo[sel], rtol=tol)


def test_radiometer_production_closed_form(gpu_lib, po):
    """Philox rays, FP32 march: isothermal black enclosure => VRIntensity = I (1 + cos(viewAngle/2)) / 2"""
    lv = po.parallel_plates(16, dtype=np.float32, T_gas=1000.0, T_wall=1000.0, mirror_faces=())
    I = float(lv.sigmaT4[5, 5, 5])
    for half in (False, True):
        g = make(gpu_lib, threshold=0.0001, allowReflect=0, half_records=half)
        g.upload_level(0, lv)
        for view, vec in ((40.0, (0, 1, 0)), (150.0, (1, 1, 1))):
            r = g.radiometer(view, vec, (8, 8, 8), (9, 9, 9), 20000, is_normal_vector=True)
            flux, inten = g.trace_radiometer([r], dtype=np.float32)
            hv = np.radians(view / 2)
            assert inten[9, 9, 9] == pytest.approx(I * (1 + np.cos(hv)) / 2, rel=(2e-3 if half else 5e-4))
            assert flux[9, 9, 9] == pytest.approx(inten[9, 9, 9] * 2 * np.pi * (1 - np.cos(hv)), rel=1e-6)


# ---------------------------------------------------------------- RMCRT_coarseLevel: Ray::refine_Q on the device
@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("rr", [(2, 2, 2), (4, 4, 4), (2, 4, 1)])
def test_refine_Q_bit_exact(gpu_lib, po, order, rr):
    """selectInterpolator (Core/Grid/AMR.h) on the device: FP64, the reference's operation order -> identical bits"""
    from test_oracle import _refine_case
    q, glo, ghi, fl, fh, olo, ohi = _refine_case(rr)
    g = make(gpu_lib)
    ref = po.refine_Q(q, glo, ghi, glo, ghi, rr, order, fl, fh, olo, ohi)
    out = g.refine_Q(q, glo, ghi, glo, ghi, rr, order, fl, fh, olo, ohi)
    assert np.array_equal(out, ref)
    # a coarse window cut down to what the interpolation reads (the getRegion box of getCoarseLevelRange) gives the same result
    reach = 0 if order == 0 else 1
    clo = tuple(fl[d] // rr[d] - reach for d in range(3)); chi = tuple((fh[d] - 1) // rr[d] + 1 + reach for d in range(3))
    sub = q[clo[2] + 1:chi[2] + 1, clo[1] + 1:chi[1] + 1, clo[0] + 1:chi[0] + 1]
    assert np.array_equal(g.refine_Q(sub, clo, chi, glo, ghi, rr, order, fl, fh, olo, ohi), ref)


def test_refine_Q_errors(gpu_lib, po):
    from test_oracle import _refine_case
    from uintah_b200 import RmcrtError
    q, glo, ghi, fl, fh, olo, ohi = _refine_case((2, 2, 2))
    g = make(gpu_lib)
    with pytest.raises(RmcrtError, match="orderOfInterpolation"):
        g.refine_Q(q, glo, ghi, glo, ghi, (2, 2, 2), 3, fl, fh, olo, ohi)
    with pytest.raises(RmcrtError, match="does not cover"):
        g.refine_Q(q[1:], (-1, -1, 0), ghi, glo, ghi, (2, 2, 2), 1, (0, 0, 0), (4, 4, 4), (0, 0, 0), (4, 4, 4))


def test_coarse_level_solve_then_refine(gpu_lib, po):
    """<algorithm type="RMCRT_coarseLevel"> (RMCRT_Test.cc:464-519): coarsen the fine properties, ray trace the coarse level,
    interpolate divQ back to the fine level -- every stage against the oracle in deterministic mode"""
    N, RR = 32, 4
    fine = po.burns_christon(N)
    fine.rr = (RR,) * 3
    coarse = po.coarsen_level(fine, (RR,) * 3)
    p = po.default_params(nDivQRays=5, threshold=0.05, isSeedRandom=0)
    ref_c = po.ray_trace(p, coarse)["divQ"]
    g = make(gpu_lib, nDivQRays=5, threshold=0.05, isSeedRandom=0, marchPrecision=1, skipZeroWeightLegs=0)
    g.upload_level(0, coarse)
    dq_c = g.trace_divq(0)["divQ"]
    np.testing.assert_allclose(dq_c, ref_c, rtol=1e-9, atol=1e-10 * np.abs(ref_c).max())
    n = N // RR
    for order in (1, 2):
        ref_f = po.refine_Q(ref_c, coarse.lo, coarse.hi, coarse.lo, coarse.hi, (RR,) * 3, order, (0,) * 3, (N,) * 3, fine.lo, fine.hi)
        out_f = g.refine_Q(ref_c, coarse.lo, coarse.hi, coarse.lo, coarse.hi, (RR,) * 3, order, (0,) * 3, (N,) * 3, fine.lo, fine.hi)
        assert np.array_equal(out_f, ref_f)
        # the interpolated coarse solve approximates the fine-level solve (same physics, 64x fewer ray origins)
        assert out_f.shape == fine.abskg.shape
    fine_ref = po.ray_trace(po.default_params(nDivQRays=20, threshold=0.05), fine, seed=9, patches=(1, 1, 8), threads=8)["divQ"]
    inner = (slice(1 + RR, -1 - RR),) * 3
    assert np.linalg.norm(out_f[inner] - fine_ref[inner]) / np.linalg.norm(fine_ref[inner]) < 0.25
    assert n == coarse.interior_hi[0]


# ---------------------------------------------------------------- the Arches caller's property tasks on the device
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_arches_property_tasks_match_oracle(gpu_lib, po, dtype):
    """RMCRT_Radiation::sigmaT4 (gas only and gas + 3 particle classes), sumAbsk and stencilToDBLs: bit-exact"""
    rng = np.random.default_rng(11)
    shp = (12, 11, 13)                      # a 10x9x11-cell patch with its extra cells
    lo, hi, ilo, ihi = (-1, -1, -1), (12, 10, 11), (0, 0, 0), (11, 9, 10)
    gv = (rng.uniform(size=shp) > 0.1).astype(np.float64)       # 10 % solid cells
    temp = [rng.uniform(300, 2000, shp) for _ in range(4)]
    absk = [rng.uniform(0, 2, shp) * (rng.uniform(size=shp) > 0.2) for _ in range(4)]
    zero = rng.uniform(size=shp) < 0.05
    for a in absk:
        a[zero] = 0.0                       # cells with sumAbsk == 0 -> sigmaT4 = 0 (RMCRT.cc:1000-1003)
    g = make(gpu_lib)
    for part in (False, True):
        ref = po.arches_sigmaT4(gv, absk, temp, part, dtype)
        out = g.arches_sigmaT4(gv, absk, temp, part, dtype)
        assert np.array_equal(out, ref) and out.dtype == dtype
    assert (po.arches_sigmaT4(gv, absk, temp, True, dtype) == 0).sum() >= int((zero & (gv > 0)).sum()) > 0
    ref = po.arches_sumAbsk(gv, absk, lo, hi, ilo, ihi, dtype)
    out = g.arches_sumAbsk(gv, absk, lo, hi, ilo, ihi, dtype)
    assert np.array_equal(out, ref) and (out[0] == 1).all() and (out[:, :, -1] == 1).all()
    bf = rng.normal(size=shp + (7,))
    r6, o6 = po.arches_stencilToDBLs(bf), g.stencil_to_dbls(bf)
    for k in "EWNSTB":
        assert np.array_equal(o6[k], r6[k])
    assert np.array_equal(o6["E"], bf[..., 1]) and np.array_equal(o6["B"], bf[..., 4])


# ---------------------------------------------------------------- isotropic scattering (<sigmaScat> > 0; RMCRT_isoScat.ups)
@pytest.mark.parametrize("mode", ["reseed", "fixed", "reseed_float_lhc"])
def test_scattering_deterministic_matches_oracle(gpu_lib, po, mode):
    """RMCRT_isoScat.ups / RMCRT_isoScat_LHC pattern in the regression modes: divQ and boundary fluxes of the FP64 scattering march against
    the oracle (itself bit-exact against the reference's --enable-ray-scatter build, tests/test_oracle_ref.py)"""
    dtype = np.float32 if "float" in mode else np.float64
    lv = po.burns_christon(15, dtype=dtype)
    add = dict(fixedRandom=1) if mode == "fixed" else {}
    lhc = dict(rayDirSampleAlgo=1) if "lhc" in mode else {}
    kw = dict(nDivQRays=4, nFluxRays=6, threshold=0.05, isSeedRandom=0, solveBoundaryFlux=1, sigmaScat=2.0, **add, **lhc)
    ref = po.ray_trace(po.default_params(**kw), lv)
    g = make(gpu_lib, marchPrecision=1, skipZeroWeightLegs=0, **kw)
    g.upload_level(0, lv)
    out = g.trace_divq(0)
    c = g.counters()
    # libm vs CUDA log / sincos differ in the last bits: a DDA tie can flip for a few rays out of 2e4
    assert abs(int(c.nRaySteps) - ref["steps"]) <= 2e-4 * ref["steps"], (c.nRaySteps, ref["steps"])
    assert abs(int(c.nFluxRaySteps) - ref["fluxSteps"]) <= 2e-4 * ref["fluxSteps"]
    scale = np.abs(ref["divQ"]).max()
    bad = np.abs(out["divQ"] - ref["divQ"]) > 1e-9 * scale
    assert bad.mean() < 2e-3, bad.mean()
    assert np.array_equal(out["boundFlux"][..., 6], ref["boundFlux"][..., 6])
    fb = np.abs(out["boundFlux"][..., :6] - ref["boundFlux"][..., :6]) > 1e-9 * np.abs(ref["boundFlux"][..., :6]).max()
    assert fb.mean() < 2e-3
    # scattering really happened: the same solve without it marches a different number of steps
    p0 = po.default_params(**{**kw, "sigmaScat": 0.0})
    assert po.ray_trace(p0, lv)["steps"] != ref["steps"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_scattering_production_properties(gpu_lib, po, dtype):
    """Philox rays: (i) isotropic scattering cannot change the radiation field of an isothermal black enclosure: divQ == 0 to round-off
    whatever sigmaScat is; (ii) Burns & Christon with scattering agrees with the oracle's scattering solve within Monte-Carlo error and
    differs from the non-scattering solution"""
    N = 24
    rng = np.random.default_rng(4)
    M = N + 2
    ab = rng.uniform(0.1, 2.0, (M, M, M)).astype(dtype)
    s4 = np.full((M, M, M), 0.6, dtype)
    ct = np.full((M, M, M), -1, np.int32)
    for ax in range(3):
        for idx in (0, M - 1):
            sl = [slice(None)] * 3; sl[ax] = idx
            ct[tuple(sl)] = 8; ab[tuple(sl)] = 1.0
    iso = po.Level(ab, s4, ct, (-1,) * 3, (N + 1,) * 3, (1.0 / N,) * 3, interior_lo=(0,) * 3, interior_hi=(N,) * 3)
    g = make(gpu_lib, nDivQRays=8, threshold=0.001, seed=5, allowReflect=0, sigmaScat=6.0)
    g.upload_level(0, iso)
    out = g.trace_divq(0)
    assert np.abs(out["divQ"]).max() < 2e-5 * 4 * np.pi * 2.0 * 0.6          # FP32 production march whatever the property type
    n_scat = g.counters().nRaySteps
    g0 = make(gpu_lib, nDivQRays=8, threshold=0.001, seed=5, allowReflect=0)
    g0.upload_level(0, iso); g0.trace_divq(0)
    assert n_scat != g0.counters().nRaySteps
    # Burns & Christon, sigmaScat = 3
    lv = po.burns_christon(21, dtype=dtype)
    p = po.default_params(nDivQRays=60, threshold=0.01, sigmaScat=3.0)
    runs = [po.ray_trace(p, lv, seed=10 * s, patches=(1, 1, 7), threads=7)["divQ"] for s in range(1, 5)]
    mu = np.mean(runs, 0)
    spread = max(np.linalg.norm(r - mu) / np.linalg.norm(mu) for r in runs)
    gs = make(gpu_lib, nDivQRays=60, threshold=0.01, seed=8, sigmaScat=3.0, rayCoherence=0)
    gs.upload_level(0, lv)
    d = gs.trace_divq(0)["divQ"]
    assert np.linalg.norm(d - mu) / np.linalg.norm(mu) < 1.5 * spread, (np.linalg.norm(d - mu) / np.linalg.norm(mu), spread)
    noscat = po.ray_trace(po.default_params(nDivQRays=60, threshold=0.01), lv, seed=3, patches=(1, 1, 7), threads=7)["divQ"]
    assert np.linalg.norm(noscat - mu) / np.linalg.norm(mu) > 2 * spread          # scattering is a visible effect at this optical thickness


# ------------------------------------------------------------------------------------------------ coarsening (Ray::coarsen_Q, computeCellType)
def _device_coarse_level(g, fine, coarse_like, dtype):
    """coarsen level 1 (uploaded from `fine`) into level 0 on the device and read its records back: (abskg, sigmaT4, nonFlow)"""
    from uintah_b200 import LevelData
    g.upload_level(1, fine)
    g.define_level(0, LevelData.of(coarse_like).desc())
    g.levels[0] = LevelData.of(coarse_like)
    g.coarsen_level_device(1, 0)
    rec, rb = g.level_records(0)
    raw = rec.cpu().numpy()
    a = raw.view(np.float64 if rb == 16 else np.float32).reshape(coarse_like.abskg.shape + (2,))
    k, s = a[..., 0], a[..., 1]
    return np.abs(k), s, np.signbit(k)


@pytest.mark.parametrize("rr,logic,wtype", [((2, 2, 2), 0, 8), ((4, 4, 4), 0, 8), ((2, 2, 2), 1, 8), ((4, 4, 4), 1, 10), ((2, 4, 1), 0, 8), ((2, 2, 2), 1, 10)])
def test_coarsened_cell_type_follows_the_integer_sum(gpu_lib, po, rr, logic, wtype):
    """Ray::computeCellType (Ray.cc:2272-2345) sums the INTEGER cell types of the fine cells (flow -1, RMCRT_Test walls 8): with RR = 8 and
    ONE non-flow fine cell the sum is 1 and ROUNDUP keeps the coarse cell a FLOW cell (|sum| < 10); RR = 64 with five or six likewise.
    Partial intrusions of every fill level against the oracle's restatement of the reference logic."""
    rng = np.random.default_rng(12)
    N = 32
    fine = po.burns_christon(N, dtype=np.float64)
    fine.rr = rr
    ct = fine.cellType
    RR = rr[0] * rr[1] * rr[2]
    # every coarse cell gets a random number 0..RR of non-flow fine cells
    for kc in range(N // rr[2]):
        for jc in range(N // rr[1]):
            for ic in range(N // rr[0]):
                u = rng.uniform()
                nw = RR if u < 0.1 else (int(rng.integers(0, RR + 1)) if u < 0.55 else 0)
                if nw:
                    idx = rng.choice(RR, nw, replace=False)
                    for q in idx:
                        a, b, c = q % rr[0], (q // rr[0]) % rr[1], q // (rr[0] * rr[1])
                        ct[1 + kc * rr[2] + c, 1 + jc * rr[1] + b, 1 + ic * rr[0] + a] = wtype
    ref = po.coarsen_level(fine, rr, logic=logic)
    g = make(gpu_lib, marchPrecision=1, cellTypeCoarsenLogic=logic, nonFlowCellType=wtype)
    k, s, wall = _device_coarse_level(g, fine, ref, np.float64)
    inner = (slice(1, -1),) * 3
    assert np.array_equal(wall[inner], ref.cellType[inner] != -1)
    nwall = int((ref.cellType[inner] != -1).sum())
    assert nwall < ref.cellType[inner].size
    # ROUNDDOWN needs |sum| > 10 (RR - 1): out of reach for the wall id 8 of RMCRT_Test (every coarse cell stays a flow cell, in the
    # reference too), reached by completely filled cells of the intrusion id 10
    assert (nwall > 0) == (logic == 0 or wtype == 10)
    np.testing.assert_allclose(k[inner], ref.abskg[inner], rtol=1e-13)
    np.testing.assert_allclose(s[inner], ref.sigmaT4[inner], rtol=1e-13)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_coarsen_extra_cells(gpu_lib, po, dtype):
    """<coarsenExtraCells> true (Ray::coarsen_Q, Ray.cc:2148-2245): coarse extra cells on the domain faces = tangential average of the fine
    extra cells (ratio 1 along the normal), edges / corners 0, all of them non-flow; then the Data-Onion solve on the device-derived
    coarse level equals the oracle's on the oracle-built one"""
    rng = np.random.default_rng(14)
    fine = po.parallel_plates(32, dtype=dtype, mirror_faces=())
    fine.rr = (4, 4, 4)
    wall = fine.cellType != -1
    fine.abskg[wall] = rng.uniform(0.2, 1.0, int(wall.sum())).astype(dtype)          # walls that vary along the faces
    fine.sigmaT4[wall] = (fine.sigmaT4[wall] * rng.uniform(0.5, 1.5, int(wall.sum()))).astype(dtype)
    ref = po.coarsen_level(fine, (4, 4, 4), coarsen_extra_cells=True)
    prec = 1 if dtype == np.float64 else 0
    g = make(gpu_lib, marchPrecision=prec, coarsenExtraCells=1, haloCells=(4, 4, 4), nDivQRays=3, threshold=0.01, isSeedRandom=0, skipZeroWeightLegs=0)
    k, s, nonflow = _device_coarse_level(g, fine, ref, dtype)
    M = ref.abskg.shape[0]
    ext = np.ones(ref.abskg.shape, bool); ext[1:-1, 1:-1, 1:-1] = False
    assert nonflow[ext].all()
    tol = dict(rtol=1e-13) if dtype == np.float64 else dict(rtol=2e-7)
    np.testing.assert_allclose(k[ext], ref.abskg[ext], **tol)
    np.testing.assert_allclose(s[ext], ref.sigmaT4[ext], **tol)
    edges = ext.copy()
    for a in range(3):
        sl = [slice(1, -1)] * 3
        for idx in (0, M - 1):
            sl[a] = idx; edges[tuple(sl)] = False
    assert np.all(k[edges] == 0) and np.all(s[edges] == 0) and (k[ext & ~edges] > 0).all()
    if dtype == np.float64:          # the solve itself (deterministic regression mode) on the device-derived coarse level
        g.finalize_level(0)
        kw = dict(nDivQRays=3, threshold=0.01, isSeedRandom=0)
        want = po.ray_trace_data_onion(po.default_params(**kw), [ref, fine], (4, 4, 4), patches=(2, 2, 2), threads=8)
        divQ = np.zeros_like(want["divQ"]); vol = np.zeros_like(divQ)
        for plo, phi in po._split((0, 0, 0), (32, 32, 32), (2, 2, 2)):
            g.trace_dataonion(2, plo, phi, zero_init=False, outputs=(divQ, vol, None))
        assert g.counters().nRaySteps == want["steps"]
        np.testing.assert_allclose(divQ, want["divQ"], rtol=1e-9, atol=1e-10 * np.abs(want["divQ"]).max())


@pytest.mark.parametrize("prec", [1, 0])
def test_data_onion_fixed_roi(gpu_lib, po, prec):
    """<ROI_extents type="fixed">: one ROI box for every fine patch (Ray.cc:1303-1314) -- patches inside it (ROI-copy fast path in FP32),
    straddling it and outside it (their rays start on the fine level and leave it at the first step); extents equal the oracle's,
    deterministic solve to round-off (FP64) / FP32 round-off"""
    fine = po.burns_christon(32, dtype=np.float64 if prec else np.float32)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    rlo, rhi = (6, 4, 8), (26, 30, 24)
    kw = dict(nDivQRays=3, threshold=0.05, isSeedRandom=0)
    want = po.ray_trace_data_onion(po.default_params(roiFixed=1, roiFixedLo=rlo, roiFixedHi=rhi, **kw), [coarse, fine], (4, 4, 4), patches=(2, 2, 2), threads=8)
    g = make(gpu_lib, marchPrecision=prec, skipZeroWeightLegs=0, haloCells=(4, 4, 4), roiFixed=1, roiFixedLo=rlo, roiFixedHi=rhi, **kw)
    g.upload_level(0, coarse)
    g.upload_level(1, fine)
    for plo, phi in (((0, 0, 0), (16, 16, 16)), ((8, 8, 8), (24, 24, 24))):
        a = g.compute_extents(2, plo, phi)
        b = po.compute_extents([coarse, fine], plo, phi, (4, 4, 4), rlo, rhi)
        assert a[0] == b[0] == rlo and a[1] == b[1] == rhi and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    divQ = np.zeros_like(want["divQ"]); vol = np.zeros_like(divQ)
    for plo, phi in list(po._split((0, 0, 0), (32, 32, 32), (2, 2, 2))) + [((8, 8, 8), (24, 24, 24))]:      # the last patch lies inside the ROI
        g.trace_dataonion(2, plo, phi, zero_init=False, outputs=(divQ, vol, None))
    if prec:
        np.testing.assert_allclose(divQ, want["divQ"], rtol=1e-9, atol=1e-10 * np.abs(want["divQ"]).max())
    else:
        assert np.abs(divQ - want["divQ"]).max() <= 2e-4 * np.abs(want["divQ"]).max()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_dynamic_roi_extents(gpu_lib, po, dtype):
    """Ray::ROI_Extents (Ray.cc:1984-2043): gradient magnitudes, flags and the ROI_loCell / ROI_hiCell reduction bit for bit against the
    numpy restatement (float records for float decks, double records for the FP64 march); then the Data-Onion solve with that ROI"""
    rng = np.random.default_rng(21)
    fine = po.burns_christon(32, dtype=dtype)
    fine.rr = (4, 4, 4)
    # a flat medium with a hot, optically thick blob inside the first octant: only the patch it lies in gets flagged (the walls are given
    # the medium's sigmaT4 and nearly its abskg so that they do not flag every patch)
    z, y, x = np.meshgrid(*[np.arange(-1, 33)] * 3, indexing="ij")
    blob = np.exp(-((x - 9.0) ** 2 + (y - 8.0) ** 2 + (z - 10.0) ** 2) / 18.0)
    base_s = float(fine.sigmaT4[5, 5, 5])
    fine.abskg[...] = (0.9 + 3.0 * blob).astype(dtype)
    fine.abskg[fine.cellType != -1] = 1.0
    fine.sigmaT4[...] = (base_s * (1 + 4.0 * blob)).astype(dtype)
    fine.sigmaT4[fine.cellType != -1] = base_s
    coarse = po.coarsen_level(fine, (4, 4, 4))
    patches = [(lo, hi, tuple(l - (1 if l == 0 else 0) for l in lo), tuple(h + (1 if h == 32 else 0) for h in hi))
               for lo, hi in po._split((0, 0, 0), (32, 32, 32), (2, 2, 2))]
    thrA, thrS = 8.0, 1.0e9
    wantA, wantS, wantF, (rlo, rhi) = po.roi_extents(fine, patches, thrA, thrS)
    prec = 1 if dtype == np.float64 else 0
    kw = dict(nDivQRays=3, threshold=0.05, isSeedRandom=0)
    g = make(gpu_lib, marchPrecision=prec, skipZeroWeightLegs=0, haloCells=(4, 4, 4), **kw)
    g.upload_level(0, coarse)
    g.upload_level(1, fine)
    gotA = np.zeros_like(wantA); gotS = np.zeros_like(wantS); gotF = np.zeros_like(wantF)
    lo_, hi_ = [32767] * 3, [-32767] * 3
    for lo, hi, elo, ehi in patches:
        a, s_, f, any_ = g.roi_extents(1, lo, hi, thrA, thrS, dtype=dtype)
        gotA += a; gotS += s_; gotF += f
        assert any_ == bool(f.any())
        if any_:
            lo_ = [min(p, q) for p, q in zip(lo_, elo)]; hi_ = [max(p, q) for p, q in zip(hi_, ehi)]
    assert np.array_equal(gotF, wantF) and 0 < wantF.sum() < 32 ** 3
    assert np.array_equal(gotA, wantA) and np.array_equal(gotS, wantS)
    assert (tuple(lo_), tuple(hi_)) == (rlo, rhi)
    assert (rlo, rhi) == ((-1, -1, -1), (16, 16, 33))             # the two patches the blob reaches (of eight)
    # the solve with the dynamic ROI = the fixed-ROI solve with that box (Ray::computeExtents: one box for every patch)
    want = po.ray_trace_data_onion(po.default_params(roiFixed=1, roiFixedLo=rlo, roiFixedHi=rhi, **kw), [coarse, fine], (4, 4, 4), patches=(2, 2, 2), threads=8)
    g.set_params(roiFixed=1, roiFixedLo=rlo, roiFixedHi=rhi)
    divQ = np.zeros_like(want["divQ"]); vol = np.zeros_like(divQ)
    for lo, hi, _, _ in patches:
        g.trace_dataonion(2, lo, hi, zero_init=False, outputs=(divQ, vol, None))
    if prec:
        np.testing.assert_allclose(divQ, want["divQ"], rtol=1e-9, atol=1e-10 * np.abs(want["divQ"]).max())
    else:
        assert np.abs(divQ - want["divQ"]).max() <= 3e-4 * np.abs(want["divQ"]).max()
