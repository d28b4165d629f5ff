"""The oracle against the REFERENCE'S OWN per-ray code.

`make -C oracle ref` (run by __graft_entry__.build() where /root/reference exists) cuts RMCRTCommon::raySignStep / findRayDirection /
ray_Origin / reflect / updateSumI<T> and Ray::containsCell / updateSumI_ML<T> out of the reference's RMCRTCommon.cc and Ray.cc
(oracle/extract_ref.py) and compiles them, unmodified, with the reference's own MersenneTwister.h into oracle/_ref/libref_rmcrt.so
(oracle/ref_rmcrt_shim.cpp; git-ignored, travels to the GPU box).  These tests feed the same rays through that library and through the
oracle's restatement: sumI must agree BIT FOR BIT and the step counts exactly -- which pins every line of the marching routines the CUDA
path is checked against.
"""
import numpy as np
import pytest

from util import add_random_walls, random_rays


@pytest.fixture(scope="module")
def ref(po):
    if po.ref_lib() is None:
        pytest.skip("oracle/_ref/libref_rmcrt.so was not built (no reference tree at build time)")
    return po.ref_lib()


def cases(po, dtype):
    rng = np.random.default_rng(7)
    bc = po.burns_christon(24, dtype=dtype)
    add_random_walls(bc, 0.02, rng, abskg=0.6, sigmaT4=0.2)
    pl = po.parallel_plates(20, dtype=dtype, mirror_faces=("x-", "y-", "z+"))
    add_random_walls(pl, 0.01, rng, abskg=0.3, sigmaT4=5000.0)
    M = 18
    rnd = po.Level(rng.uniform(0.05, 3.0, (M, M, M)).astype(dtype), rng.uniform(0, 2, (M, M, M)).astype(dtype), np.full((M, M, M), -1, np.int32),
                   (-1,) * 3, (M - 1,) * 3, (0.07, 0.05, 0.11), anchor=(-0.3, 0.2, 1.0), interior_lo=(0,) * 3, interior_hi=(M - 2,) * 3)
    for ax in range(3):
        for idx in (0, M - 1):
            sl = [slice(None)] * 3; sl[ax] = idx
            rnd.cellType[tuple(sl)] = 8
    rnd.abskg[rnd.cellType != -1] = rng.uniform(0, 1.4, int((rnd.cellType != -1).sum())).astype(dtype)     # emissivities above 1 get clamped
    return {"burns_christon+intrusions": bc, "plates+mirrors": pl, "anisotropic_random": rnd}


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kw", [dict(threshold=0.05), dict(threshold=0.001, allowReflect=0), dict(threshold=0.01, maxRayLength=0.35)])
def test_updateSumI_bit_exact(po, ref, dtype, kw):
    """RMCRTCommon::updateSumI<T> (RMCRTCommon.cc:607-904): 1500 random rays per field, thresholds / reflections / bounded length"""
    rng = np.random.default_rng(11)
    p = po.default_params(**kw)
    total = 0
    for name, lv in cases(po, dtype).items():
        cells, dirs, orgs = random_rays(lv, 1500, rng)
        for c, d, o in zip(cells, dirs, orgs):
            if lv.cellType[c[2] - lv.lo[2], c[1] - lv.lo[1], c[0] - lv.lo[0]] != -1:
                continue
            s_ref, n_ref, rc = po.ref_update_sumI(p, lv, d, o, c)
            s, n, _ = po.update_sumI(p, lv, d, o, c)
            assert rc == 0
            assert n == n_ref, (name, c, d)
            assert s == s_ref or (np.isnan(s) and np.isnan(s_ref)), (name, c, d, s, s_ref)      # bit for bit
            total += n
    assert total > 30000


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_updateSumI_ML_bit_exact(po, ref, dtype):
    """Ray::updateSumI_ML<T> (Ray.cc:2349-2627) on a 3-level Data Onion: level jumps, region boxes, reflections on every level"""
    rng = np.random.default_rng(13)
    fine = po.parallel_plates(32, dtype=dtype, mirror_faces=("x+", "z-"))
    fine.abskg[1:-1, 1:-1, 1:-1] = rng.uniform(0.05, 1.5, (32, 32, 32)).astype(dtype)
    fine.sigmaT4[1:-1, 1:-1, 1:-1] = rng.uniform(0, 1e4, (32, 32, 32)).astype(dtype)
    walls = fine.cellType != -1
    fine.abskg[walls] = rng.uniform(0, 1.3, int(walls.sum())).astype(dtype)     # grey walls: fs = fs (1 - abskg) matters, > 1 gets clamped
    fine.rr = (2, 2, 2)
    mid = po.coarsen_level(fine, (2, 2, 2))
    mid.rr = (4, 4, 4)
    coarse = po.coarsen_level(mid, (4, 4, 4))
    levels = [coarse, mid, fine]
    for thr in (0.05, 0.002):
        p = po.default_params(threshold=thr)
        for patch_lo, patch_hi in (((0, 0, 0), (16, 16, 16)), ((16, 8, 16), (32, 24, 32))):
            roi_lo, roi_hi, reg_lo, reg_hi = po.compute_extents(levels, patch_lo, patch_hi, (3, 3, 3))
            cells, dirs, orgs = random_rays(fine, 600, rng, lo=patch_lo, hi=patch_hi)
            for c, d, o in zip(cells, dirs, orgs):
                s_ref, n_ref, rc = po.ref_update_sumI_ML(p, levels, roi_lo, roi_hi, reg_lo, reg_hi, d, o, c)
                s, n, _ = po.update_sumI_ML(p, levels, roi_lo, roi_hi, reg_lo, reg_hi, d, o, c)
                assert rc == 0 and n == n_ref and s == s_ref, (thr, c, d, n, n_ref, s, s_ref)


@pytest.mark.parametrize("kw", [dict(isSeedRandom=1), dict(isSeedRandom=0), dict(isSeedRandom=1, CCRays=1)])
def test_ray_generation_bit_exact(po, ref, kw):
    """findRayDirection + ray_Origin as the divQ loop calls them (Ray.cc:752-760; RMCRTCommon.cc:499-576), random and per-ray re-seeded"""
    lv = po.burns_christon(12)
    lv2 = po.Level(lv.abskg, lv.sigmaT4, lv.cellType, lv.lo, lv.hi, (0.07, 0.05, 0.11), anchor=(-0.3, 0.2, 1.0), interior_lo=(0,) * 3, interior_hi=(12,) * 3)
    p = po.default_params(**kw)
    for level in (lv, lv2):
        for cell in ((0, 0, 0), (3, 7, 11), (11, 11, 11)):
            for seed in (1, 12345):
                d0, o0 = po.generate_rays(p, level, cell, 300, seed, reference=True)
                d1, o1 = po.generate_rays(p, level, cell, 300, seed)
                assert np.array_equal(d0, d1) and np.array_equal(o0, o1)
    np.testing.assert_allclose(np.linalg.norm(d0, axis=1), 1.0, rtol=1e-12)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("seed_random", [1, 0])
def test_divQ_loop_bit_exact(po, ref, dtype, seed_random):
    """the divQ loop of Ray::rayTrace<T> (Ray.cc:742-779) driven by the reference's own per-ray routines == oracle_rayTrace, bit for bit,
    including the MT19937 stream shared by all cells of a task and the task split / seeds"""
    lv = po.burns_christon(14, dtype=dtype)
    add_random_walls(lv, 0.03, np.random.default_rng(3), abskg=0.5, sigmaT4=0.1)
    p = po.default_params(nDivQRays=7, threshold=0.02, isSeedRandom=seed_random)
    a = po.ref_ray_trace(p, lv, seed=5, patches=(1, 2, 2), threads=4)
    b = po.ray_trace(p, lv, seed=5, patches=(1, 2, 2), threads=4)
    assert a["steps"] == b["steps"] and a["steps"] > 1e5
    assert np.array_equal(a["divQ"], b["divQ"]) and np.array_equal(a["radiationVolq"], b["radiationVolq"])


# ---------------------------------------------------------------- isotropic scattering: the reference's --enable-ray-scatter build
@pytest.fixture(scope="module")
def ref_scatter(po):
    lib = po.ref_scatter_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libref_rmcrt_scatter.so was not built (no reference tree at build time)")
    assert lib.ref_has_ray_scatter() == 1 and po.ref_lib().ref_has_ray_scatter() == 0
    return lib


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kw", [dict(sigmaScat=1.0, threshold=0.05), dict(sigmaScat=8.0, threshold=0.01, isSeedRandom=0),
                                dict(sigmaScat=3.0, threshold=0.02, allowReflect=0), dict(sigmaScat=25.0, threshold=0.01, maxRayLength=0.6)])
def test_scattering_rays_bit_exact(po, ref_scatter, dtype, kw):
    """RMCRTCommon::updateSumI<T> compiled with RAY_SCATTER (RMCRTCommon.cc:670-676, 790-836): free-path draws, re-direction (with the
    per-event MT reseed (i+j+k)*(-9)+1 of the regression mode), back scattering into prevCell, tMax rebuilt from the ray location"""
    rng = np.random.default_rng(17)
    p = po.default_params(**kw)
    scattered = 0
    for name, lv in cases(po, dtype).items():
        cells, dirs, orgs = random_rays(lv, 400, rng)
        for n_, (c, d, o) in enumerate(zip(cells, dirs, orgs)):
            if lv.cellType[c[2] - lv.lo[2], c[1] - lv.lo[1], c[0] - lv.lo[0]] != -1:
                continue
            s_ref, n_ref = po.update_sumI_seeded(p, lv, d, o, c, 100 + n_, reference=ref_scatter)
            s, n = po.update_sumI_seeded(p, lv, d, o, c, 100 + n_)
            assert n == n_ref and s == s_ref, (name, c, d, n, n_ref, s, s_ref)
            p0 = po.default_params(**{**kw, "sigmaScat": 0.0})
            scattered += (po.update_sumI(p0, lv, d, o, c)[1] != n)
    assert scattered > 200          # scattering really changed the paths


@pytest.mark.parametrize("seed_random", [1, 0])
def test_scattering_divQ_loop_bit_exact(po, ref_scatter, seed_random):
    """RMCRT_isoScat.ups pattern (benchmark 1, sigmaScat 1): whole-patch divQ with the reference's scatter build == the oracle"""
    lv = po.burns_christon(13)
    p = po.default_params(nDivQRays=5, threshold=0.05, isSeedRandom=seed_random, sigmaScat=1.0)
    a = po.ref_ray_trace(p, lv, seed=9, patches=(1, 1, 2), threads=2, library=ref_scatter)
    b = po.ray_trace(p, lv, seed=9, patches=(1, 1, 2), threads=2)
    assert a["steps"] == b["steps"]
    assert np.array_equal(a["divQ"], b["divQ"]) and np.array_equal(a["radiationVolq"], b["radiationVolq"])
    c = po.ray_trace(po.default_params(nDivQRays=5, threshold=0.05, isSeedRandom=seed_random), lv, seed=9, patches=(1, 1, 2), threads=2)
    assert c["steps"] != b["steps"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("seeded", [1, 0])
def test_boundary_flux_loop_bit_exact(po, ref, dtype, seeded):
    """Ray::has_a_boundary, rayDirection_cellFace, rayLocation_cellFace and the d_dirIndexOrder / d_dirSignSwap tables (Ray.cc:112-125,
    1362-1390, 1465-1572), the reference's own code, under the boundary-flux loop of Ray::rayTrace: the whole Stencil7 field -- face
    selection, boundFlux.p and every flux -- equals the oracle's restatement bit for bit, and so do the executed steps"""
    rng = np.random.default_rng(17)
    for name, lv in cases(po, dtype).items():
        if name == "anisotropic_random":
            add_random_walls(lv, 0.03, rng, abskg=0.5, sigmaT4=0.7)
        p = po.default_params(nDivQRays=1, nFluxRays=7, threshold=0.01, isSeedRandom=seeded, solveBoundaryFlux=1, solveDivQ=0)
        R, nR = po.ref_ray_trace_boundflux(p, lv, seed=31)
        O = po.ray_trace(p, lv, seed=31)
        assert nR == O["fluxSteps"], name
        assert (R[..., 6] == 1).sum() > 100
        assert np.array_equal(R, O["boundFlux"]), (name, np.abs(R - O["boundFlux"]).max())


def test_computeExtents_is_the_reference(po, ref):
    """Ray::computeExtents (Ray.cc:1279-1358), the reference's own code, against the oracle's restatement: patch_based ROI for two- and
    three-level stacks, patches in corners / on faces / inside, several halos; mapCellToCoarser of negative indices included"""
    fine = po.burns_christon(64, dtype=np.float32); fine.rr = (2, 2, 2)
    mid = po.coarsen_level(fine, (2, 2, 2)); mid.rr = (4, 4, 4)
    coarse = po.coarsen_level(mid, (4, 4, 4))
    f2 = po.burns_christon(32, dtype=np.float32); f2.rr = (4, 4, 4)
    c2 = po.coarsen_level(f2, (4, 4, 4))
    n = 0
    for levels, N in (([coarse, mid, fine], 64), ([c2, f2], 32)):
        for halo in ((4, 4, 4), (1, 2, 3), (10, 10, 10), (0, 0, 0)):
            for lo, hi in (((0, 0, 0), (N // 2,) * 3), ((N // 2,) * 3, (N,) * 3), ((N // 4, 0, N // 2), (N // 2, N // 4, N)), ((3, 5, 7), (11, 9, 8))):
                a = po.ref_compute_extents(levels, lo, hi, halo)
                b = po.compute_extents(levels, lo, hi, halo)
                assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), (N, halo, lo, hi, a, b)
                n += 1
    assert n == 32


def test_computeExtents_fixed_roi_is_the_reference(po, ref):
    """the fixed branch of Ray::computeExtents (Ray.cc:1303-1314): ROI = fineLevel->getCellIndex(min / max point), clamped to the level,
    the same for every patch; points off the fine level make the reference throw"""
    fine = po.burns_christon(64, dtype=np.float32); fine.rr = (2, 2, 2)
    mid = po.coarsen_level(fine, (2, 2, 2)); mid.rr = (4, 4, 4)
    coarse = po.coarsen_level(mid, (4, 4, 4))
    levels = [coarse, mid, fine]
    dx = 1.0 / 64
    n = 0
    for pmin, pmax in (((0.25, 0.25, 0.25), (0.75, 0.75, 0.75)), ((0.0, 0.1, 0.3), (0.999, 0.6, 0.31)), ((-0.01, -0.01, -0.01), (1.01, 1.01, 1.01))):
        lo = tuple(int(np.floor(x / dx)) for x in pmin)
        hi = tuple(int(np.floor(x / dx)) for x in pmax)
        for plo, phi in (((0, 0, 0), (32, 32, 32)), ((32, 32, 32), (64, 64, 64))):
            a = po.ref_compute_extents(levels, plo, phi, (4, 4, 4), "fixed", pmin, pmax)
            b = po.compute_extents(levels, plo, phi, (4, 4, 4), lo, hi)
            assert a is not None
            assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), (pmin, pmax, a, b)
            n += 1
    assert n == 6
    assert po.ref_compute_extents(levels, (0, 0, 0), (32, 32, 32), (4, 4, 4), "fixed", (0.2, 0.2, 0.2), (1.5, 0.5, 0.5)) is None
