"""CPU tests that PIN the oracle (oracle/rmcrt_oracle.cpp) against what the reference tree itself holds:
the reference's own MersenneTwister.h (golden fixture + live _ref build), the Burns & Christon exact centre line
(parametricStudies/postProcessTools/RMCRT.m:22-38) and the Modest parallel-plate closed form (RMCRT.m:44-94)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_mt19937_matches_reference_golden(po):
    g = json.load(open(os.path.join(HERE, "golden", "mt19937.json")))
    for seed_s, v in g["seeds"].items():
        seed = int(seed_s)
        u = po.mt_u32(seed, 1300)
        assert [int(x) for x in u[:8]] == v["first8_u32"]
        assert [int(x) for x in u[623:626]] == v["u32_at_623_624_625"]
        assert [int(u[1247]), int(u[1248]), int(u[1299])] == v["u32_at_1247_1248_1299"]
        assert int(np.bitwise_xor.reduce(u)) == v["xor_all_1300"]
        assert [float(x).hex() for x in po.mt_rand(seed, 8)] == v["rand_first8_hex"]
        assert [float(x).hex() for x in po.mt_randDblExc(seed, 8)] == v["randDblExc_first8_hex"]


def test_mt19937_matches_reference_header_live(po):
    so = os.path.join(ROOT, "oracle", "_ref", "libref_mt.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ref = C.CDLL(so)
    for seed in (1, 7, 99991, 2 ** 32 - 1):
        b = np.empty(3000, np.uint32)
        ref.ref_mt_u32(C.c_uint32(seed), 3000, b.ctypes.data_as(C.c_void_p))
        assert (po.mt_u32(seed, 3000) == b).all()


def test_fixed_random_num_live(po):
    """FIXED_RANDOM_NUM build of the reference header returns 0.3 for rand() and randDblExc()"""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_mt_fixed.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built")
    ref = C.CDLL(so)
    b = np.empty(4)
    ref.ref_mt_rand(C.c_uint32(1), 4, b.ctypes.data_as(C.c_void_p))
    assert (b == 0.3).all()
    ref.ref_mt_randDblExc(C.c_uint32(1), 4, b.ctypes.data_as(C.c_void_p))
    assert (b == 0.3).all()


@pytest.fixture(scope="module")
def bc41(po):
    """RMCRT_bm1_1L.ups at 100 rays/cell, randomSeed=true style (C1 of BASELINE.json)"""
    lv = po.burns_christon(41)
    p = po.default_params(nDivQRays=100, threshold=0.05, isSeedRandom=1)
    r = po.ray_trace(p, lv, seed=1, patches=(1, 1, 8), threads=8)
    return lv, p, r


def test_burns_christon_centreline(po, bc41):
    lv, p, r = bc41
    L2, line, exact = po.centreline_L2(r["divQ"], 41)
    # 0.010-0.014 at 100 rays; exact value at x=0.5 is -3.07561
    assert L2 < 0.02, L2
    assert abs(line[20] - (-3.07561)) < 0.05
    # symmetric problem: y and z centre lines agree with the exact line as well
    mid = 21
    for ln in (r["divQ"][mid, 1:42, mid], r["divQ"][1:42, mid, mid]):
        assert np.sqrt(np.mean((ln - exact) ** 2)) < 0.02


def test_steps_per_ray_reference_semantics(po, bc41):
    lv, p, r = bc41
    # default allowReflect=true marches a zero-weight second leg: 62.4 steps/ray at 41^3
    spr = r["steps"] / (41 ** 3 * 100)
    assert 61.5 < spr < 63.2, spr


def test_zero_weight_second_leg_is_exact_noop(po):
    lv = po.burns_christon(17)
    a = po.ray_trace(po.default_params(nDivQRays=20, threshold=0.05, isSeedRandom=0, allowReflect=1), lv)
    b = po.ray_trace(po.default_params(nDivQRays=20, threshold=0.05, isSeedRandom=0, allowReflect=0), lv)
    assert np.array_equal(a["divQ"], b["divQ"])
    assert a["steps"] > 1.8 * b["steps"]


def modest_exact(n, Tw=1000.0, Tm=1500.0, sigma=5.67051e-8, ntheta=100000):
    """benchMark2 of RMCRT.m:44-94 (k = 1, eps = 1) at the n cell centres between the plates"""
    x = np.linspace(1 / (2 * n), 1 - 1 / (2 * n), n)
    mu = np.cos(np.linspace(0, np.pi / 2, ntheta))
    dmu = mu[:-1] - mu[1:]
    E2 = np.zeros(n)
    with np.errstate(divide="ignore"):
        for ix in range(n):
            f = np.exp(-x[ix] / mu)
            f[~np.isfinite(f)] = 0
            E2[ix] = np.sum((f[:-1] + f[1:]) / 2 * dmu)
    D = np.zeros(n)
    for ix in range(1, n):
        D[ix - 1] = sigma * (Tw ** 4 - Tm ** 4) * -2 * (E2[ix - 1] + E2[n - ix])
    D[n - 1] = sigma * (Tw ** 4 - Tm ** 4) * -2 * (E2[n - 1] + 1.0)
    D[n - 1] = (D[n - 1] + D[n - 2]) / 2
    return D


def test_modest_parallel_plates(po):
    """RMCRT_1L_reflect.ups: 4 mirror walls, 2 black plates at 1000 K, medium 1500 K, kappa 1.
    Plane-averaged divQ follows Modest's closed form (RMCRT's divQ has the opposite sign convention)."""
    N = 41
    lv = po.parallel_plates(N, mirror_faces=("x-", "x+", "y-", "y+"))
    p = po.default_params(nDivQRays=16, threshold=0.001, isSeedRandom=1, allowReflect=1)
    r = po.ray_trace(p, lv, seed=3, patches=(1, 1, 8), threads=8)
    plane = r["divQ"][1:N + 1, 1:N + 1, 1:N + 1].mean(axis=(1, 2))
    D = modest_exact(N)
    rel = np.sqrt(np.mean((plane + D) ** 2)) / np.abs(D).mean()
    assert rel < 0.03, rel


def test_deterministic_mode_is_decomposition_invariant(po):
    """<randomSeed> false: per-ray reseed makes the result independent of the patch layout"""
    lv = po.burns_christon(12)
    p = po.default_params(nDivQRays=3, threshold=0.05, isSeedRandom=0)
    a = po.ray_trace(p, lv, patches=(1, 1, 1))
    b = po.ray_trace(p, lv, patches=(2, 3, 2), threads=4)
    assert np.array_equal(a["divQ"], b["divQ"]) and a["steps"] == b["steps"]


def test_boundary_flux_mask_and_slots(po):
    """boundFlux.p and the set of written Stencil7 slots (Ray.cc:664,716-717; has_a_boundary 1517-1572)"""
    rng = np.random.default_rng(5)
    from util import add_random_walls
    lv = po.burns_christon(10)
    add_random_walls(lv, 0.05, rng)
    p = po.default_params(nDivQRays=1, nFluxRays=2, threshold=0.05, isSeedRandom=0, solveBoundaryFlux=1, solveDivQ=0)
    r = po.ray_trace(p, lv)
    ct = lv.cellType
    flow = ct == -1
    S = r["boundFlux"]
    nb = {0: (0, 0, -1), 1: (0, 0, 1), 2: (0, -1, 0), 3: (0, 1, 0), 4: (-1, 0, 0), 5: (1, 0, 0)}  # slot -> (dz,dy,dx)
    any_face = np.zeros(ct.shape, bool)
    for slot, (dz, dy, dx) in nb.items():
        wall_nb = np.zeros(ct.shape, bool)
        wall_nb[1:-1, 1:-1, 1:-1] = ct[1 + dz:ct.shape[0] - 1 + dz, 1 + dy:ct.shape[1] - 1 + dy, 1 + dx:ct.shape[2] - 1 + dx] != -1
        expect = wall_nb & flow
        expect[[0, -1], :, :] = False; expect[:, [0, -1], :] = False; expect[:, :, [0, -1]] = False
        # gas at 64.8 K around cold walls: every traced face receives a strictly positive flux
        assert np.array_equal(S[..., slot] != 0, expect), slot
        any_face |= expect
    assert np.array_equal(S[..., 6] == 1.0, any_face)
    assert set(np.unique(S[..., 6])) <= {0.0, 1.0}


def test_map_cell_to_coarser_negative_indices(po):
    """Level::mapCellToCoarser quirk (Level.cc:1285-1307): -1 -> -1, but -2 -> -2 for ratio 4"""
    out = (C.c_int * 3)()
    po.lib().oracle_mapCellToCoarser((C.c_int * 3)(-1, 7, -4), (C.c_int * 3)(4, 4, 4), out)
    assert tuple(out) == (-1, 1, -1)
    po.lib().oracle_mapCellToCoarser((C.c_int * 3)(-2, 8, -5), (C.c_int * 3)(4, 4, 4), out)
    assert tuple(out) == (-2, 2, -2)


def test_data_onion_close_to_single_level(po):
    """2-level Data Onion (fine 32^3, coarse 8^3, halo 4) approximates the single-level answer"""
    fine = po.burns_christon(32)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    p = po.default_params(nDivQRays=24, threshold=0.05, isSeedRandom=1)
    one = po.ray_trace(p, fine, seed=11, patches=(1, 1, 8), threads=8)
    do = po.ray_trace_data_onion(p, [coarse, fine], (4, 4, 4), seed=11, patches=(2, 2, 2), threads=8)
    a = one["divQ"][1:-1, 1:-1, 1:-1]
    b = do["divQ"][1:-1, 1:-1, 1:-1]
    assert abs(a.mean() - b.mean()) / abs(a.mean()) < 0.02
    assert do["steps"] < one["steps"]          # coarse cells are 4x longer


def test_extents_patch_based(po):
    fine = po.burns_christon(32)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    roiLo, roiHi, regLo, regHi = po.compute_extents([coarse, fine], (16, 0, 0), (32, 16, 16), (4, 4, 4))
    assert roiLo == (12, -1, -1) and roiHi == (33, 20, 20)           # clamped to the level incl. extra cells
    assert tuple(regLo[1]) == (13, 0, 0) and tuple(regHi[1]) == (32, 19, 19)
    assert tuple(regLo[0]) == (0, 0, 0) and tuple(regHi[0]) == (8, 8, 8)


def test_randVector_is_the_reference_fisher_yates(po):
    """RMCRTCommon::randVector (RMCRTCommon.cc:998-1024) restated independently in python on the pinned MT19937 stream:
    seed(i+j+k), then for i = n-1..1 swap a[i] with a[randInt(i)] (randInt(n) = rejection on the masked word)"""
    def rand_int(stream, n):
        used = n
        for s in (1, 2, 4, 8, 16):
            used |= used >> s
        while True:
            v = int(next(stream)) & used
            if v <= n:
                return v
    p = po.default_params(isSeedRandom=0)
    for cell, n in (((1, 2, 3), 10), ((0, 0, 0), 7), ((20, 3, 40), 100), ((5, 5, 5), 600)):
        words = iter(po.mt_u32(sum(cell), 4 * n + 64))
        a = list(range(n))
        for i in range(n - 1, 0, -1):
            r = rand_int(words, i)
            a[i], a[r] = a[r], a[i]
        assert list(po.rand_vector(p, cell, n)) == a
    fixed = po.default_params(fixedRandom=1)
    a = list(range(10))
    for i in range(9, 0, -1):
        r = int(0.3 * i)
        a[i], a[r] = a[r], a[i]
    assert list(po.rand_vector(fixed, (1, 1, 1), 10)) == a


def test_latin_hyper_cube_lowers_the_variance(po):
    """LHC stratifies cos(theta) and phi (Ray.cc:1438-1460): same mean, smaller seed-to-seed spread than naive sampling"""
    lv = po.burns_christon(11)
    runs = {0: [], 1: []}
    for algo in (0, 1):
        for seed in range(6):
            p = po.default_params(nDivQRays=64, threshold=0.05, rayDirSampleAlgo=algo)
            runs[algo].append(po.ray_trace(p, lv, seed=100 + seed)["divQ"][1:-1, 1:-1, 1:-1])
    m0, m1 = np.mean(runs[0], 0), np.mean(runs[1], 0)
    s0, s1 = np.std(runs[0], 0).mean(), np.std(runs[1], 0).mean()
    assert abs(m0.mean() - m1.mean()) < 0.02 * abs(m0.mean())
    assert s1 < 0.9 * s0


def test_radiometer_isothermal_enclosure(po):
    """closed form: in an isothermal black enclosure every ray returns sigmaT4/pi, so
    VRIntensity = I * mean(cos theta) and VRFlux = VRIntensity * solidAngle (Radiometer.cc:483-493);
    for the uniform-in-cos(theta) cone sampling (530) mean(cos theta) = (1 + cos(viewAngle/2)) / 2"""
    lv = po.parallel_plates(12, T_gas=1000.0, T_wall=1000.0, mirror_faces=())
    I = float(lv.sigmaT4[5, 5, 5])
    p = po.default_params(threshold=0.0001, allowReflect=0)      # (with reflections the un-weighted wall term, SURVEY Q1, adds to I)
    for view, vec in ((30.0, (0, 0, 1)), (120.0, (1, 1, 0)), (60.0, (-1, -2, 0.5))):
        r = po.radiometer(view, vec, (5, 5, 5), (7, 6, 6), 4000, is_normal_vector=True)
        flux, inten, steps = po.radiometer_flux(p, lv, [r], seed=3)
        half = np.radians(view / 2)
        expect = I * (1 + np.cos(half)) / 2
        for c in ((6, 6, 6), (6, 6, 7)):          # [z][y][x] with lo = -1
            assert inten[c] == pytest.approx(expect, rel=0.01)
            assert flux[c] == pytest.approx(inten[c] * 2 * np.pi * (1 - np.cos(half)), rel=1e-12)
        assert np.count_nonzero(flux) == 2 and steps > 0


def test_radiometer_direction_cone(po):
    """every rayDirection_VR direction lies inside the view cone around the requested normal (Radiometer.cc:505-560).
    The rotation (theta about y, then xi about z) is exact for normals in quadrants 1, 2 and 4 of the x-y plane."""
    lv = po.parallel_plates(8, T_gas=1000.0, T_wall=1000.0, mirror_faces=())
    # a cold black wall on x+ only: the radiometer sees less when it looks at it
    lv.sigmaT4[:, :, -1] = 0.0
    p = po.default_params(threshold=0.0001, allowReflect=0)
    I = float(lv.sigmaT4[4, 4, 4])
    towards = po.radiometer(20.0, (1, 0, 0), (4, 4, 4), (5, 5, 5), 3000)
    away = po.radiometer(20.0, (-1, 0, 0), (4, 4, 4), (5, 5, 5), 3000)
    f1, i1, _ = po.radiometer_flux(p, lv, [towards])
    f2, i2, _ = po.radiometer_flux(p, lv, [away])
    # looking at the cold wall: only the gas column (kappa 1, ~0.44 m) emits; looking away: full black-body intensity
    assert i1[5, 5, 5] < 0.5 * i2[5, 5, 5]
    assert i2[5, 5, 5] == pytest.approx(I * (1 + np.cos(np.radians(10))) / 2, rel=0.01)


# ---------------------------------------------------------------- the reference's own exact-solution tables (postProcessTools)
@pytest.mark.parametrize("bm,N,axis,rays,tol", [
    (3, 27, 2, 2000, 0.006),     # bench3Exact.txt sub-sampled exactly like RMCRT.m:113-131 (Exact27)
    (3, 81, 2, 500, 0.006),      # Exact81
    (7, 41, 0, 1000, 0.03),      # kexpDivQx.dat (pchip to 41 cells)
    (7, 41, 1, 1000, 0.03),
    (7, 41, 2, 1000, 0.03),
    (8, 41, 0, 1000, 0.06),      # cloudsDivQx.dat
    (8, 41, 1, 1000, 0.03),      # cloudsDivQy.dat
])
def test_oracle_matches_reference_exact_tables(po, bm, N, axis, rays, tol):
    """Known answers the reference keeps for this path besides Burns & Christon: the oracle's centre-line divQ of
    RMCRT_Test benchmarks 3, 7 and 8 against tests/golden/rmcrt_benchmarks.json (made from bench3Exact.txt, kexpDivQ*.dat,
    cloudsDivQ*.dat).  Measured here: 0.0016-0.0023 (bm 3), 0.010-0.015 (bm 7), 0.010 / 0.038 (bm 8 y / x at 41 cells)."""
    lv = po.benchmark_level(bm, N, pDir=0)
    lo, hi = po.centreline_box(N, axis)
    r = po.ray_trace(po.default_params(nDivQRays=rays, threshold=0.001, isSeedRandom=1), lv, patch_lo=lo, patch_hi=hi, seed=5)
    line = po.centreline_of(r["divQ"], N, axis)
    ex = po.benchmark_exact(bm, N, axis)
    rel = np.linalg.norm(line - ex) / np.linalg.norm(ex)
    assert rel < tol, rel
    # only the centre line was traced
    assert np.count_nonzero(r["divQ"]) <= N


def test_benchmark_golden_fixture_shape(po):
    g = po.golden_benchmarks()
    assert g["bench3"]["n"] == 243 and len(g["bench3"]["values"]) == 243
    assert sorted(g["kexp"].keys()) == ["scale_to_code_problem", "x", "y", "z"]
    assert len(g["clouds"]["x"]["x"]) == len(g["clouds"]["x"]["divQ"]) == 247
    # the 243-point table is symmetric about the mid plane to its own Monte-Carlo noise (it is itself a computed solution)
    v = np.array(g["bench3"]["values"])
    assert np.linalg.norm(v - v[::-1]) / np.linalg.norm(v) < 5e-3


# ---------------------------------------------------------------- RMCRT_coarseLevel: Ray::refine_Q (selectInterpolator, AMR.h)
def _refine_case(rr, n_coarse=6, seed=3):
    """one coarse level of n_coarse^3 interior cells (+1 extra layer), fine patch = a sub-box of the rr-times finer level"""
    rng = np.random.default_rng(seed)
    glo, ghi = (-1, -1, -1), (n_coarse + 1,) * 3
    q = rng.normal(size=(n_coarse + 2,) * 3)
    fl, fh = (rr[0] * 1, rr[1] * 0, rr[2] * 2), (rr[0] * 4, rr[1] * 3, rr[2] * n_coarse)      # touches y-, z+ domain faces
    olo, ohi = (fl[0], fl[1] - 1, fl[2]), (fh[0], fh[1], fh[2] + 1)                           # patch window incl. its extra cells
    return q, glo, ghi, fl, fh, olo, ohi


@pytest.mark.parametrize("rr", [(2, 2, 2), (4, 4, 4), (2, 4, 1)])
def test_refine_Q_reproduces_polynomials(po, rr):
    """order 0 copies the parent, order 1 is exact for tri-linear fields, order 2 for tri-quadratic ones (cell-centred)"""
    q, glo, ghi, fl, fh, olo, ohi = _refine_case(rr)
    n = q.shape[0] - 2
    cc = np.arange(-1, n + 1) + 0.5                                    # coarse cell centres in coarse-cell units
    Z, Y, X = np.meshgrid(cc, cc, cc, indexing="ij")
    fx = [(np.arange(olo[d], ohi[d]) + 0.5) / rr[d] for d in range(3)]
    FZ, FY, FX = np.meshgrid(fx[2], fx[1], fx[0], indexing="ij")
    sl = tuple(slice(fl[d] - olo[d], fh[d] - olo[d]) for d in (2, 1, 0))
    lin = lambda x, y, z: 1.5 + 0.25 * x - 0.5 * y + 2 * z + 0.125 * x * y * z
    quad = lambda x, y, z: lin(x, y, z) + 0.5 * x * x - 0.25 * y * y + z * z * x
    for order, fn, tol in ((1, lin, 1e-13), (2, quad, 1e-12)):
        out = po.refine_Q(fn(X, Y, Z), glo, ghi, glo, ghi, rr, order, fl, fh, olo, ohi)
        np.testing.assert_allclose(out[sl], fn(FX, FY, FZ)[sl], rtol=tol, atol=tol)
        mask = np.ones(out.shape, bool); mask[sl] = False
        assert not out[mask].any()                                         # only [fl,fh) is written (Ray.cc:1899-1934)
    out0 = po.refine_Q(q, glo, ghi, glo, ghi, rr, 0, fl, fh, olo, ohi)
    k, j, i = np.meshgrid(*[np.arange(fl[d], fh[d]) // rr[d] + 1 for d in (2, 1, 0)], indexing="ij")
    assert np.array_equal(out0[sl], q[k, j, i])


def test_arches_property_tasks_known_answers(po):
    """RMCRT_Radiation::sigmaT4 / sumAbsk / stencilToDBLs restatements on hand-checkable inputs"""
    T = np.array([[[1000.0, 500.0, 300.0]]]); Tp = np.array([[[2000.0, 500.0, 900.0]]])
    ak = np.array([[[1.0, 0.0, 0.3]]]); akp = np.array([[[3.0, 0.0, 0.1]]])
    gv = np.array([[[1.0, 1.0, 0.0]]])
    s = po.arches_sigmaT4(gv, [ak, akp], [T, Tp], True)
    sop = po.SIGMA / np.pi
    np.testing.assert_allclose(s[0, 0], [sop * (1e12 * 1 + 16e12 * 3) / 4, 0.0, sop * 300.0 ** 4], rtol=1e-15)
    np.testing.assert_allclose(po.arches_sigmaT4(gv, None, [T], False)[0, 0], sop * T[0, 0] ** 4, rtol=1e-15)
    # sumAbsk: 3x3x3 window with a 1-cell extra layer around one interior cell
    gvw = np.ones((3, 3, 3)); a1 = np.full((3, 3, 3), 0.25); a2 = np.full((3, 3, 3), 0.5)
    out = po.arches_sumAbsk(gvw, [a1, a2], (-1,) * 3, (2,) * 3, (0,) * 3, (1,) * 3)
    assert out[1, 1, 1] == 0.75 and (np.delete(out.ravel(), 13) == 1.0).all()
    gvw[1, 1, 1] = 0.0
    assert po.arches_sumAbsk(gvw, [a1, a2], (-1,) * 3, (2,) * 3, (0,) * 3, (1,) * 3)[1, 1, 1] == 1.0
    bf = np.arange(14.0).reshape(2, 7)
    d = po.arches_stencilToDBLs(bf)
    assert [d[k][0] for k in "EWNSTB"] == [1, 0, 3, 2, 5, 4] and d["T"][1] == 12
