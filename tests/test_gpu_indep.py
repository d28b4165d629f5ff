"""divq_indep_kernel (indep.cuh, rmcrt_params.blockedMarch = 1): the alternative kernel of the default sampling (rayCoherence 0) --
blocked half-packed records, one 256-bit load per 2 x 2 x 2 block, per-lane ray refill -- against the production one-cell-per-lane
kernel (blockedMarch = 0) on the SAME Philox rays (both key direction and origin on (cell, ray, timestep)), and against the oracle."""
import numpy as np
import pytest

from util import add_random_walls

pytestmark = pytest.mark.gpu


def make(**kw):
    from uintah_b200 import RMCRT
    return RMCRT(device=0, **kw)


def solve(lv, patch=None, **kw):
    g = make(**kw)
    g.upload_level(0, lv)
    plo, phi = patch if patch else (None, None)
    out = g.trace_divq(0, patch_lo=plo, patch_hi=phi, timestep=6)
    c = g.counters()
    g.close()
    return out["divQ"], out["radiationVolq"], int(c.nRaySteps), int(c.nDivQCells)


@pytest.mark.parametrize("bm", [1, 2, 3])
@pytest.mark.parametrize("case", ["bc_walls_odd_patch", "plates_mirrors", "bc_refsem"])
def test_blocked_march_matches_one_cell_kernel(gpu_lib, po, case, bm):
    """same rays, same fp16 properties (half-packed level records for the one-cell kernel): divQ to FP32 round-off, executed steps and
    written cells equal (ties of the axis choice aside)"""
    rng = np.random.default_rng(4)
    kw = dict(nDivQRays=24, threshold=0.05, seed=13, half_records=True)
    patch = None
    if case == "bc_walls_odd_patch":
        lv = po.burns_christon(37, dtype=np.float32)
        add_random_walls(lv, 0.03, rng, abskg=0.8, sigmaT4=0.25)
        patch = ((3, 5, 2), (36, 30, 11))                  # partial 32 x 2 x 2 blocks on every side, odd offsets
    elif case == "plates_mirrors":                         # several reflections per ray, hot black walls
        lv = po.parallel_plates(24, dtype=np.float32, mirror_faces=("x-", "y-"))
        kw["threshold"] = 0.001
    else:                                                  # the reference's zero-weight second legs are marched
        lv = po.burns_christon(24, dtype=np.float32)
        kw["skipZeroWeightLegs"] = 0
    d1, v1, n1, c1 = solve(lv, patch, blockedMarch=0, **kw)
    d2, v2, n2, c2 = solve(lv, patch, blockedMarch=bm, **kw)
    assert c1 == c2 and c1 > 0
    if bm == 2:      # the same PTX leg loop on the same fp16 words: ray sums and step counts bit for bit
        assert n1 == n2 and np.array_equal(v1, v2)
    if bm == 3:      # the same rays again, added per cell in ray order instead of per warp: FP64 round-off
        assert n1 == n2 and np.allclose(v1, v2, rtol=1e-12, atol=0)
    assert np.array_equal(d1 == 0, d2 == 0) and np.array_equal(v1 == 0, v2 == 0)
    assert abs(n1 - n2) <= 3e-4 * n1, (n1, n2)
    # the ray sums (radiationVolq = 4 pi sumI / N) agree to FP32 round-off; divQ additionally differs by the fp16 rounding of the origin
    # cell's own sigmaT4/pi, which the blocked kernel's epilogue uses on purpose (<= 2^-11 relative on the emission term)
    scale = np.abs(v1).max()
    diff = np.abs(v1 - v2)
    assert np.percentile(diff, 99.9) <= 1e-4 * scale and diff.max() <= 2e-2 * scale, (np.percentile(diff, 99.9) / scale, diff.max() / scale)
    assert np.linalg.norm(v1 - v2) / np.linalg.norm(v1) < 3e-4
    assert np.abs(d1 - d2).max() <= 2e-2 * np.abs(d1).max() and np.linalg.norm(d1 - d2) / np.linalg.norm(d1) < 1.5e-3


@pytest.mark.parametrize("bm", [1, 2, 3])
def test_blocked_march_on_float_records(gpu_lib, po, bm):
    """float (8-B) level records: the blocked copy is their fp16 rounding -- divQ within the fp16 rounding of the properties along the
    rays of the one-cell kernel marching the float records, same masks"""
    rng = np.random.default_rng(5)
    lv = po.burns_christon(32, dtype=np.float32)
    add_random_walls(lv, 0.01, rng)
    kw = dict(nDivQRays=32, threshold=0.05, seed=21)
    d1, _, n1, c1 = solve(lv, None, blockedMarch=0, **kw)
    d2, _, n2, c2 = solve(lv, None, blockedMarch=bm, **kw)
    assert c1 == c2 and np.array_equal(d1 == 0, d2 == 0)
    assert abs(n1 - n2) <= 5e-4 * n1
    scale = np.abs(d1).max()
    diff = np.abs(d1 - d2)[d1 != 0]
    assert np.percentile(diff, 99.5) < 1.5e-3 * scale and np.linalg.norm(d1 - d2) / np.linalg.norm(d1) < 1e-3


@pytest.mark.parametrize("bm", [1, 2, 3])
def test_isothermal_enclosure_is_exact(gpu_lib, po, bm):
    """sigmaT4/pi = 0.6 is NOT exact in fp16: the epilogue uses the same fp16 value the rays integrate, so divQ == 0 to FP32 round-off"""
    rng = np.random.default_rng(6)
    M = 27
    ab = rng.uniform(0.1, 2.0, (M, M, M)).astype(np.float32); s4 = np.full((M, M, M), 0.6, np.float32); ct = np.full((M, M, M), -1, np.int32)
    for ax in range(3):
        for idx in (0, M - 1):
            sl = [slice(None)] * 3; sl[ax] = idx; ct[tuple(sl)] = 8; ab[tuple(sl)] = 1.0
    walls = (rng.uniform(size=(M, M, M)) < 0.03) & (ct == -1); ct[walls] = 8; ab[walls] = 1.0
    iso = po.Level(ab, s4, ct, (-1,) * 3, (M - 1,) * 3, (1.0 / 25,) * 3, interior_lo=(0,) * 3, interior_hi=(25,) * 3)
    for half in (False, True):
        d, _, _, c = solve(iso, None, nDivQRays=8, threshold=0.001, seed=5, allowReflect=0, half_records=half, blockedMarch=bm)
        assert np.abs(d).max() < 2e-5 * 4 * np.pi * 2.0 * 0.6
        assert c == int((ct == -1).sum())


def test_refill_threshold_does_not_change_results(gpu_lib, po):
    """the rays of a cell and their order are fixed by (cell, ray index): any refill batching gives bit-identical divQ"""
    lv = po.burns_christon(20, dtype=np.float32)
    ref = None
    for rl in (1, 8, 32):
        d, _, n, _ = solve(lv, None, nDivQRays=10, threshold=0.05, seed=2, refillLanes=rl, blockedMarch=1)
        if ref is None:
            ref = (d, n)
        assert np.array_equal(d, ref[0]) and n == ref[1]


@pytest.mark.parametrize("bm", [1, 2, 3])
def test_monte_carlo_tolerance_41(gpu_lib, po, bm):
    """BASELINE config 1 through the blocked kernel: SURVEY 8c (i)-(iii) against 8 oracle seeds, unrelaxed"""
    from util import candidate_stats, loo_control
    lv = po.burns_christon(41)
    p = po.default_params(nDivQRays=100, threshold=0.05, isSeedRandom=1)
    runs = [po.ray_trace(p, lv, seed=100 * s, patches=(1, 1, 8), threads=8)["divQ"] for s in range(1, 9)]
    g = make(nDivQRays=100, threshold=0.05, seed=2024, blockedMarch=bm)
    g.upload_level(0, lv)
    d = g.trace_divq(0)["divQ"]
    inner = (slice(1, -1),) * 3
    c3, c5, cl2 = loo_control(runs, inner)
    f3, f5, l2 = candidate_stats(d, runs, inner)
    assert f3 >= min(c3) and f5 >= min(c5) and l2 <= max(cl2), (f3, c3, f5, c5, l2, cl2)
    L2_ref = [po.centreline_L2(r, 41)[0] for r in runs]
    assert po.centreline_L2(d, 41)[0] <= max(L2_ref), (po.centreline_L2(d, 41)[0], L2_ref)


@pytest.mark.parametrize("halo", [2, 6])
def test_origin_brick_staging_is_bit_identical(gpu_lib, po, halo):
    """divq_brick_kernel (originBrickHalo, A/B knob): records inside the staged neighbourhood of the origin tile come from shared memory,
    the rest from global memory -- same rays, same arithmetic: divQ and step counts equal the generic march bit for bit"""
    rng = np.random.default_rng(8)
    lv = po.burns_christon(32, dtype=np.float32)
    add_random_walls(lv, 0.02, rng, abskg=0.7, sigmaT4=0.3)
    kw = dict(nDivQRays=12, threshold=0.05, seed=4, genericMarch=1)
    d0, v0, n0, c0 = solve(lv, None, **kw)
    d1, v1, n1, c1 = solve(lv, None, originBrickHalo=halo, **kw)
    assert n0 == n1 and c0 == c1
    assert np.array_equal(d0, d1) and np.array_equal(v0, v1)
    # mirrors: reflected rays re-enter the brick
    pl = po.parallel_plates(32, dtype=np.float32, mirror_faces=("x-", "y-"))
    kw["threshold"] = 0.001
    a = solve(pl, None, **kw)
    b = solve(pl, None, originBrickHalo=halo, **kw)
    assert a[2] == b[2] and np.array_equal(a[0], b[0])


def test_sorted_kernel_is_independent_of_the_launch_shape(gpu_lib, po):
    """divq_sorted_kernel (blockedMarch = 3) adds the rays of a cell in ray order whatever the sort, the chunk hand-out and the pass
    length did: bit-identical divQ for any rays-per-pass (refillLanes doubles as that A/B knob), any patch decomposition, run to run"""
    rng = np.random.default_rng(9)
    lv = po.burns_christon(30, dtype=np.float32)
    add_random_walls(lv, 0.02, rng, abskg=0.7, sigmaT4=0.3)
    kw = dict(nDivQRays=37, threshold=0.05, seed=4, blockedMarch=3)
    d0, v0, n0, c0 = solve(lv, None, **kw)
    for rp in (5, 16, 37):
        d1, v1, n1, c1 = solve(lv, None, refillLanes=rp, **kw)
        assert n0 == n1 and c0 == c1 and np.array_equal(d0, d1) and np.array_equal(v0, v1), rp
    # two patches (the 4 x 4 x 2-cell tiles start elsewhere): every cell keeps its rays and their order
    g = make(**kw)
    g.upload_level(0, lv)
    a = g.trace_divq(0, patch_lo=(0, 0, 0), patch_hi=(30, 30, 13), timestep=6)["divQ"]
    b = g.trace_divq(0, patch_lo=(0, 0, 13), patch_hi=(30, 30, 30), timestep=6)["divQ"]
    g.close()
    assert np.array_equal(a + b, d0)


def test_sorted_kernel_many_rays_per_pass(gpu_lib, po):
    """ray ids of a pass are 16 bits (32 cells x at most 2048 rays): a larger refillLanes is clamped, not wrapped -- thousands of rays per
    cell give bit-identical divQ for any pass length"""
    lv = po.burns_christon(10, dtype=np.float32)
    kw = dict(nDivQRays=2500, threshold=0.05, seed=8, blockedMarch=3)
    d0, v0, n0, c0 = solve(lv, None, **kw)
    for rp in (2048, 2500, 100000):
        d1, v1, n1, c1 = solve(lv, None, refillLanes=rp, **kw)
        assert n0 == n1 and c0 == c1 and np.array_equal(d0, d1) and np.array_equal(v0, v1), rp


@pytest.mark.parametrize("case", ["bc", "bc_walls", "plates_mirrors"])
def test_sorted_data_onion_matches_roi_copy_path(gpu_lib, po, case):
    """2-level Data Onion, library default (divq_sorted_kernel<ONION>: blocked copies of the patch's ROI and of level 0, march_onion2_blk)
    against the linear ROI-copy path (blockedMarch 0, march_onion2_fast) on half-packed levels: the same Philox rays through the same fp16
    words with the same leg arithmetic -- executed steps equal, ray sums equal up to the order they are added in (FP64 round-off)"""
    rng = np.random.default_rng(31)
    kw = dict(nDivQRays=20, haloCells=(4, 4, 4), seed=17, half_records=True)
    if case == "plates_mirrors":
        fine, kw["threshold"] = po.parallel_plates(32, dtype=np.float32, mirror_faces=("x-", "y-")), 0.001
    else:
        fine, kw["threshold"] = po.burns_christon(32, dtype=np.float32), 0.05
        if case == "bc_walls":
            add_random_walls(fine, 0.02, rng, abskg=0.7, sigmaT4=0.3)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    res = {}
    for bm in (0, 3):
        g = make(blockedMarch=bm, **kw)
        g.upload_level(0, coarse)
        g.upload_level(1, fine)
        divQ = np.zeros(fine.abskg.shape); vol = np.zeros_like(divQ)
        # (halo 4 on patches of 16 x 16 x 32: the ROI of the inner patch faces ends exactly at the interior's edge on some sides)
        for plo, phi in po._split((0, 0, 0), (32, 32, 32), (2, 2, 1)) + [((4, 4, 4), (12, 12, 12))]:
            g.trace_dataonion(2, plo, phi, zero_init=False, outputs=(divQ, vol, None), timestep=4)
        c = g.counters()
        res[bm] = (divQ, vol, int(c.nRaySteps), int(c.nDivQCells))
        g.close()
    (d0, v0, n0, c0), (d3, v3, n3, c3) = res[0], res[3]
    assert c0 == c3 and c0 > 0 and n0 == n3
    assert np.array_equal(v0 == 0, v3 == 0)
    assert np.allclose(v0, v3, rtol=1e-12, atol=0)
    # divQ: the default kernel's epilogue takes the origin cell's sigmaT4/pi from the fp16 word its rays integrate
    assert np.abs(d0 - d3).max() <= 2e-3 * np.abs(d0).max()


@pytest.mark.parametrize("case", ["bc", "bc_refsem", "plates_mirrors"])
def test_sorted_data_onion_three_levels(gpu_lib, po, case):
    """3-level Data Onion (fine 48^3, RR 2, 2): the default kernel marches blocked boxes of the fine ROI and of level 1's region and the
    blocked copy of level 0 (march_onion_blk: two level jumps per ray) -- against the generic multi-level march (blockedMarch 0) on the same
    Philox rays and the same half-packed records: same cells written, executed steps equal up to FP32 ties, divQ to FP32 round-off.
    Patches at the domain boundary, in the interior and with an ROI that ends exactly at the interior's edge."""
    rng = np.random.default_rng(33)
    kw = dict(nDivQRays=16, haloCells=(2, 2, 2), seed=23, half_records=True)
    if case == "plates_mirrors":
        fine, kw["threshold"] = po.parallel_plates(48, dtype=np.float32, mirror_faces=("x-", "y-")), 0.001
    else:
        fine, kw["threshold"] = po.burns_christon(48, dtype=np.float32), 0.05
        if case == "bc_refsem":      # march the zero-weight legs behind the black walls: which level a ray reflects on shows in the steps
            kw["skipZeroWeightLegs"] = 0
    # (no reflective intrusions here: a level jump that lands in a coarse wall cell makes the reference reflect into an index of the finer
    # level -- SURVEY Q7, undefined behaviour that the generic march, the oracle and this march resolve differently)
    fine.rr = (2, 2, 2)
    mid = po.coarsen_level(fine, (2, 2, 2)); mid.rr = (2, 2, 2)
    coarse = po.coarsen_level(mid, (2, 2, 2))
    patches = [((0, 0, 0), (16, 16, 16)), ((16, 16, 16), (32, 32, 32)), ((2, 32, 20), (18, 48, 36))]
    res = {}
    for bm in (0, 3):
        g = make(blockedMarch=bm, **kw)
        g.upload_level(0, coarse); g.upload_level(1, mid); g.upload_level(2, fine)
        divQ = np.zeros(fine.abskg.shape); vol = np.zeros_like(divQ)
        for plo, phi in patches:
            g.trace_dataonion(3, plo, phi, zero_init=False, outputs=(divQ, vol, None), timestep=4)
        c = g.counters()
        res[bm] = (divQ, vol, int(c.nRaySteps), int(c.nDivQCells))
        g.close()
    (d0, v0, n0, c0), (d3, v3, n3, c3) = res[0], res[3]
    assert c0 == c3 and c0 > 0
    assert abs(n0 - n3) <= 3e-4 * n0, (n0, n3)
    assert np.array_equal(v0 == 0, v3 == 0)
    scale = np.abs(v0).max()
    diff = np.abs(v0 - v3)
    assert np.percentile(diff, 99.9) <= 2e-4 * scale and diff.max() <= 2e-2 * scale, (np.percentile(diff, 99.9) / scale, diff.max() / scale)
    assert np.abs(d0 - d3).max() <= 2e-2 * np.abs(d0).max() and np.linalg.norm(d0 - d3) / np.linalg.norm(d0) < 2e-3


@pytest.mark.parametrize("rays,ccrays", [(1, 0), (3, 0), (7, 1), (300, 0)])
def test_sorted_kernel_odd_shapes(gpu_lib, po, rays, ccrays):
    """fewer rays than warps, more rays than one pass sorts (256), cell-centred rays, a level of 5 x 4 x 3 cells and a patch smaller than a
    tile: the default kernel against the one-cell kernel on the same rays (half-packed level)"""
    rng = np.random.default_rng(51)
    for shape_case in ("tiny_level", "small_patch"):
        if shape_case == "tiny_level":
            M = (3 + 2, 4 + 2, 5 + 2)            # (z, y, x) incl. extra cells
            ab = rng.uniform(0.2, 1.5, M).astype(np.float32); s4 = rng.uniform(0.1, 1.0, M).astype(np.float32); ct = np.full(M, -1, np.int32)
            for ax in range(3):
                for idx in (0, M[ax] - 1):
                    sl = [slice(None)] * 3; sl[ax] = idx; ct[tuple(sl)] = 8; ab[tuple(sl)] = 1.0
            lv = po.Level(ab, s4, ct, (-1,) * 3, (5 + 1, 4 + 1, 3 + 1), (0.2, 0.25, 1.0 / 3), interior_lo=(0,) * 3, interior_hi=(5, 4, 3))
            patch = None
        else:
            lv = po.burns_christon(21, dtype=np.float32)
            add_random_walls(lv, 0.05, rng, abskg=0.6, sigmaT4=0.2)
            patch = ((7, 9, 11), (12, 11, 13))
        kw = dict(nDivQRays=rays, threshold=0.01, seed=3, half_records=True, CCRays=ccrays)
        d0, v0, n0, c0 = solve(lv, patch, blockedMarch=0, **kw)
        d3, v3, n3, c3 = solve(lv, patch, blockedMarch=3, **kw)
        assert c0 == c3 and c0 > 0 and n0 == n3, (shape_case, c0, c3, n0, n3)
        assert np.array_equal(v0 == 0, v3 == 0) and np.allclose(v0, v3, rtol=1e-12, atol=0), shape_case
