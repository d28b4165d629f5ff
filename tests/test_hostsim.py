"""The production march / RNG headers (uintah_b200/csrc/march.cuh, rng.cuh) compiled for the HOST and stepped
ray-by-ray against the oracle.  Catches logic errors in the device code on a machine without a GPU; the same
comparisons run on the real kernels in test_gpu_parity.py."""
import numpy as np
import pytest

from hostsim import sim
from util import random_rays


def _compare_1L(po, lv, thr, refl, n=200, seed=0):
    rng = np.random.default_rng(seed)
    cells, dirs, org = random_rays(lv, n, rng)
    p = po.default_params(threshold=thr, allowReflect=refl)
    ref = [po.update_sumI(p, lv, dirs[t], org[t], cells[t], max_path=4000) for t in range(n)]
    s64, n64, p64 = sim.trace_1L(lv, 1, cells, dirs, org, thr, refl)
    s32, n32, _ = sim.trace_1L(lv, 0, cells, dirs, org, thr, refl)
    sfa, nfa = sim.trace_1L_fast(lv, cells, dirs, org, thr, refl)      # the production FP32 kernel's march
    assert (nfa == n32).mean() > 0.98 and (nfa == np.array([r[1] for r in ref])).mean() > 0.98
    for t, (s, ns, path) in enumerate(ref):
        assert s == pytest.approx(sfa[t], rel=5e-4, abs=1e-6 * max(1.0, abs(s)))
        assert ns == n64[t]
        assert np.array_equal(path, p64[t, :ns, :3])
        assert s == pytest.approx(s64[t], rel=1e-13, abs=1e-300)
        assert s == pytest.approx(s32[t], rel=5e-4, abs=1e-6 * max(1.0, abs(s)))
    assert (np.array([r[1] for r in ref]) == n32).mean() > 0.98   # FP32 ties may differ on a few rays


def test_single_level_burns_christon(po):
    _compare_1L(po, po.burns_christon(21), 0.05, 1)
    _compare_1L(po, po.burns_christon(21), 0.05, 0, seed=1)


def test_single_level_reflections(po):
    # mirrors + hot black walls: exercises reflect(), fs, the un-weighted wall term
    _compare_1L(po, po.parallel_plates(16), 0.001, 1, seed=2)
    _compare_1L(po, po.parallel_plates(12, mirror_faces=("x-", "x+", "y-", "y+")), 0.001, 1, seed=3)


def test_single_level_float_properties(po):
    _compare_1L(po, po.burns_christon(15, dtype=np.float32), 0.05, 1, seed=4)


@pytest.mark.parametrize("case", ["bc", "plates"])
def test_multi_level(po, case):
    rng = np.random.default_rng(7)
    fine = po.burns_christon(32) if case == "bc" else po.parallel_plates(32)
    thr = 0.05 if case == "bc" else 0.001
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    levels = [coarse, fine]
    p = po.default_params(threshold=thr, allowReflect=1)
    for plo, phi in (((0, 0, 0), (16, 16, 16)), ((16, 16, 0), (32, 32, 16)), ((8, 8, 8), (24, 24, 24))):
        roiLo, roiHi, regLo, regHi = po.compute_extents(levels, plo, phi, (4, 4, 4))
        cells, dirs, org = random_rays(fine, 120, rng, plo, phi)
        s64, n64, p64 = sim.trace_ML(levels, 1, (roiLo, roiHi), (regLo, regHi), cells, dirs, org, thr, 1)
        s32, n32, _ = sim.trace_ML(levels, 0, (roiLo, roiHi), (regLo, regHi), cells, dirs, org, thr, 1)
        sfa, nfa = sim.trace_DO2_fast(levels, (roiLo, roiHi), (regLo, regHi), cells, dirs, org, thr, 1)   # production DO march
        assert (nfa == n32).mean() >= 0.97
        same32 = 0
        for t in range(len(cells)):
            s, ns, path = po.update_sumI_ML(p, levels, roiLo, roiHi, regLo, regHi, dirs[t], org[t], cells[t], max_path=4000)
            assert ns == n64[t]
            assert np.array_equal(path, p64[t, :ns])          # (i,j,k,level) of every cell entered
            assert s == pytest.approx(s64[t], rel=1e-11)
            assert s == pytest.approx(s32[t], rel=1e-3, abs=1e-6 * max(1.0, abs(s)))
            if nfa[t] == ns:
                assert s == pytest.approx(sfa[t], rel=1e-3, abs=1e-6 * max(1.0, abs(s)))
            same32 += ns == n32[t]
        assert same32 >= 0.97 * len(cells)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10"""
    assert sim.philox((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert sim.philox((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert sim.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_mt_first_outputs_match_full_generator(po):
    """the stateless 'first outputs after seed()' used on the GPU == MTRand::seed(s); randInt() x n"""
    for seed in (1, 2, 3, 1234567, 4294967295, (5 + 6 + 7) * 99 + 1):
        full = [int(x) for x in po.mt_u32(seed, 5)]
        assert sim.mt_first(seed, 5) == full
        assert sim.mt_first(seed, 4) == full[:4]


@pytest.mark.parametrize("max_len", [0.05, 0.2, 0.6])
def test_single_level_bounded_ray_length(po, max_len):
    """<ROI_extents type="boundedRayLength">: inner and outer loops also stop on rayLength >= d_maxRayLength
    (RMCRTCommon.cc:680,684) and the cell the ray stops in is added like a wall (846-854)"""
    lv = po.parallel_plates(16)
    rng = np.random.default_rng(11)
    cells, dirs, org = random_rays(lv, 150, rng)
    p = po.default_params(threshold=0.001, allowReflect=1, maxRayLength=max_len)
    s64, n64, p64 = sim.trace_1L(lv, 1, cells, dirs, org, 0.001, 1, max_ray_length=max_len)
    s32, n32, _ = sim.trace_1L(lv, 0, cells, dirs, org, 0.001, 1, max_ray_length=max_len)
    unbounded = sim.trace_1L(lv, 1, cells, dirs, org, 0.001, 1)[1]
    assert (n64 < unbounded).any()            # the bound really cut rays short
    for t in range(len(cells)):
        s, ns, path = po.update_sumI(p, lv, dirs[t], org[t], cells[t], max_path=4000)
        assert ns == n64[t]
        assert np.array_equal(path, p64[t, :ns, :3])
        assert s == pytest.approx(s64[t], rel=1e-13, abs=1e-300)
        if ns == n32[t]:
            assert s == pytest.approx(s32[t], rel=5e-4, abs=1e-6 * max(1.0, abs(s)))
    assert (n32 == n64).mean() > 0.97


def test_keyed_permutation_is_a_bijection():
    """Latin-Hyper-Cube strata with <randomSeed> true: every n, every key must give a permutation of [0,n)"""
    rng = np.random.default_rng(3)
    seen = set()
    for n in (1, 2, 3, 5, 10, 16, 17, 100, 127, 128, 1000, 4097):
        for _ in range(4):
            key = [int(x) for x in rng.integers(0, 2 ** 32, 4)]
            perm = sim.keyed_permutation(n, key)
            assert np.array_equal(np.sort(perm), np.arange(n))
            if n == 100:
                seen.add(tuple(perm))
    assert len(seen) == 4                     # different keys, different permutations


@pytest.mark.parametrize("case", ["bc", "plates_float", "bounded"])
def test_scattering_march_is_the_oracle_bit_for_bit(po, case):
    """march_single_level_scatter (the device code of <sigmaScat> > 0) compiled for the host, where it shares glibc's log / sincos with
    the oracle: FP64 instantiation == oracle (== the reference's RAY_SCATTER build, test_oracle_ref.py) BIT FOR BIT, the FP32
    instantiation to round-off"""
    rng = np.random.default_rng(21)
    kw = dict(threshold=0.02, isSeedRandom=0, sigmaScat=4.0)
    if case == "bc":
        lv = po.burns_christon(17)
    elif case == "plates_float":
        lv = po.parallel_plates(14, dtype=np.float32, mirror_faces=("x-", "z+"))
        lv.abskg[lv.cellType != -1] = rng.uniform(0, 1.2, int((lv.cellType != -1).sum())).astype(np.float32)
        kw.update(threshold=0.005, sigmaScat=9.0)
    else:
        lv = po.burns_christon(17)
        kw.update(maxRayLength=0.4, sigmaScat=12.0)
    n = 300
    cells, dirs, org = random_rays(lv, n, rng)
    seeds = rng.integers(1, 2 ** 31, n).astype(np.uint32)
    p = po.default_params(**kw)
    ref = [po.update_sumI_seeded(p, lv, dirs[t], org[t], cells[t], int(seeds[t])) for t in range(n)]
    s64, n64 = sim.trace_1L_scatter(lv, 1, cells, dirs, org, seeds, kw["threshold"], 1, kw["sigmaScat"], kw.get("maxRayLength", 0.0))
    assert np.array_equal(n64, np.array([r[1] for r in ref], np.uint32))
    assert np.array_equal(s64, np.array([r[0] for r in ref]))
    s32, n32 = sim.trace_1L_scatter(lv, 0, cells, dirs, org, seeds, kw["threshold"], 1, kw["sigmaScat"], kw.get("maxRayLength", 0.0))
    same = n32 == n64
    assert same.mean() > 0.9
    np.testing.assert_allclose(s32[same], s64[same], rtol=2e-3, atol=1e-6 * np.abs(s64).max())
    p0 = po.default_params(**{**kw, "sigmaScat": 0.0})
    assert sum(po.update_sumI(p0, lv, dirs[t], org[t], cells[t])[1] != ref[t][1] for t in range(n)) > 100


@pytest.mark.parametrize("case", ["bc_walls", "plates_mirrors", "bc_refsem"])
def test_two_cells_per_lane_march_equals_two_single_marches(po, case):
    """march_single_level_dual (one DDA drives the rays of cells (i,j,k) and (i,j+1,k), which are exact translates) == the single-ray
    production march run twice: same sums and step counts, including when the two rays meet walls at different steps, reflect off
    mirrors (they separate at the first wall) or one of them starts next to an intrusion"""
    rng = np.random.default_rng(31)
    if case == "plates_mirrors":
        lv = po.parallel_plates(18, dtype=np.float32, mirror_faces=("x-", "y+"))
        thr, refl, skip = 0.001, 1, 0
    else:
        lv = po.burns_christon(20, dtype=np.float32)
        from util import add_random_walls
        add_random_walls(lv, 0.03, rng, abskg=0.8, sigmaT4=0.05)
        thr, refl, skip = 0.05, 1, (1 if case == "bc_walls" else 0)
    n = 400
    cells, dirs, org = random_rays(lv, n, rng)
    N = lv.interior_hi[1]
    ct = lv.cellType
    keep = [t for t in range(n) if cells[t][1] + 1 < N and ct[cells[t][2] + 1, cells[t][1] + 1, cells[t][0] + 1] == -1
            and ct[cells[t][2] + 1, cells[t][1] + 2, cells[t][0] + 1] == -1]
    cells, dirs, org = cells[keep], dirs[keep], org[keep]
    assert len(cells) > 250
    cellsB = cells + np.array([0, 1, 0], np.int32)
    orgB = org + np.array([0.0, lv.dx[1], 0.0])
    sa, sb, na, nb = sim.trace_1L_dual(lv, cells, dirs, org, thr, refl, skip)
    ra, nra = sim.trace_1L_fast(lv, cells, dirs, org, thr, refl, skip)
    rb, nrb = sim.trace_1L_fast(lv, cellsB, dirs, orgB, thr, refl, skip)
    # the in-cell offset of B's origin is rebuilt from an absolute position in the single march: allow FP32 round-off, not more
    assert (na == nra).mean() > 0.99 and (nb == nrb).mean() > 0.97
    same = (na == nra) & (nb == nrb)
    np.testing.assert_allclose(sa[same], ra[same], rtol=1e-5, atol=1e-6 * np.abs(ra).max())
    np.testing.assert_allclose(sb[same], rb[same], rtol=2e-4, atol=1e-5 * np.abs(rb).max())
    assert (na != nb).mean() > 0.2          # the two rays of a lane really end at different steps often


@pytest.mark.parametrize("case", ["bc_walls", "plates_mirrors", "bc_refsem"])
def test_four_cells_per_lane_march_equals_four_single_marches(po, case):
    """march_single_level_quad (one DDA drives the rays of the 2 x 2 cells (i, j..j+1, k..k+1), exact translates of each other) == the
    single-ray production march run four times: same sums and the same step total, including when the rays meet walls at different
    steps, reflect off mirrors or start next to an intrusion"""
    rng = np.random.default_rng(37)
    if case == "plates_mirrors":
        lv = po.parallel_plates(18, dtype=np.float32, mirror_faces=("x-", "z+"))
        thr, refl, skip = 0.001, 1, 0
    else:
        lv = po.burns_christon(20, dtype=np.float32)
        from util import add_random_walls
        add_random_walls(lv, 0.03, rng, abskg=0.8, sigmaT4=0.05)
        thr, refl, skip = 0.05, 1, (1 if case == "bc_walls" else 0)
    n = 500
    cells, dirs, org = random_rays(lv, n, rng)
    N = lv.interior_hi
    ct = lv.cellType
    offs = [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1)]
    keep = [t for t in range(n) if cells[t][1] + 1 < N[1] and cells[t][2] + 1 < N[2]
            and all(ct[cells[t][2] + 1 + o[2], cells[t][1] + 1 + o[1], cells[t][0] + 1] == -1 for o in offs)]
    cells, dirs, org = cells[keep], dirs[keep], org[keep]
    assert len(cells) > 250
    sm, st = sim.trace_1L_quad(lv, cells, dirs, org, thr, refl, skip)
    tot = np.zeros(len(cells), np.int64)
    ok = np.ones(len(cells), bool)
    singles = []
    for o in offs:
        r, nr = sim.trace_1L_fast(lv, cells + np.array(o, np.int32), dirs, org + np.array(o) * np.array(lv.dx), thr, refl, skip)
        singles.append(r); tot += nr
    # the in-cell offset of the translated origins is rebuilt from an absolute position in the single march: allow FP32 round-off
    same = st.astype(np.int64) == tot
    assert same.mean() > 0.95, same.mean()
    np.testing.assert_allclose(sm[same, 0], singles[0][same], rtol=1e-5, atol=1e-6 * np.abs(singles[0]).max())
    for q in (1, 2, 3):
        np.testing.assert_allclose(sm[same, q], singles[q][same], rtol=2e-4, atol=1e-5 * np.abs(singles[q]).max())


@pytest.mark.parametrize("pair_k", [0, 1])
@pytest.mark.parametrize("case", ["plates_mirrors_xy", "plates_mirrors_xz", "bc_walls"])
def test_pair_march_rejoins_after_a_wall_both_rays_hit_at_once(po, case, pair_k):
    """march_single_level_dual<REJOIN> (the variant used when walls reflect): when both rays of the pair enter a non-flow cell at the
    same step and are reflected, the joint leg goes on.  Pairs stacked in y (pair_k = 0) or in z (1): equal to two single marches --
    same sums, same step counts -- through several mirror reflections, staggered hits on the walls normal to the pair's axis and
    intrusions"""
    rng = np.random.default_rng(41 + pair_k)
    if case.startswith("plates"):
        lv = po.parallel_plates(18, dtype=np.float32, mirror_faces=("x-", "y-") if case.endswith("xy") else ("x-", "z+"))
        thr, refl, skip = 0.001, 1, 0
    else:
        lv = po.burns_christon(20, dtype=np.float32)
        from util import add_random_walls
        add_random_walls(lv, 0.03, rng, abskg=0.8, sigmaT4=0.05)
        thr, refl, skip = 0.05, 1, 0
    n = 500
    cells, dirs, org = random_rays(lv, n, rng)
    N = lv.interior_hi
    ct = lv.cellType
    o = np.array([0, 0, 1] if pair_k else [0, 1, 0], np.int32)
    keep = [t for t in range(n) if cells[t][1] + o[1] < N[1] and cells[t][2] + o[2] < N[2]
            and ct[cells[t][2] + 1, cells[t][1] + 1, cells[t][0] + 1] == -1
            and ct[cells[t][2] + 1 + o[2], cells[t][1] + 1 + o[1], cells[t][0] + 1] == -1]
    cells, dirs, org = cells[keep], dirs[keep], org[keep]
    assert len(cells) > 250
    sa, sb, na, nb = sim.trace_1L_dual(lv, cells, dirs, org, thr, refl, skip, rejoin=1, pair_k=pair_k)
    ra, nra = sim.trace_1L_fast(lv, cells, dirs, org, thr, refl, skip)
    rb, nrb = sim.trace_1L_fast(lv, cells + o, dirs, org + o * np.array(lv.dx), thr, refl, skip)
    assert (na == nra).mean() > 0.99 and (nb == nrb).mean() > 0.95, ((na == nra).mean(), (nb == nrb).mean())
    same = (na == nra) & (nb == nrb)
    np.testing.assert_allclose(sa[same], ra[same], rtol=1e-5, atol=1e-6 * np.abs(ra).max())
    np.testing.assert_allclose(sb[same], rb[same], rtol=2e-4, atol=1e-5 * np.abs(rb).max())
    if case.startswith("plates"):
        assert na.mean() > 25          # 18 cells across: two to three legs per ray, the rejoin path is what is being exercised


@pytest.mark.parametrize("case", ["bc", "plates_mirrors"])
def test_data_onion_quad_march_equals_four_single_marches(po, case):
    """march_onion2_quad (2-level Data Onion, four cells per lane: one DDA marches the four translates through the ROI copy, then each
    ray leaves the ROI / finishes on the coarse level through march_onion2_fast's continuation) == march_onion2_fast run four times"""
    rng = np.random.default_rng(53)
    fine = po.burns_christon(32, dtype=np.float32) if case == "bc" else po.parallel_plates(32, dtype=np.float32, mirror_faces=("x-", "y-"))
    thr = 0.05 if case == "bc" else 0.001
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    levels = [coarse, fine]
    offs = [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1)]
    for plo, phi in (((0, 0, 0), (16, 16, 16)), ((16, 16, 0), (32, 32, 16)), ((8, 8, 8), (24, 24, 24))):
        roiLo, roiHi, regLo, regHi = po.compute_extents(levels, plo, phi, (4, 4, 4))
        cells, dirs, org = random_rays(fine, 150, rng, plo, (phi[0], phi[1] - 1, phi[2] - 1))      # the whole 2 x 2 block inside the patch
        sm, st = sim.trace_DO2_quad(levels, (roiLo, roiHi), (regLo, regHi), cells, dirs, org, thr, 1)
        tot = np.zeros(len(cells), np.int64)
        singles = []
        for o in offs:
            s1, n1 = sim.trace_DO2_fast(levels, (roiLo, roiHi), (regLo, regHi), cells + np.array(o, np.int32), dirs,
                                        org + np.array(o) * np.array(fine.dx), thr, 1)
            singles.append(s1); tot += n1
        # the in-cell offset of a translated origin is rebuilt from an absolute position in the single march: FP32 round-off can move a
        # ray across a cell face, rarely
        same = st.astype(np.int64) == tot
        assert same.mean() > 0.9, same.mean()
        np.testing.assert_allclose(sm[same, 0], singles[0][same], rtol=2e-5, atol=2e-6 * np.abs(singles[0]).max())
        for q in (1, 2, 3):
            np.testing.assert_allclose(sm[same, q], singles[q][same], rtol=5e-4, atol=2e-5 * np.abs(singles[q]).max())


def test_blocked_addressing_algebra():
    """indep.cuh, march_leg_blk: the word index of a cell in the blocked copy is 8 * block + (x&1 | (y&1)<<1 | (z&1)<<2); a ray keeps one
    signed delta per axis, a step on an axis adds it and toggles it with ONE XOR by a constant of the level (6, 2 ^ (8 nbx - 2),
    4 ^ (8 nbx nby - 4) -- the same for both directions), a reflection steps back by the previous delta d and goes on with d - sign(d) S.
    Walks random paths with reflections and checks every index against the closed form (integer arithmetic as the device does it)."""
    rng = np.random.default_rng(0)

    def i32(v):
        return ((int(v) + 2 ** 31) % 2 ** 32) - 2 ** 31

    for nbx, nby, nbz in ((7, 5, 6), (66, 66, 66), (131, 131, 131)):
        M = nbx * nby
        dims = (2 * nbx, 2 * nby, 2 * nbz)

        def widx(x, y, z):
            return 8 * ((z >> 1) * M + (y >> 1) * nbx + (x >> 1)) + ((x & 1) | ((y & 1) << 1) | ((z & 1) << 2))

        A, S = (1, 2, 4), (8, 8 * nbx, 8 * M)
        T = (6, 2 ^ (8 * nbx - 2), 4 ^ (8 * M - 4))
        for trial in range(300):
            p = [int(rng.integers(2, d - 2)) for d in dims]
            sg = [int(rng.choice([-1, 1])) for _ in range(3)]
            # blk_stride_init
            dl = [(A[a] if not (p[a] & 1) else S[a] - A[a]) if sg[a] > 0 else (-A[a] if (p[a] & 1) else A[a] - S[a]) for a in range(3)]
            w = widx(*p)
            for step in range(40):
                a = int(rng.integers(3))
                q = list(p); q[a] += sg[a]
                if not (0 <= q[a] < dims[a]):
                    continue
                w2 = i32(w + dl[a]); dl[a] = i32(dl[a] ^ T[a])           # RB_PEEK / RB_COMMIT
                assert w2 == widx(*q)
                if rng.random() < 0.25:                                   # the cell entered is a wall: reflect (blk_ray_legs)
                    dprev = i32(dl[a] ^ T[a])
                    w2 = i32(w2 - dprev)
                    dl[a] = dprev - S[a] if dprev > 0 else dprev + S[a]
                    sg[a] = -sg[a]
                    assert w2 == widx(*p)
                else:
                    p = q
                w = w2


def _fp16_level(po, lv):
    """the level as the blocked copy sees it: abskg and sigmaT4/pi * 2^-5 rounded to fp16 (exactly representable in float)"""
    import copy
    q = copy.copy(lv)
    q.abskg = lv.abskg.astype(np.float16).astype(lv.abskg.dtype)
    q.sigmaT4 = ((lv.sigmaT4 / 32).astype(np.float16).astype(lv.sigmaT4.dtype)) * 32
    return q


@pytest.mark.parametrize("case", ["bc", "bc_walls", "plates_mirrors", "bc_refsem"])
def test_blocked_march_equals_the_fast_march_and_the_oracle(po, case):
    """march_single_level_blk (indep.cuh: the default solve's march on the 2 x 2 x 2 blocked half-packed copy, coordinate-free addressing)
    against march_single_level_fast on the same fp16-rounded properties -- same cells, same arithmetic: equal step counts, sums to FP32
    round-off -- against the oracle on those properties, and resumed leg by leg through the 12 words divq_sorted_kernel<RESORT> parks:
    bit-identical to the uninterrupted march."""
    rng = np.random.default_rng(41)
    if case == "plates_mirrors":
        lv = po.parallel_plates(17, dtype=np.float32, mirror_faces=("x-", "y+"))
        thr, refl, skip = 0.001, 1, 0
    else:
        lv = po.burns_christon(19, dtype=np.float32)       # odd sizes: partial blocks at the high faces
        if case == "bc_walls":
            from util import add_random_walls
            add_random_walls(lv, 0.03, rng, abskg=0.8, sigmaT4=0.05)
        thr, refl, skip = 0.05, 1, (0 if case == "bc_refsem" else 1)
    n = 300
    cells, dirs, org = random_rays(lv, n, rng)
    lv16 = _fp16_level(po, lv)
    sb, nb = sim.trace_1L_blk(lv, cells, dirs, org, thr, refl, skip)
    sf, nf = sim.trace_1L_fast(lv16, cells, dirs, org, thr, refl, skip)
    assert np.array_equal(nb, nf)
    np.testing.assert_allclose(sb, sf, rtol=1e-6, atol=1e-7 * np.abs(sf).max())
    # resumed after every leg / every second leg
    for legs in (1, 2):
        sr, nr = sim.trace_1L_blk(lv, cells, dirs, org, thr, refl, skip, legs_per_call=legs)
        assert np.array_equal(nr, nb) and np.array_equal(sr, sb)
    if skip == 0:      # the oracle marches every leg the reference marches
        p = po.default_params(threshold=thr, allowReflect=refl)
        ref = [po.update_sumI(p, lv16, dirs[t], org[t], cells[t]) for t in range(n)]
        rs = np.array([r[0] for r in ref]); rn = np.array([r[1] for r in ref])
        assert (rn == nb).mean() > 0.97
        same = rn == nb
        np.testing.assert_allclose(sb[same], rs[same], rtol=5e-4, atol=1e-6 * max(1.0, np.abs(rs).max()))


@pytest.mark.parametrize("case", ["bc2", "plates2", "bc3", "plates3"])
def test_blocked_data_onion_march(po, case):
    """march_onion_blk (the default Data-Onion march: one blocked half-packed box per level, marker shell = level jump, the level's own
    extra cells where a box ends at the domain boundary) with two and three levels against the oracle's updateSumI_ML (pinned bit for
    bit on the reference's own routine) on the same fp16-rounded properties: same executed steps up to FP32 ties, sums to FP32
    round-off.  Patches at the boundary, inside, and with an ROI that ends exactly at the interior's edge; zero-weight legs marched."""
    rng = np.random.default_rng(43)
    three = case.endswith("3")
    N = 48 if three else 32
    fine = po.burns_christon(N, dtype=np.float32) if case.startswith("bc") else po.parallel_plates(N, dtype=np.float32, mirror_faces=("x-", "y-"))
    thr = 0.05 if case.startswith("bc") else 0.001
    if three:
        fine.rr = (2, 2, 2)
        mid = po.coarsen_level(fine, (2, 2, 2)); mid.rr = (2, 2, 2)
        levels = [po.coarsen_level(mid, (2, 2, 2)), mid, fine]
        halo, patches = (2, 2, 2), (((0, 0, 0), (16, 16, 16)), ((16, 16, 16), (32, 32, 32)), ((2, 32, 20), (18, 48, 36)))
    else:
        fine.rr = (4, 4, 4)
        levels = [po.coarsen_level(fine, (4, 4, 4)), fine]
        halo, patches = (4, 4, 4), (((0, 0, 0), (16, 16, 16)), ((8, 8, 8), (24, 24, 24)), ((4, 4, 4), (12, 12, 12)))
    levels16 = [_fp16_level(po, l) for l in levels]
    for l16, l in zip(levels16, levels):
        l16.rr = getattr(l, "rr", (1, 1, 1))
    p = po.default_params(threshold=thr, allowReflect=1)
    for plo, phi in patches:
        roiLo, roiHi, regLo, regHi = po.compute_extents(levels, plo, phi, halo)
        cells, dirs, org = random_rays(fine, 100, rng, plo, phi)
        sb, nb = sim.trace_DO_blk(levels, (roiLo, roiHi), (regLo, regHi), cells, dirs, org, thr, 1)
        same = 0
        for t in range(len(cells)):
            s, ns, _ = po.update_sumI_ML(p, levels16, roiLo, roiHi, regLo, regHi, dirs[t], org[t], cells[t], max_path=0)
            if ns == nb[t]:
                same += 1
                assert s == pytest.approx(sb[t], rel=1e-3, abs=1e-6 * max(1.0, abs(s))), (plo, t)
        assert same >= 0.96 * len(cells), (plo, same)
