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
