"""ctypes front end of tests/hostsim/march_host.cpp: the production march/rng headers compiled for the host."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libmarch_host.so")
_lib = None


class HsLevel(C.Structure):
    _fields_ = [("rec", C.c_void_p), ("lo", C.c_int * 3), ("hi", C.c_int * 3), ("intLo", C.c_int * 3), ("intHi", C.c_int * 3),
                ("rr", C.c_int * 3), ("dx", C.c_double * 3), ("anchor", C.c_double * 3)]


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, "march_host.cpp"), os.path.join(HERE, "cuda_shim.h"),
                os.path.join(HERE, "..", "..", "uintah_b200", "csrc", "march.cuh"),
                os.path.join(HERE, "..", "..", "uintah_b200", "csrc", "rng.cuh"),
                os.path.join(HERE, "..", "..", "uintah_b200", "csrc", "indep.cuh")]
        if (not os.path.exists(SO)) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, srcs[0]], check=True)
        _lib = C.CDLL(SO)
    return _lib


def pack(lv, kind):
    n = lv.abskg.size
    out = np.zeros(n * (8 if kind == 0 else 16), np.uint8)
    ab = np.ascontiguousarray(lv.abskg, np.float64)
    s4 = np.ascontiguousarray(lv.sigmaT4, np.float64)
    lib().hs_pack(ab.ctypes.data_as(C.c_void_p), s4.ctypes.data_as(C.c_void_p), lv.cellType.ctypes.data_as(C.c_void_p),
                  C.c_longlong(n), -1, kind, out.ctypes.data_as(C.c_void_p))
    return out


def hslevel(lv, rec):
    h = HsLevel()
    h.rec = rec.ctypes.data
    h.lo = (C.c_int * 3)(*lv.lo); h.hi = (C.c_int * 3)(*lv.hi)
    h.intLo = (C.c_int * 3)(*lv.interior_lo); h.intHi = (C.c_int * 3)(*lv.interior_hi)
    h.rr = (C.c_int * 3)(*lv.rr); h.dx = (C.c_double * 3)(*lv.dx); h.anchor = (C.c_double * 3)(*lv.anchor)
    return h


def _vp(a):
    return np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


def trace_1L(lv, kind, cells, dirs, org, thr, refl, skip=0, max_path=4000, max_ray_length=0.0):
    rec = pack(lv, kind)
    h = hslevel(lv, rec)
    n = len(cells)
    sumI = np.zeros(n); steps = np.zeros(n, np.uint32); path = np.zeros((n, max_path, 4), np.int32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    lib().hs_trace_1L(C.byref(h), kind, n, _vp(cells), _vp(dirs), _vp(org), C.c_double(thr), refl, skip, _vp(sumI), _vp(steps),
                      path.ctypes.data_as(C.c_void_p), max_path, C.c_double(max_ray_length))
    return sumI, steps, path


def keyed_permutation(n, key):
    out = np.zeros(n, np.uint32)
    lib().hs_keyed_permutation(C.c_uint32(n), (C.c_uint32 * 4)(*key), _vp(out))
    return out


def trace_1L_fast(lv, cells, dirs, org, thr, refl, skip=0):
    rec = pack(lv, 0)
    h = hslevel(lv, rec)
    n = len(cells)
    sumI = np.zeros(n); steps = np.zeros(n, np.uint32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    lib().hs_trace_1L_fast(C.byref(h), n, _vp(cells), _vp(dirs), _vp(org), C.c_double(thr), refl, skip, _vp(sumI), _vp(steps))
    return sumI, steps


def trace_1L_blk(lv, cells, dirs, org, thr, refl, skip=0, legs_per_call=0):
    """the default solve's march (indep.cuh): blocked half-packed copy of the level, march_leg_blk; legs_per_call > 0: resumed leg by leg"""
    rec = pack(lv, 0)
    h = hslevel(lv, rec)
    n = len(cells)
    sumI = np.zeros(n); steps = np.zeros(n, np.uint32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    lib().hs_trace_1L_blk(C.byref(h), n, _vp(cells), _vp(dirs), _vp(org), C.c_double(thr), refl, skip, legs_per_call, _vp(sumI), _vp(steps))
    return sumI, steps


def trace_ML(levels, kind, roi, reg, cells, dirs, org, thr, refl, skip=0, max_path=4000):
    recs = [pack(l, kind) for l in levels]
    nl = len(levels)
    arr = (HsLevel * nl)(*[hslevel(l, r) for l, r in zip(levels, recs)])
    n = len(cells)
    sumI = np.zeros(n); steps = np.zeros(n, np.uint32); path = np.zeros((n, max_path, 4), np.int32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    regLo = np.ascontiguousarray(reg[0], np.int32); regHi = np.ascontiguousarray(reg[1], np.int32)
    lib().hs_trace_ML(arr, nl, (C.c_int * 3)(*roi[0]), (C.c_int * 3)(*roi[1]), _vp(regLo), _vp(regHi), kind, n, _vp(cells),
                      _vp(dirs), _vp(org), C.c_double(thr), refl, skip, _vp(sumI), _vp(steps), path.ctypes.data_as(C.c_void_p), max_path)
    return sumI, steps, path


def trace_DO_blk(levels, roi, reg, cells, dirs, org, thr, refl, skip=0):
    """the default Data-Onion march (indep.cuh, march_onion_blk): one blocked half-packed box per level, 2 to 4 levels"""
    recs = [pack(l, 0) for l in levels]
    nl = len(levels)
    arr = (HsLevel * nl)(*[hslevel(l, r) for l, r in zip(levels, recs)])
    n = len(cells)
    sumI = np.zeros(n); steps = np.zeros(n, np.uint32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    regLo = np.ascontiguousarray(reg[0], np.int32); regHi = np.ascontiguousarray(reg[1], np.int32)
    lib().hs_trace_DO_blk(arr, nl, (C.c_int * 3)(*roi[0]), (C.c_int * 3)(*roi[1]), _vp(regLo), _vp(regHi), n, _vp(cells), _vp(dirs), _vp(org),
                          C.c_double(thr), refl, skip, _vp(sumI), _vp(steps))
    return sumI, steps


def trace_DO2_fast(levels, roi, reg, cells, dirs, org, thr, refl, skip=0):
    assert len(levels) == 2
    recs = [pack(l, 0) for l in levels]
    arr = (HsLevel * 2)(*[hslevel(l, r) for l, r in zip(levels, recs)])
    n = len(cells)
    sumI = np.zeros(n); steps = np.zeros(n, np.uint32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    regLo = np.ascontiguousarray(reg[0], np.int32); regHi = np.ascontiguousarray(reg[1], np.int32)
    lib().hs_trace_DO2_fast(arr, (C.c_int * 3)(*roi[0]), (C.c_int * 3)(*roi[1]), _vp(regLo), _vp(regHi), n, _vp(cells), _vp(dirs),
                            _vp(org), C.c_double(thr), refl, skip, _vp(sumI), _vp(steps))
    return sumI, steps


def trace_DO2_quad(levels, roi, reg, cells, dirs, org, thr, refl, skip=0):
    """march_onion2_quad: the 2 x 2 cells c, c+(0,1,0), c+(0,0,1), c+(0,1,1) of the fine level share direction and in-cell origin;
    returns sums [n, 4] and the step total of the four rays [n]"""
    assert len(levels) == 2
    recs = [pack(l, 0) for l in levels]
    arr = (HsLevel * 2)(*[hslevel(l, r) for l, r in zip(levels, recs)])
    n = len(cells)
    sm = np.zeros((n, 4)); st = np.zeros(n, np.uint32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    regLo = np.ascontiguousarray(reg[0], np.int32); regHi = np.ascontiguousarray(reg[1], np.int32)
    lib().hs_trace_DO2_quad(arr, (C.c_int * 3)(*roi[0]), (C.c_int * 3)(*roi[1]), _vp(regLo), _vp(regHi), n, _vp(cells), _vp(dirs),
                            _vp(org), C.c_double(thr), refl, skip, _vp(sm), _vp(st))
    return sm, st


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    lib().hs_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return tuple(out)


def mt_first(seed, n):
    out = (C.c_uint32 * n)()
    getattr(lib(), f"hs_mt_first{n}")(C.c_uint32(seed), out)
    return [int(x) for x in out]


def trace_1L_scatter(lv, kind, cells, dirs, org, seeds, thr, refl, sigma_scat, max_ray_length=0.0):
    """march_single_level_scatter on the host with the regression-mode MT19937 stream seeded per ray"""
    rec = pack(lv, kind)
    h = hslevel(lv, rec)
    n = len(cells)
    sumI = np.zeros(n); steps = np.zeros(n, np.uint32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    seeds = np.ascontiguousarray(seeds, np.uint32)
    lib().hs_trace_1L_scatter(C.byref(h), kind, n, _vp(cells), _vp(dirs), _vp(org), _vp(seeds), C.c_double(thr), refl, C.c_double(sigma_scat),
                              C.c_double(max_ray_length), 1 if lv.is_float else 0, _vp(sumI), _vp(steps))
    return sumI, steps


def trace_1L_dual(lv, cells, dirs, org, thr, refl, skip=0, rejoin=0, pair_k=0):
    """march_single_level_dual: rays of cell c (A) and c + (0,1,0) (B) sharing direction and in-cell origin"""
    rec = pack(lv, 0)
    h = hslevel(lv, rec)
    n = len(cells)
    sa = np.zeros(n); sb = np.zeros(n); na = np.zeros(n, np.uint32); nb = np.zeros(n, np.uint32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    lib().hs_trace_1L_dual(C.byref(h), n, _vp(cells), _vp(dirs), _vp(org), C.c_double(thr), refl, skip, rejoin, pair_k, _vp(sa), _vp(sb), _vp(na), _vp(nb))
    return sa, sb, na, nb


def trace_1L_quad(lv, cells, dirs, org, thr, refl, skip=0):
    """march_single_level_quad: rays of the 2 x 2 cells c, c+(0,1,0), c+(0,0,1), c+(0,1,1) sharing direction and in-cell origin;
    returns sums [n, 4] and the step total of the four rays [n]"""
    rec = pack(lv, 0)
    h = hslevel(lv, rec)
    n = len(cells)
    sm = np.zeros((n, 4)); st = np.zeros(n, np.uint32)
    cells = np.ascontiguousarray(cells, np.int32); dirs = np.ascontiguousarray(dirs); org = np.ascontiguousarray(org)
    lib().hs_trace_1L_quad(C.byref(h), n, _vp(cells), _vp(dirs), _vp(org), C.c_double(thr), refl, skip, _vp(sm), _vp(st))
    return sm, st
