"""ctypes front end of the CPU oracle (oracle/rmcrt_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product (uintah_b200/)
never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librmcrt_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(_HERE, "rmcrt_oracle.cpp")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return _LIB_PATH


class Params(C.Structure):
    _fields_ = [("nDivQRays", C.c_int), ("nFluxRays", C.c_int), ("threshold", C.c_double),
                ("allowReflect", C.c_int), ("CCRays", C.c_int), ("isSeedRandom", C.c_int),
                ("solveDivQ", C.c_int), ("solveBoundaryFlux", C.c_int), ("fixedRandom", C.c_int),
                ("flowCell", C.c_int), ("dataOnionFluxFlowGuard", C.c_int), ("rayDirSampleAlgo", C.c_int),
                ("maxRayLength", C.c_double), ("sigmaScat", C.c_double),
                ("roiFixed", C.c_int), ("roiFixedLo", C.c_int * 3), ("roiFixedHi", C.c_int * 3)]


class CRadiometer(C.Structure):
    _fields_ = [("cellLo", C.c_int * 3), ("cellHi", C.c_int * 3), ("nRays", C.c_int), ("solidAngle", C.c_double),
                ("theta_viewAngle", C.c_double), ("range", C.c_double), ("phi_rotate", C.c_double),
                ("theta_rotate", C.c_double), ("xi_rotate", C.c_double)]


class CLevel(C.Structure):
    _fields_ = [("abskg", C.c_void_p), ("sigmaT4", C.c_void_p), ("cellType", C.c_void_p),
                ("isFloat", C.c_int), ("lo", C.c_int * 3), ("hi", C.c_int * 3),
                ("dx", C.c_double * 3), ("anchor", C.c_double * 3), ("refineRatio", C.c_int * 3),
                ("interiorLo", C.c_int * 3), ("interiorHi", C.c_int * 3)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_default_params.argtypes = [C.POINTER(Params)]
    return _lib


def default_params(**kw) -> Params:
    p = Params()
    lib().oracle_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, (C.c_int * 3)(*[int(x) for x in v]) if k in ("roiFixedLo", "roiFixedHi") else v)
    return p


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


class Level:
    """A level's property arrays, numpy [z][y][x] (x fastest) over window [lo,hi)."""

    def __init__(self, abskg, sigmaT4, cellType, lo, hi, dx, anchor=(0, 0, 0), refine_ratio=(1, 1, 1),
                 interior_lo=None, interior_hi=None):
        self.abskg = np.ascontiguousarray(abskg)
        self.sigmaT4 = np.ascontiguousarray(sigmaT4, dtype=self.abskg.dtype)
        self.cellType = np.ascontiguousarray(cellType, dtype=np.int32)
        assert self.abskg.dtype in (np.float32, np.float64)
        self.lo = tuple(int(x) for x in lo)
        self.hi = tuple(int(x) for x in hi)
        shape = (self.hi[2] - self.lo[2], self.hi[1] - self.lo[1], self.hi[0] - self.lo[0])
        assert self.abskg.shape == shape and self.sigmaT4.shape == shape and self.cellType.shape == shape, \
            (self.abskg.shape, shape)
        self.dx = tuple(float(x) for x in dx)
        self.anchor = tuple(float(x) for x in anchor)
        self.rr = tuple(int(x) for x in refine_ratio)
        self.interior_lo = tuple(interior_lo) if interior_lo is not None else tuple(x + 1 for x in self.lo)
        self.interior_hi = tuple(interior_hi) if interior_hi is not None else tuple(x - 1 for x in self.hi)

    @property
    def is_float(self):
        return self.abskg.dtype == np.float32

    def c(self) -> CLevel:
        L = CLevel()
        L.abskg = self.abskg.ctypes.data
        L.sigmaT4 = self.sigmaT4.ctypes.data
        L.cellType = self.cellType.ctypes.data
        L.isFloat = 1 if self.is_float else 0
        L.lo = _i3(self.lo); L.hi = _i3(self.hi)
        L.dx = _d3(self.dx); L.anchor = _d3(self.anchor)
        L.refineRatio = _i3(self.rr)
        L.interiorLo = _i3(self.interior_lo); L.interiorHi = _i3(self.interior_hi)
        return L


# ---------------------------------------------------------------- MT19937
def mt_u32(seed, n):
    out = np.empty(n, np.uint32)
    lib().oracle_mt_u32(C.c_uint32(seed), n, out.ctypes.data_as(C.c_void_p))
    return out


def mt_rand(seed, n):
    out = np.empty(n, np.float64)
    lib().oracle_mt_rand(C.c_uint32(seed), n, out.ctypes.data_as(C.c_void_p))
    return out


def mt_randDblExc(seed, n):
    out = np.empty(n, np.float64)
    lib().oracle_mt_randDblExc(C.c_uint32(seed), n, out.ctypes.data_as(C.c_void_p))
    return out


# ---------------------------------------------------------------- problem set-up
SIGMA = 5.67051e-8  # Ray.cc:165


def sigmaT4(temp, dtype=np.float64, sigma=SIGMA):
    temp = np.ascontiguousarray(temp, dtype=np.float64)
    out = np.empty(temp.shape, dtype)
    lib().oracle_sigmaT4(temp.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                         1 if dtype == np.float32 else 0, C.c_int64(temp.size), C.c_double(sigma))
    return out


def burns_christon(N, dtype=np.float64, T_gas=64.804, abskg_wall=1.0, T_wall=0.0, wall_type=8) -> Level:
    """RMCRT_bm1_1L.ups at resolution N^3: unit cube, one layer of extra cells, cold black walls."""
    M = N + 2
    abskg = np.empty((M, M, M), np.float64)
    temp = np.empty((M, M, M), np.float64)
    ct = np.empty((M, M, M), np.int32)
    lib().oracle_burns_christon(N, abskg.ctypes.data_as(C.c_void_p), temp.ctypes.data_as(C.c_void_p),
                                ct.ctypes.data_as(C.c_void_p), C.c_double(T_gas), C.c_double(abskg_wall),
                                C.c_double(T_wall), wall_type)
    s4 = sigmaT4(temp, dtype)
    return Level(abskg.astype(dtype), s4, ct, (-1, -1, -1), (N + 1, N + 1, N + 1), (1.0 / N,) * 3,
                 interior_lo=(0, 0, 0), interior_hi=(N, N, N))


def parallel_plates(N, dtype=np.float64, T_gas=1500.0, T_wall=1000.0, kappa=1.0, mirror_faces=("x-", "y-"),
                    wall_type=8, black_abskg=1.0) -> Level:
    """Benchmark 2 as configured by RMCRT_1L_perf.ups:24-90,108-110: abskg==1, T_m 1500 K, walls 1000 K,
    listed faces are perfect mirrors (abskg 0), the others black."""
    M = N + 2
    abskg = np.full((M, M, M), kappa, np.float64)
    temp = np.full((M, M, M), T_gas, np.float64)
    ct = np.full((M, M, M), -1, np.int32)
    faces = {"x-": (2, 0), "x+": (2, M - 1), "y-": (1, 0), "y+": (1, M - 1), "z-": (0, 0), "z+": (0, M - 1)}
    for name, (axis, idx) in faces.items():
        sl = [slice(None)] * 3
        sl[axis] = idx
        sl = tuple(sl)
        ct[sl] = wall_type
        temp[sl] = T_wall
    # abskg on walls: black first then mirrors, so shared edges follow the face order x-,x+,y-,y+,z-,z+ of setBC
    for name in ("x-", "x+", "y-", "y+", "z-", "z+"):
        axis, idx = faces[name]
        sl = [slice(None)] * 3
        sl[axis] = idx
        abskg[tuple(sl)] = 0.0 if name in mirror_faces else black_abskg
    s4 = sigmaT4(temp, dtype)
    return Level(abskg.astype(dtype), s4, ct, (-1, -1, -1), (N + 1, N + 1, N + 1), (1.0 / N,) * 3,
                 interior_lo=(0, 0, 0), interior_hi=(N, N, N))


def benchmark_level(bm, N, pDir=0, dtype=np.float64, wall_type=8) -> Level:
    """RMCRT_Test::initialize benchmark fields (Examples/RMCRT_Test.cc:653-779) on the unit cube at N^3 with cold black
    walls (abskg 1, T 0: the Dirichlet values of RMCRT_bm1_1L.ups / RMCRT_bm3.ups).  bm: 1 Burns & Christon; 2 abskg==1;
    3 Burns & Christon abskg with T = 1000*abskg; 6 sinusoidal T; 7 exponential abskg, linear T^4; 8 hot and cold clouds."""
    base = burns_christon(N, np.float64, wall_type=wall_type)
    flow = base.cellType == -1
    dx = 1.0 / N
    c = 0.0 + dx * np.arange(-1, N + 1) + dx * 0.5    # Level::getCellPosition (Level.cc:473-477): anchor + dcell*v + dcell*0.5
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    p = (x, y, z)[pDir]
    ab = base.abskg.copy()
    T = np.full(ab.shape, 64.804)
    if bm == 1:
        pass
    elif bm == 2:
        ab[:] = 1.0                                                      # 683-685
    elif bm == 3:
        T = 1000.0 * ab                                                  # 687-692
    elif bm == 6:
        ab[:] = 5.0                                                      # 710-731
        T = (1e7 * (1.0 + np.sin(4.5 * 2 * np.pi * p))) ** 0.25
    elif bm == 7:
        ab = 5.0 * np.exp(-2 * p - 0.1)                                  # 736-756
        T = (1.0e7 * (2.0 * p + 1.0)) ** 0.25
    elif bm == 8:                                                        # 759-779
        pi, mx = np.pi, np.maximum
        ab = (mx(np.sin(x * 5 * pi) * np.abs(np.sin(z * 3 * pi + pi) + np.sin(y * 3 * pi)), 0.0)
              + mx(np.sin(x * 5.0 * pi + pi), 0.0) * mx(np.sin(z * 3 * pi + pi) + np.sin(y * 3 * pi + pi), 0.0)) * 5.0
        T = (mx(np.sin(x * 5.0 * pi) * mx(np.sin(z * 3.0 * pi) + np.sin(y * 3.0 * pi), 0.0), 0.0)
             + mx(np.sin(x * 5.0 * pi), 0.0) * mx(np.sin(z * 3.0 * pi + pi) + np.sin(y * 3.0 * pi + pi), 0.0)) * 1000.0
    else:
        raise ValueError(f"benchmark {bm}: not initialised here (4 and 5 are the RAY_SCATTER problems)")
    ab = np.where(flow, ab, 1.0)
    T = np.where(flow, T, 0.0)
    return Level(ab.astype(dtype), sigmaT4(T, dtype), base.cellType, base.lo, base.hi, base.dx,
                 interior_lo=(0, 0, 0), interior_hi=(N, N, N))


def centreline_box(N, axis):
    """origin cells RMCRT.m:387-398 extracts with lineextract: the centre line along `axis` (0 x, 1 y, 2 z)"""
    m = N // 2
    lo, hi = [m, m, m], [m + 1, m + 1, m + 1]
    lo[axis], hi[axis] = 0, N
    return tuple(lo), tuple(hi)


def centreline_of(field_level, N, axis):
    """the same line out of an (N+2)^3 level window [z][y][x]"""
    m = N // 2
    sl = [m + 1, m + 1, m + 1]
    sl[2 - axis] = slice(1, N + 1)
    return field_level[tuple(sl)]


_GOLDEN = None


def golden_benchmarks():
    """tests/golden/rmcrt_benchmarks.json (made from the reference's postProcessTools tables by make_benchmark_golden.py)"""
    global _GOLDEN
    if _GOLDEN is None:
        import json
        with open(os.path.join(os.path.dirname(_HERE), "tests", "golden", "rmcrt_benchmarks.json")) as f:
            _GOLDEN = json.load(f)
    return _GOLDEN


def benchmark_exact(bm, N, axis):
    """exact divQ at the N cell centres of the centre line, in Ray.cc:775's sign, for the problem benchmark_level builds.
    bm 3: RMCRT.m:100-135 (243-point table, sub-sampled for 81/27/9/3, pchip otherwise); bm 7/8: RMCRT.m:202-277 (pchip)."""
    from scipy.interpolate import PchipInterpolator
    g = golden_benchmarks()
    xc = (np.arange(N) + 0.5) / N
    if bm == 3:
        ex = -np.array(g["bench3"]["values"])
        sub = {243: ex}
        sub[81] = sub[243][1:242:3]; sub[27] = sub[81][1:80:3]; sub[9] = sub[27][1:26:3]; sub[3] = sub[9][1:8:3]
        if N in sub:
            return sub[N]
        return PchipInterpolator((np.arange(243) + 0.5) / 243, ex)(xc)
    name, scale = {7: ("kexp", g["kexp"]["scale_to_code_problem"]), 8: ("clouds", 1.0)}[bm]
    t = g[name]["xyz"[axis]]
    return scale * PchipInterpolator(np.array(t["x"]), np.array(t["divQ"]))(xc)


# ---------------------------------------------------------------- single rays
def update_sumI(p: Params, lv: Level, direction, ray_origin, origin_cell, max_path=0):
    cl = lv.c()
    sumI = C.c_double(0)
    n = C.c_uint64(0)
    npath = C.c_int(0)
    path = np.zeros((max(max_path, 1), 3), np.int32)
    lib().oracle_updateSumI(C.byref(p), C.byref(cl), _d3(direction), _d3(ray_origin), _i3(origin_cell),
                            C.byref(sumI), C.byref(n), path.ctypes.data_as(C.c_void_p) if max_path else None,
                            max_path, C.byref(npath))
    return sumI.value, n.value, path[:min(npath.value, max_path)]


def compute_extents(levels, patch_lo, patch_hi, halo, fixed_lo=None, fixed_hi=None):
    """Ray::computeExtents: patch_based ROI (patch +- halo), or the fixed cell box [fixed_lo, fixed_hi)"""
    nl = len(levels)
    arr = (CLevel * nl)(*[l.c() for l in levels])
    roiLo = (C.c_int * 3)(); roiHi = (C.c_int * 3)()
    regLo = (C.c_int * (3 * nl))(); regHi = (C.c_int * (3 * nl))()
    lib().oracle_computeExtents_roi(arr, nl, _i3(patch_lo), _i3(patch_hi), _i3(halo), _i3(fixed_lo) if fixed_lo is not None else None,
                                    _i3(fixed_hi) if fixed_hi is not None else None, roiLo, roiHi, regLo, regHi)
    return (tuple(roiLo), tuple(roiHi), np.array(regLo).reshape(nl, 3).copy(), np.array(regHi).reshape(nl, 3).copy())


def update_sumI_ML(p: Params, levels, roi_lo, roi_hi, region_lo, region_hi, direction, ray_origin, origin_cell,
                   max_path=0):
    nl = len(levels)
    arr = (CLevel * nl)(*[l.c() for l in levels])
    regLo = np.ascontiguousarray(region_lo, np.int32)
    regHi = np.ascontiguousarray(region_hi, np.int32)
    sumI = C.c_double(0)
    n = C.c_uint64(0)
    npath = C.c_int(0)
    path = np.zeros((max(max_path, 1), 4), np.int32)
    lib().oracle_updateSumI_ML(C.byref(p), arr, nl, _i3(roi_lo), _i3(roi_hi), regLo.ctypes.data_as(C.c_void_p),
                               regHi.ctypes.data_as(C.c_void_p), _d3(direction), _d3(ray_origin), _i3(origin_cell),
                               C.byref(sumI), C.byref(n), path.ctypes.data_as(C.c_void_p) if max_path else None,
                               max_path, C.byref(npath))
    return sumI.value, n.value, path[:min(npath.value, max_path)]


# ---------------------------------------------------------------- the reference's own per-ray routines (oracle/_ref/libref_rmcrt.so)
_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libref_rmcrt.so")
_ref = None


def ref_lib():
    """the reference's updateSumI / updateSumI_ML / findRayDirection / ray_Origin compiled from the reference tree by `make -C oracle ref`
    (oracle/extract_ref.py + ref_rmcrt_shim.cpp); None when it was never built (no reference tree at build time)"""
    global _ref
    if _ref is None and os.path.exists(_REF_LIB_PATH):
        _ref = C.CDLL(_REF_LIB_PATH)
    return _ref


_ref_scatter = None


def ref_scatter_lib():
    """the same routines compiled with -DRAY_SCATTER (the reference's --enable-ray-scatter build)"""
    global _ref_scatter
    path = os.path.join(_HERE, "_ref", "libref_rmcrt_scatter.so")
    if _ref_scatter is None and os.path.exists(path):
        _ref_scatter = C.CDLL(path)
    return _ref_scatter


def update_sumI_seeded(p: Params, lv: Level, direction, ray_origin, origin_cell, seed, reference=None):
    """one ray with the task generator seeded (scattering draws from it inside the march); reference = a ref_lib() handle or None = oracle"""
    cl = lv.c()
    sumI = C.c_double(0)
    n = C.c_uint64(0)
    fn = lib().oracle_updateSumI_seeded if reference is None else reference.ref_updateSumI_seeded
    fn(C.byref(p), C.byref(cl), _d3(direction), _d3(ray_origin), _i3(origin_cell), C.c_uint32(seed), C.byref(sumI), C.byref(n))
    return sumI.value, n.value


def ref_update_sumI(p: Params, lv: Level, direction, ray_origin, origin_cell):
    cl = lv.c()
    sumI = C.c_double(0)
    n = C.c_uint64(0)
    rc = ref_lib().ref_updateSumI(C.byref(p), C.byref(cl), _d3(direction), _d3(ray_origin), _i3(origin_cell), C.byref(sumI), C.byref(n))
    return sumI.value, n.value, rc


def ref_update_sumI_ML(p: Params, levels, roi_lo, roi_hi, region_lo, region_hi, direction, ray_origin, origin_cell):
    nl = len(levels)
    arr = (CLevel * nl)(*[l.c() for l in levels])
    regLo = np.ascontiguousarray(region_lo, np.int32)
    regHi = np.ascontiguousarray(region_hi, np.int32)
    sumI = C.c_double(0)
    n = C.c_uint64(0)
    rc = ref_lib().ref_updateSumI_ML(C.byref(p), arr, nl, _i3(roi_lo), _i3(roi_hi), regLo.ctypes.data_as(C.c_void_p), regHi.ctypes.data_as(C.c_void_p),
                                     _d3(direction), _d3(ray_origin), _i3(origin_cell), C.byref(sumI), C.byref(n))
    return sumI.value, n.value, rc


def ref_ray_trace(p: Params, lv: Level, patch_lo=None, patch_hi=None, seed=1, patches=(1, 1, 1), threads=1, library=None):
    """the divQ part of Ray::rayTrace with the reference's own per-ray routines (ref_rayTrace_divQ); same task split / seeds as ray_trace"""
    patch_lo = lv.interior_lo if patch_lo is None else tuple(patch_lo)
    patch_hi = lv.interior_hi if patch_hi is None else tuple(patch_hi)
    shape = lv.abskg.shape
    divQ = np.zeros(shape, np.float64)
    volq = np.zeros(shape, np.float64)
    cl = lv.c()
    tasks = _split(patch_lo, patch_hi, patches)
    steps = [0] * len(tasks)

    def run(t):
        lo, hi = tasks[t]
        n = C.c_uint64(0)
        rc = (library or ref_lib()).ref_rayTrace_divQ(C.byref(p), C.byref(cl), _i3(lo), _i3(hi), C.c_uint32(seed + t), divQ.ctypes.data_as(C.c_void_p),
                                         volq.ctypes.data_as(C.c_void_p), _i3(lv.lo), _i3(lv.hi), C.byref(n))
        assert rc == 0
        steps[t] = n.value

    if threads > 1 and len(tasks) > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(run, range(len(tasks))))
    else:
        for t in range(len(tasks)):
            run(t)
    return dict(divQ=divQ, radiationVolq=volq, steps=sum(steps))


def ref_ray_trace_boundflux(p: Params, lv: Level, patch_lo=None, patch_hi=None, seed=1):
    """the boundary-flux part of Ray::rayTrace (Ray.cc:645-730) with the reference's OWN has_a_boundary / rayDirection_cellFace /
    rayLocation_cellFace / updateSumI and face tables (ref_rayTrace_boundFlux); one task, MT seed `seed`.  Returns (boundFlux, steps)."""
    patch_lo = lv.interior_lo if patch_lo is None else tuple(patch_lo)
    patch_hi = lv.interior_hi if patch_hi is None else tuple(patch_hi)
    bflux = np.zeros(lv.abskg.shape + (7,), np.float64)
    cl = lv.c()
    n = C.c_uint64(0)
    rc = ref_lib().ref_rayTrace_boundFlux(C.byref(p), C.byref(cl), _i3(patch_lo), _i3(patch_hi), C.c_uint32(seed), bflux.ctypes.data_as(C.c_void_p),
                                          _i3(lv.lo), _i3(lv.hi), C.byref(n))
    assert rc == 0
    return bflux, n.value


def ref_compute_extents(levels, patch_lo, patch_hi, halo, roi_algo="patch_based", roi_min=None, roi_max=None):
    """Ray::computeExtents, the reference's own code (Ray.cc:1279-1358).  roi_algo: 'fixed' (roi_min / roi_max physical points),
    'dynamic' (roi_min / roi_max = the ROI_loCell / ROI_hiCell reduction values) or 'patch_based'.  Returns the tuple of
    compute_extents, or None when the reference throws (fixed ROI outside the fine level)."""
    nl = len(levels)
    arr = (CLevel * nl)(*[l.c() for l in levels])
    roiLo = (C.c_int * 3)(); roiHi = (C.c_int * 3)()
    regLo = (C.c_int * (3 * nl))(); regHi = (C.c_int * (3 * nl))()
    algo = {"fixed": 0, "dynamic": 1, "patch_based": 2}[roi_algo]
    rc = ref_lib().ref_computeExtents(arr, nl, _i3(patch_lo), _i3(patch_hi), _i3(halo), algo, _d3(roi_min) if roi_min is not None else None,
                                      _d3(roi_max) if roi_max is not None else None, roiLo, roiHi, regLo, regHi)
    if rc != 0:
        return None
    return (tuple(roiLo), tuple(roiHi), np.array(regLo).reshape(nl, 3).copy(), np.array(regHi).reshape(nl, 3).copy())


def generate_rays(p: Params, lv: Level, cell, n_rays, seed=1, reference=False):
    """(dirs, origins) of rays 0..n_rays-1 of one cell as the divQ loop draws them (Ray.cc:752-760); reference=True: the reference's own code"""
    cl = lv.c()
    dirs = np.zeros((n_rays, 3)); orgs = np.zeros((n_rays, 3))
    fn = ref_lib().ref_generate_rays if reference else lib().oracle_generate_rays
    fn(C.byref(p), C.byref(cl), _i3(cell), int(n_rays), C.c_uint32(seed), dirs.ctypes.data_as(C.c_void_p), orgs.ctypes.data_as(C.c_void_p))
    return dirs, orgs


# ---------------------------------------------------------------- radiometer / LHC
def radiometer(view_angle, vec, cell_lo, cell_hi, n_rays, is_normal_vector=False) -> CRadiometer:
    """one <Radiometer> block: angles per Radiometer::problemSetup, cell box = getCellIndex(locationsMin/Max)"""
    r = CRadiometer()
    lib().oracle_radiometer_setup(C.c_double(view_angle), _d3(vec), 1 if is_normal_vector else 0, C.byref(r))
    r.cellLo = _i3(cell_lo); r.cellHi = _i3(cell_hi); r.nRays = int(n_rays)
    return r


def radiometer_flux(p: Params, lv: Level, rads, patch_lo=None, patch_hi=None, seed=1):
    """Radiometer::radiometerFlux over one patch; returns (VRFlux, VRIntensity, steps) over the level window, dtype of lv"""
    patch_lo = lv.interior_lo if patch_lo is None else tuple(patch_lo)
    patch_hi = lv.interior_hi if patch_hi is None else tuple(patch_hi)
    flux = np.zeros(lv.abskg.shape, lv.abskg.dtype)
    inten = np.zeros(lv.abskg.shape, lv.abskg.dtype)
    arr = (CRadiometer * len(rads))(*rads)
    cl = lv.c()
    n = C.c_uint64(0)
    lib().oracle_radiometerFlux(C.byref(p), C.byref(cl), arr, len(rads), _i3(patch_lo), _i3(patch_hi), C.c_uint32(seed),
                                flux.ctypes.data_as(C.c_void_p), inten.ctypes.data_as(C.c_void_p), _i3(lv.lo), _i3(lv.hi),
                                C.byref(n))
    return flux, inten, n.value


def rand_vector(p: Params, cell, n, seed=1):
    out = np.zeros(n, np.int32)
    lib().oracle_randVector(C.byref(p), C.c_uint32(seed), _i3(cell), n, out.ctypes.data_as(C.c_void_p))
    return out


# ---------------------------------------------------------------- patch solves
def _split(lo, hi, parts):
    """Split box [lo,hi) into parts[0]*parts[1]*parts[2] patches like <patches> does."""
    edges = []
    for d in range(3):
        n = hi[d] - lo[d]
        e = [lo[d] + (n * i) // parts[d] for i in range(parts[d] + 1)]
        edges.append(e)
    out = []
    for k in range(parts[2]):
        for j in range(parts[1]):
            for i in range(parts[0]):
                out.append(((edges[0][i], edges[1][j], edges[2][k]), (edges[0][i + 1], edges[1][j + 1], edges[2][k + 1])))
    return out


def ray_trace(p: Params, lv: Level, patch_lo=None, patch_hi=None, seed=1, patches=(1, 1, 1), threads=1):
    """Ray::rayTrace over [patch_lo,patch_hi) (default: interior), split into `patches` tasks.
    Task t uses MT seed `seed + t` when isSeedRandom.  Returns dict(divQ, radiationVolq, boundFlux, steps, fluxSteps);
    output arrays cover the whole level window (extra cells stay 0)."""
    patch_lo = lv.interior_lo if patch_lo is None else tuple(patch_lo)
    patch_hi = lv.interior_hi if patch_hi is None else tuple(patch_hi)
    shape = lv.abskg.shape
    divQ = np.zeros(shape, np.float64)
    volq = np.zeros(shape, np.float64)
    bflux = np.zeros(shape + (7,), np.float64)
    cl = lv.c()
    tasks = _split(patch_lo, patch_hi, patches)
    steps = [0] * len(tasks)
    fsteps = [0] * len(tasks)

    def run(t):
        lo, hi = tasks[t]
        n = C.c_uint64(0)
        nf = C.c_uint64(0)
        lib().oracle_rayTrace(C.byref(p), C.byref(cl), _i3(lo), _i3(hi), C.c_uint32(seed + t),
                              divQ.ctypes.data_as(C.c_void_p), volq.ctypes.data_as(C.c_void_p),
                              bflux.ctypes.data_as(C.c_void_p), _i3(lv.lo), _i3(lv.hi), C.byref(n), C.byref(nf))
        steps[t] = n.value
        fsteps[t] = nf.value

    if threads > 1 and len(tasks) > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(run, range(len(tasks))))
    else:
        for t in range(len(tasks)):
            run(t)
    return dict(divQ=divQ, radiationVolq=volq, boundFlux=bflux, steps=sum(steps), fluxSteps=sum(fsteps))


def ray_trace_data_onion_sub(p: Params, levels, halo, patch_lo, patch_hi, sub_lo, sub_hi, seed=1, parts=(1, 1, 1), threads=1):
    """Ray::rayTrace_dataOnion of ONE fine patch [patch_lo,patch_hi) (ROI and region extents of that patch), restricted to the origin
    cells of [sub_lo,sub_hi); the sub-box is split into `parts` pieces run on `threads` threads (piece t: MT seed `seed + t`)."""
    fine = levels[-1]
    nl = len(levels)
    arr = (CLevel * nl)(*[l.c() for l in levels])
    shape = fine.abskg.shape
    divQ = np.zeros(shape, np.float64)
    volq = np.zeros(shape, np.float64)
    bflux = np.zeros(shape + (7,), np.float64) if p.solveBoundaryFlux else np.zeros((1, 7), np.float64)
    tasks = _split(tuple(sub_lo), tuple(sub_hi), parts)
    steps = [0] * len(tasks)

    def run(t):
        lo, hi = tasks[t]
        n = C.c_uint64(0)
        nf = C.c_uint64(0)
        lib().oracle_rayTrace_dataOnion_sub(C.byref(p), arr, nl, _i3(patch_lo), _i3(patch_hi), _i3(halo), _i3(lo), _i3(hi), C.c_uint32(seed + t),
                                            divQ.ctypes.data_as(C.c_void_p), volq.ctypes.data_as(C.c_void_p),
                                            bflux.ctypes.data_as(C.c_void_p), _i3(fine.lo), _i3(fine.hi), C.byref(n), C.byref(nf))
        steps[t] = n.value

    with ThreadPoolExecutor(max_workers=max(threads, 1)) as ex:
        list(ex.map(run, range(len(tasks))))
    return dict(divQ=divQ, radiationVolq=volq, boundFlux=bflux if p.solveBoundaryFlux else None, steps=sum(steps))


def ray_trace_data_onion(p: Params, levels, halo, patch_lo=None, patch_hi=None, seed=1, patches=(1, 1, 1), threads=1):
    """Ray::rayTrace_dataOnion over the fine level's interior split into `patches` fine patches."""
    fine = levels[-1]
    patch_lo = fine.interior_lo if patch_lo is None else tuple(patch_lo)
    patch_hi = fine.interior_hi if patch_hi is None else tuple(patch_hi)
    nl = len(levels)
    arr = (CLevel * nl)(*[l.c() for l in levels])
    shape = fine.abskg.shape
    divQ = np.zeros(shape, np.float64)
    volq = np.zeros(shape, np.float64)
    bflux = np.zeros(shape + (7,), np.float64)
    tasks = _split(patch_lo, patch_hi, patches)
    steps = [0] * len(tasks)
    fsteps = [0] * len(tasks)

    def run(t):
        lo, hi = tasks[t]
        n = C.c_uint64(0)
        nf = C.c_uint64(0)
        lib().oracle_rayTrace_dataOnion(C.byref(p), arr, nl, _i3(lo), _i3(hi), _i3(halo), C.c_uint32(seed + t),
                                        divQ.ctypes.data_as(C.c_void_p), volq.ctypes.data_as(C.c_void_p),
                                        bflux.ctypes.data_as(C.c_void_p), _i3(fine.lo), _i3(fine.hi),
                                        C.byref(n), C.byref(nf))
        steps[t] = n.value
        fsteps[t] = nf.value

    if threads > 1 and len(tasks) > 1:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(run, range(len(tasks))))
    else:
        for t in range(len(tasks)):
            run(t)
    return dict(divQ=divQ, radiationVolq=volq, boundFlux=bflux, steps=sum(steps), fluxSteps=sum(fsteps))


def roi_extents(lv: Level, patches, abskg_thld, sigmaT4_thld):
    """Ray::ROI_Extents (Ray.cc:1984-2043) with compute_Mag_gradient (Core/Grid/AMR.cc:464-483): for every patch (lo, hi, extraLo, extraHi)
    the first-order central-difference gradient magnitudes of abskg and sigmaT4 (T inv_dx = (T)0.5/dx, differences and product in T,
    Vector::length() in double, stored as T), the flags mag > threshold, and the ROI_loCell / ROI_hiCell reduction = the union of the
    extra-cell boxes of the patches that hold a flagged cell (SHRT_MAX / -SHRT_MAX when there is none).
    Returns (mag_grad_abskg, mag_grad_sigmaT4, flaggedCells) over the level window (zero outside the patches) and (roi_lo, roi_hi)."""
    T = lv.abskg.dtype.type
    mags = []
    for q in (lv.abskg, lv.sigmaT4):
        g2 = np.zeros(q[1:-1, 1:-1, 1:-1].shape, np.float64)
        for axis, d in ((2, 0), (1, 1), (0, 2)):          # array axis 2 = x
            inv = T(np.float64(0.5) / np.float64(lv.dx[d]))      # (T)0.5 / dx: a double division, rounded to T
            r = [slice(1, -1)] * 3; l = [slice(1, -1)] * 3
            r[axis] = slice(2, None); l[axis] = slice(0, -2)
            g = ((q[tuple(r)] - q[tuple(l)]).astype(T) * inv).astype(T).astype(np.float64)
            g2 = g2 + g * g if d == 0 else (g2 + g * g)
        m = np.zeros(q.shape, T)
        m[1:-1, 1:-1, 1:-1] = np.sqrt(g2).astype(T)
        mags.append(m)
    # (x*x + y*y) + z*z in that order, like Vector::length()
    flagged = np.zeros(lv.abskg.shape, np.int32)
    roi_lo, roi_hi = [32767] * 3, [-32767] * 3
    outA, outS = np.zeros_like(mags[0]), np.zeros_like(mags[1])
    for lo, hi, elo, ehi in patches:
        sl = tuple(slice(lo[d] - lv.lo[d], hi[d] - lv.lo[d]) for d in (2, 1, 0))
        outA[sl] = mags[0][sl]; outS[sl] = mags[1][sl]
        f = (mags[0][sl].astype(np.float64) > abskg_thld) | (mags[1][sl].astype(np.float64) > sigmaT4_thld)
        flagged[sl] = f
        if f.any():
            roi_lo = [min(a, b) for a, b in zip(roi_lo, elo)]
            roi_hi = [max(a, b) for a, b in zip(roi_hi, ehi)]
    return outA, outS, flagged, (tuple(roi_lo), tuple(roi_hi))


def coarsen_level(fine: Level, rr=(4, 4, 4), logic=0, wall_abskg=None, wall_sigmaT4=None, wall_type=None, coarsen_extra_cells=False) -> Level:
    """Build the coarse level the Data-Onion needs: Ray::coarsen_Q interior average of abskg and sigmaT4 and
    Ray::computeCellType; coarse extra cells are then filled the way sched_setBoundaryConditions does for
    Dirichlet walls (taken from the fine level's extra-cell corner value unless given) -- or, with coarsen_extra_cells
    (<coarsenExtraCells> true, Ray.cc:2148-2245), with the average of the fine extra cells each coarse face cell covers (ratio 1 along
    the face normal; faces without edges, edge / corner extra cells stay 0; accumulation in T, scaled by the double 1/RR)."""
    fl, fh = fine.interior_lo, fine.interior_hi
    cil = tuple(fl[d] // rr[d] for d in range(3))
    cih = tuple(fh[d] // rr[d] for d in range(3))
    clo = tuple(x - 1 for x in cil)
    chi = tuple(x + 1 for x in cih)
    shape = (chi[2] - clo[2], chi[1] - clo[1], chi[0] - clo[0])
    dt = fine.abskg.dtype
    ab = np.zeros(shape, dt)
    s4 = np.zeros(shape, dt)
    ct = np.zeros(shape, np.int32)
    isf = 1 if fine.is_float else 0
    for src, dst in ((fine.abskg, ab), (fine.sigmaT4, s4)):
        lib().oracle_coarsen_interior(src.ctypes.data_as(C.c_void_p), isf, _i3(fine.lo), _i3(fine.hi),
                                      dst.ctypes.data_as(C.c_void_p), _i3(clo), _i3(chi), _i3(cil), _i3(cih), _i3(rr))
    lib().oracle_computeCellType(fine.cellType.ctypes.data_as(C.c_void_p), _i3(fine.lo), _i3(fine.hi),
                                 ct.ctypes.data_as(C.c_void_p), _i3(clo), _i3(chi), _i3(cil), _i3(cih), _i3(rr), logic)
    # extra cells: copy the fine level's face values (uniform Dirichlet in all BASELINE decks)
    def face_fill(dst, src, override):
        M = dst.shape
        for axis, idx, fidx in ((2, 0, 0), (2, M[2] - 1, -1), (1, 0, 0), (1, M[1] - 1, -1), (0, 0, 0), (0, M[0] - 1, -1)):
            sl = [slice(None)] * 3
            sl[axis] = idx
            fsl = [1, 1, 1]
            fsl[axis] = fidx
            dst[tuple(sl)] = src[tuple(fsl)] if override is None else override
    face_fill(ct, fine.cellType, wall_type)
    if coarsen_extra_cells:
        for src, dst in ((fine.abskg, ab), (fine.sigmaT4, s4)):
            for a in range(3):                      # array axis 0 = z, 1 = y, 2 = x
                t1, t2 = [x for x in range(3) if x != a]
                r1, r2 = rr[2 - t1], rr[2 - t2]
                for cidx, fidx in ((0, 0), (dst.shape[a] - 1, src.shape[a] - 1)):
                    fsl = [slice(1, -1)] * 3; fsl[a] = fidx
                    face = src[tuple(fsl)]                                  # fine extra cells of the face without its edges, 2-D
                    n1, n2 = face.shape[0] // r1, face.shape[1] // r2
                    acc = np.zeros((n1, n2), dt)
                    for i1 in range(r1):            # the reference's CellIterator runs x fastest: same summation order per coarse cell
                        for i2 in range(r2):
                            acc = (acc + face[i1::r1, i2::r2][:n1, :n2]).astype(dt)
                    val = (acc.astype(np.float64) * (1.0 / float(r1 * r2))).astype(dt)
                    csl = [slice(1, -1)] * 3; csl[a] = cidx
                    dst[tuple(csl)] = val
    else:
        face_fill(ab, fine.abskg, wall_abskg)
        face_fill(s4, fine.sigmaT4, wall_sigmaT4)
    cdx = tuple(fine.dx[d] * rr[d] for d in range(3))
    return Level(ab, s4, ct, clo, chi, cdx, fine.anchor, (1, 1, 1), cil, cih)


# ---------------------------------------------------------------- RMCRT_coarseLevel: Ray::refine_Q; Arches property tasks
def refine_Q(coarseQ, c_lo, c_hi, grid_lo, grid_hi, rr, order, fl, fh, out_lo, out_hi):
    q = np.ascontiguousarray(coarseQ, np.float64)
    assert q.shape == (c_hi[2] - c_lo[2], c_hi[1] - c_lo[1], c_hi[0] - c_lo[0])
    out = np.zeros((out_hi[2] - out_lo[2], out_hi[1] - out_lo[1], out_hi[0] - out_lo[0]), np.float64)
    lib().oracle_refine_Q(q.ctypes.data_as(C.c_void_p), _i3(c_lo), _i3(c_hi), _i3(grid_lo), _i3(grid_hi), _i3(rr), int(order),
                          _i3(fl), _i3(fh), out.ctypes.data_as(C.c_void_p), _i3(out_lo), _i3(out_hi))
    return out


def _pp(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


def arches_sigmaT4(gasVolFrac, absk, temp, do_part_radiation, dtype=np.float64, sigma=SIGMA):
    temp = [np.ascontiguousarray(t, np.float64) for t in temp]
    absk = [np.ascontiguousarray(a, np.float64) for a in (absk if absk is not None else temp)]
    gv = np.ascontiguousarray(gasVolFrac if gasVolFrac is not None else np.ones_like(temp[0]), np.float64)
    out = np.empty(temp[0].shape, dtype)
    lib().oracle_arches_sigmaT4(gv.ctypes.data_as(C.c_void_p), _pp(absk), _pp(temp), len(temp), 1 if do_part_radiation else 0,
                                out.ctypes.data_as(C.c_void_p), 1 if dtype == np.float32 else 0, C.c_int64(out.size), C.c_double(sigma))
    return out


def arches_sumAbsk(gasVolFrac, absk, lo, hi, int_lo, int_hi, dtype=np.float64):
    absk = [np.ascontiguousarray(a, np.float64) for a in absk]
    gv = np.ascontiguousarray(gasVolFrac, np.float64)
    out = np.empty(gv.shape, dtype)
    lib().oracle_arches_sumAbsk(gv.ctypes.data_as(C.c_void_p), _pp(absk), len(absk), _i3(lo), _i3(hi), _i3(int_lo), _i3(int_hi),
                                out.ctypes.data_as(C.c_void_p), 1 if dtype == np.float32 else 0)
    return out


def arches_stencilToDBLs(boundFlux):
    bf = np.ascontiguousarray(boundFlux, np.float64)
    outs = [np.empty(bf.shape[:-1], np.float64) for _ in range(6)]
    lib().oracle_arches_stencilToDBLs(bf.ctypes.data_as(C.c_void_p), C.c_int64(outs[0].size), *[o.ctypes.data_as(C.c_void_p) for o in outs])
    return dict(zip(("E", "W", "N", "S", "T", "B"), outs))


# ---------------------------------------------------------------- golden: Burns & Christon centre line
# parametricStudies/postProcessTools/RMCRT.m:22-35 (x = 0:1/40:1, symmetric about 0.5)
_BC_HALF = [3.07561, 2.94305, 2.81667, 2.69362, 2.57196, 2.45069, 2.32896, 2.20616, 2.08181, 1.95549, 1.82691,
            1.69575, 1.56180, 1.42485, 1.28473, 1.1413, .99443, .84405, .6901, .53255, .37165]


def burns_christon_exact(x_cc):
    """pchip interpolation of the exact centre line to cell centres (RMCRT.m:37-38)."""
    from scipy.interpolate import PchipInterpolator
    xs = np.linspace(0.0, 1.0, 41)
    ys = -np.array(_BC_HALF[::-1][:-1] + _BC_HALF)
    return PchipInterpolator(xs, ys)(np.asarray(x_cc))


def centreline_L2(divQ_level, N):
    """L2 = sqrt(sum((sim-exact)^2)/n) along the x centre line (RMCRT.m:387-413); divQ_level has the (N+2)^3 window."""
    mid = N // 2
    line = divQ_level[mid + 1, mid + 1, 1:N + 1]
    x = (np.arange(N) + 0.5) / N
    ex = burns_christon_exact(x)
    return float(np.sqrt(np.sum((line - ex) ** 2) / N)), line, ex
