"""ctypes binding of include/rmcrt_b200.h (the C ABI of librmcrt_b200.so).

One python function per exported symbol, same names, same argument meaning.  Errors are raised as
RmcrtError carrying rmcrt_last_error() -- the analogue of the C++ host rethrowing InternalError /
ProblemSetupException.  No compute happens in python and there is no fallback when the library is absent.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "librmcrt_b200.so")

RMCRT_MAX_LEVELS = 8

# every symbol include/rmcrt_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "rmcrt_abi_version", "rmcrt_struct_size", "rmcrt_default_params", "rmcrt_create", "rmcrt_set_params", "rmcrt_destroy",
    "rmcrt_last_error", "rmcrt_define_level", "rmcrt_upload_level", "rmcrt_pack_level_device",
    "rmcrt_finalize_level", "rmcrt_finalize_level_async", "rmcrt_level_records", "rmcrt_trace_divq_device", "rmcrt_trace_divq",
    "rmcrt_trace_dataonion_device", "rmcrt_trace_dataonion", "rmcrt_compute_extents", "rmcrt_reset_counters",
    "rmcrt_get_counters", "rmcrt_sigmaT4_device", "rmcrt_sigmaT4", "rmcrt_coarsen_level_device", "rmcrt_set_level_extra_cells",
    "rmcrt_pack_box_device", "rmcrt_assemble_level", "rmcrt_allgather_records", "rmcrt_measure_read_bandwidth", "rmcrt_debug_trace_rays",
    "rmcrt_radiometer_setup", "rmcrt_trace_radiometer_device", "rmcrt_trace_radiometer",
    "rmcrt_refine_Q_device", "rmcrt_refine_Q", "rmcrt_arches_sigmaT4_device", "rmcrt_arches_sumAbsk_device",
    "rmcrt_stencil_to_dbls_device", "rmcrt_arches_sigmaT4", "rmcrt_arches_sumAbsk", "rmcrt_stencil_to_dbls",
    "rmcrt_host_alloc", "rmcrt_host_free", "rmcrt_host_pool_trim", "rmcrt_replicate_level", "rmcrt_roi_extents",
]


class RmcrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rmcrt error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("nDivQRays", C.c_int32), ("nFluxRays", C.c_int32), ("threshold", C.c_double), ("sigma", C.c_double),
        ("allowReflect", C.c_int32), ("CCRays", C.c_int32), ("solveDivQ", C.c_int32),
        ("solveBoundaryFlux", C.c_int32), ("isSeedRandom", C.c_int32), ("flowCell", C.c_int32),
        ("haloCells", C.c_int32 * 3), ("cellTypeCoarsenLogic", C.c_int32), ("seed", C.c_uint64),
        ("marchPrecision", C.c_int32), ("rayCoherence", C.c_int32), ("skipZeroWeightLegs", C.c_int32),
        ("fixedRandom", C.c_int32), ("l2Persist", C.c_int32), ("warpsPerTile", C.c_int32),
        ("genericMarch", C.c_int32), ("maxBlocksPerSM", C.c_int32), ("rayDirSampleAlgo", C.c_int32), ("cellsPerLane", C.c_int32),
        ("maxRayLength", C.c_double), ("sigmaScat", C.c_double), ("tileX", C.c_int32), ("nonFlowCellType", C.c_int32),
        ("blockedMarch", C.c_int32), ("refillLanes", C.c_int32), ("coarsenExtraCells", C.c_int32), ("roiFixed", C.c_int32),
        ("roiFixedLo", C.c_int32 * 3), ("roiFixedHi", C.c_int32 * 3), ("originBrickHalo", C.c_int32), ("reserved2", C.c_int32),
    ]


class Radiometer(C.Structure):
    _fields_ = [
        ("cellLo", C.c_int32 * 3), ("cellHi", C.c_int32 * 3), ("nRays", C.c_int32), ("pad", C.c_int32),
        ("solidAngle", C.c_double), ("theta_viewAngle", C.c_double), ("range", C.c_double),
        ("phi_rotate", C.c_double), ("theta_rotate", C.c_double), ("xi_rotate", C.c_double),
    ]


class LevelDesc(C.Structure):
    _fields_ = [
        ("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3), ("interiorLo", C.c_int32 * 3), ("interiorHi", C.c_int32 * 3),
        ("dx", C.c_double * 3), ("anchor", C.c_double * 3), ("refineRatio", C.c_int32 * 3),
    ]


class Counters(C.Structure):
    _fields_ = [("nRaySteps", C.c_uint64), ("nFluxRaySteps", C.c_uint64), ("nDivQCells", C.c_uint64),
                ("nFluxCells", C.c_uint64), ("nRadiometerSteps", C.c_uint64)]


_lib = None


def load_library(path: str | None = None):
    """dlopen librmcrt_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("RMCRT_B200_LIB") or LIB_PATH     # RMCRT_B200_LIB: A/B builds of the same ABI
    if not os.path.exists(p):
        raise RmcrtError(-2, f"{p} not found: build it with `python -m uintah_b200.build` "
                             "(there is no CPU fallback for the RMCRT hot path)")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    i3 = C.POINTER(C.c_int32)
    vp = C.c_void_p
    lib.rmcrt_abi_version.restype = C.c_int
    lib.rmcrt_struct_size.argtypes = [C.c_int]
    lib.rmcrt_struct_size.restype = C.c_size_t
    for which, cls in ((0, Params), (1, LevelDesc), (2, Counters), (3, Radiometer)):
        if lib.rmcrt_struct_size(which) != C.sizeof(cls):
            raise RmcrtError(-1, f"ABI struct layout mismatch for {cls.__name__}: library {lib.rmcrt_struct_size(which)} B, binding {C.sizeof(cls)} B")
    lib.rmcrt_default_params.argtypes = [C.POINTER(Params)]
    lib.rmcrt_default_params.restype = None
    lib.rmcrt_create.argtypes = [C.c_int, C.POINTER(Params), C.POINTER(vp)]
    lib.rmcrt_set_params.argtypes = [vp, C.POINTER(Params)]
    lib.rmcrt_destroy.argtypes = [vp]
    lib.rmcrt_destroy.restype = None
    lib.rmcrt_last_error.argtypes = [vp]
    lib.rmcrt_last_error.restype = C.c_char_p
    lib.rmcrt_define_level.argtypes = [vp, C.c_int, C.POINTER(LevelDesc), C.c_int]
    lib.rmcrt_upload_level.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, i3, i3, vp]
    lib.rmcrt_pack_level_device.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, i3, i3, vp]
    lib.rmcrt_finalize_level.argtypes = [vp, C.c_int, vp]
    lib.rmcrt_finalize_level_async.argtypes = [vp, C.c_int, vp]
    lib.rmcrt_level_records.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    for name in ("rmcrt_trace_divq_device", "rmcrt_trace_divq", "rmcrt_trace_dataonion_device", "rmcrt_trace_dataonion"):
        getattr(lib, name).argtypes = [vp, C.c_int, i3, i3, vp, vp, vp, i3, i3, C.c_int, C.c_uint64, vp]
    lib.rmcrt_radiometer_setup.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_int, C.POINTER(Radiometer)]
    for name in ("rmcrt_trace_radiometer_device", "rmcrt_trace_radiometer"):
        getattr(lib, name).argtypes = [vp, C.c_int, C.POINTER(Radiometer), C.c_int, i3, i3, vp, vp, C.c_int, i3, i3, C.c_int,
                                       C.c_uint64, vp]
    lib.rmcrt_compute_extents.argtypes = [vp, C.c_int, i3, i3, i3, i3, i3, i3]
    lib.rmcrt_reset_counters.argtypes = [vp, vp]
    lib.rmcrt_get_counters.argtypes = [vp, C.POINTER(Counters), vp]
    lib.rmcrt_sigmaT4_device.argtypes = [vp, vp, vp, C.c_int, C.c_int64, vp]
    lib.rmcrt_sigmaT4.argtypes = [vp, vp, vp, C.c_int, C.c_int64, vp]
    lib.rmcrt_coarsen_level_device.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.rmcrt_set_level_extra_cells.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), i3, vp]
    lib.rmcrt_pack_box_device.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, i3, i3, i3, i3, vp, vp]
    lib.rmcrt_assemble_level.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, i3, i3, vp]
    lib.rmcrt_allgather_records.argtypes = [vp, vp, vp, C.c_size_t, vp, vp]
    lib.rmcrt_measure_read_bandwidth.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.rmcrt_debug_trace_rays.argtypes = [vp, C.c_int, C.c_int, i3, i3, C.c_int, i3, C.POINTER(C.c_double),
                                           C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint32), i3, C.c_int]
    lib.rmcrt_refine_Q_device.argtypes = [vp, vp, i3, i3, i3, i3, i3, C.c_int, i3, i3, vp, i3, i3, C.c_int, vp]
    lib.rmcrt_refine_Q.argtypes = [vp, vp, i3, i3, i3, i3, i3, C.c_int, i3, i3, vp, i3, i3, vp]
    pp = C.POINTER(vp)
    for name in ("rmcrt_arches_sigmaT4_device", "rmcrt_arches_sigmaT4"):
        getattr(lib, name).argtypes = [vp, vp, pp, pp, C.c_int, C.c_int, vp, C.c_int, C.c_int64, vp]
    for name in ("rmcrt_arches_sumAbsk_device", "rmcrt_arches_sumAbsk"):
        getattr(lib, name).argtypes = [vp, vp, pp, C.c_int, i3, i3, i3, i3, vp, C.c_int, vp]
    for name in ("rmcrt_stencil_to_dbls_device", "rmcrt_stencil_to_dbls"):
        getattr(lib, name).argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp]
    if path is None:
        _lib = lib
    return lib


def default_params(**overrides) -> Params:
    p = Params()
    load_library().rmcrt_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(f"rmcrt_params has no field {k!r}")
        if k in ("haloCells", "roiFixedLo", "roiFixedHi"):
            setattr(p, k, (C.c_int32 * 3)(*[int(x) for x in v]))
        else:
            setattr(p, k, v)
    return p


def i3(v):
    return (C.c_int32 * 3)(*[int(x) for x in v])


def check(ctx, rc):
    if rc != 0:
        msg = load_library().rmcrt_last_error(ctx)
        raise RmcrtError(rc, msg.decode() if msg else "")
