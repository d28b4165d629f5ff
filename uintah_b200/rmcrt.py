"""Python harness over the C ABI: what tests/ and bench.py drive.

It only moves pointers: numpy arrays for the host entry points, torch CUDA tensors (device memory, streams,
torch.distributed for the NCCL all-gather) for the device entry points.  All arithmetic happens inside
librmcrt_b200.so.  Mirrors the call sequence of the reference task bodies:

    Ray::rayTrace<T>            (Ray.cc:513-818)   -> RMCRT.upload_level / finalize_level / trace_divq
    Ray::rayTrace_dataOnion<T>  (Ray.cc:943-1274)  -> ... / trace_dataonion
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import Counters, LevelDesc, Params, Radiometer, RmcrtError, check, i3


@dataclass
class LevelData:
    """One level's property arrays, numpy [z][y][x] over index window [lo,hi) incl. extra cells."""
    abskg: np.ndarray
    sigmaT4: np.ndarray          # sigma T^4 / pi
    cellType: np.ndarray
    lo: tuple
    hi: tuple
    dx: tuple
    anchor: tuple = (0.0, 0.0, 0.0)
    rr: tuple = (1, 1, 1)        # refinement ratio to the next coarser level
    interior_lo: tuple = None
    interior_hi: tuple = None

    def __post_init__(self):
        if self.interior_lo is None:
            self.interior_lo = tuple(x + 1 for x in self.lo)
        if self.interior_hi is None:
            self.interior_hi = tuple(x - 1 for x in self.hi)

    @classmethod
    def of(cls, obj) -> "LevelData":
        """Accept any object with the same attribute names (e.g. the test oracle's Level)."""
        if isinstance(obj, cls):
            return obj
        return cls(obj.abskg, obj.sigmaT4, obj.cellType, tuple(obj.lo), tuple(obj.hi), tuple(obj.dx),
                   tuple(obj.anchor), tuple(obj.rr), tuple(obj.interior_lo), tuple(obj.interior_hi))

    def desc(self) -> LevelDesc:
        d = LevelDesc()
        d.lo = i3(self.lo); d.hi = i3(self.hi)
        d.interiorLo = i3(self.interior_lo); d.interiorHi = i3(self.interior_hi)
        d.dx = (C.c_double * 3)(*self.dx); d.anchor = (C.c_double * 3)(*self.anchor)
        d.refineRatio = i3(self.rr)
        return d


def _ptr(x):
    """host numpy array or torch tensor -> raw address"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()


def _stream_handle(stream):
    if stream is None:
        return None
    return getattr(stream, "cuda_stream", stream)


class _DevView:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class RMCRT:
    """Owner of one rmcrt_ctx (one per process / GPU, like the per-process Ray instance)."""

    def __init__(self, device: int = 0, params: Params | None = None, half_records: bool = False, **overrides):
        """half_records: levels are stored as 4-B half-packed records (recordKind 2) instead of 8-B float records"""
        self.lib = capi.load_library()
        self.half_records = bool(half_records)
        self.params = params if params is not None else capi.default_params(**overrides)
        if params is not None:
            for k, v in overrides.items():
                setattr(self.params, k, v)
        h = C.c_void_p()
        rc = self.lib.rmcrt_create(device, C.byref(self.params), C.byref(h))
        if rc != 0:
            msg = self.lib.rmcrt_last_error(None)
            raise RmcrtError(rc, msg.decode() if msg else "")
        self.ctx = h
        self.device = device
        self.levels: dict[int, LevelData | LevelDesc] = {}

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.rmcrt_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise AttributeError(k)
            if k in ("haloCells", "roiFixedLo", "roiFixedHi"):
                setattr(self.params, k, i3(v))
            else:
                setattr(self.params, k, v)
        check(self.ctx, self.lib.rmcrt_set_params(self.ctx, C.byref(self.params)))

    @property
    def record_kind(self) -> int:
        return 1 if self.params.marchPrecision == 1 else (2 if self.half_records else 0)

    # ------------------------------------------------------------------ levels
    def define_level(self, L: int, desc: LevelDesc, record_kind: int | None = None):
        rk = self.record_kind if record_kind is None else record_kind
        check(self.ctx, self.lib.rmcrt_define_level(self.ctx, L, C.byref(desc), rk))

    def upload_level(self, L: int, level, stream=None, finalize=True):
        """Host arrays -> device records (H2D inside)."""
        lv = LevelData.of(level)
        self.levels[L] = lv
        self.define_level(L, lv.desc())
        ab = np.ascontiguousarray(lv.abskg)
        s4 = np.ascontiguousarray(lv.sigmaT4, dtype=ab.dtype)
        ct = np.ascontiguousarray(lv.cellType, dtype=np.int32)
        is_float = 1 if ab.dtype == np.float32 else 0
        check(self.ctx, self.lib.rmcrt_upload_level(self.ctx, L, ab.ctypes.data, s4.ctypes.data, ct.ctypes.data, is_float,
                                                    i3(lv.lo), i3(lv.hi), _stream_handle(stream)))
        if finalize:
            self.finalize_level(L, stream)

    def pack_level_device(self, L: int, abskg, sigmaT4, cellType, src_lo, src_hi, stream=None):
        """torch CUDA tensors (float32|float64, int32), x-fastest over [src_lo,src_hi)."""
        import torch
        assert abskg.is_cuda and sigmaT4.is_cuda and cellType.is_cuda and cellType.dtype == torch.int32
        assert abskg.dtype == sigmaT4.dtype and abskg.is_contiguous() and sigmaT4.is_contiguous() and cellType.is_contiguous()
        is_float = 1 if abskg.dtype == torch.float32 else 0
        check(self.ctx, self.lib.rmcrt_pack_level_device(self.ctx, L, abskg.data_ptr(), sigmaT4.data_ptr(), cellType.data_ptr(),
                                                         is_float, i3(src_lo), i3(src_hi), _stream_handle(stream)))

    def finalize_level(self, L: int, stream=None):
        check(self.ctx, self.lib.rmcrt_finalize_level(self.ctx, L, _stream_handle(stream)))

    def finalize_level_async(self, L: int, stream=None):
        check(self.ctx, self.lib.rmcrt_finalize_level_async(self.ctx, L, _stream_handle(stream)))

    def level_records(self, L: int):
        """torch uint8 view of the packed records of level L (for the all-gather)."""
        import torch
        p = C.c_void_p(); n = C.c_size_t(); rb = C.c_size_t()
        check(self.ctx, self.lib.rmcrt_level_records(self.ctx, L, C.byref(p), C.byref(n), C.byref(rb)))
        return torch.as_tensor(_DevView(p.value, n.value), device=f"cuda:{self.device}"), rb.value

    def set_level_extra_cells(self, L: int, abskg6, sigmaT4_6, type6, stream=None):
        a = (C.c_double * 6)(*[float(x) for x in abskg6])
        s = (C.c_double * 6)(*[float(x) for x in sigmaT4_6])
        t = (C.c_int32 * 6)(*[int(x) for x in type6])
        check(self.ctx, self.lib.rmcrt_set_level_extra_cells(self.ctx, L, a, s, t, _stream_handle(stream)))

    def coarsen_level_device(self, fineL: int, coarseL: int, stream=None):
        check(self.ctx, self.lib.rmcrt_coarsen_level_device(self.ctx, fineL, coarseL, _stream_handle(stream)))

    def sigmaT4_device(self, temperature, out, stream=None):
        import torch
        assert temperature.dtype == torch.float64 and temperature.is_cuda and out.is_cuda
        is_float = 1 if out.dtype == torch.float32 else 0
        check(self.ctx, self.lib.rmcrt_sigmaT4_device(self.ctx, temperature.data_ptr(), out.data_ptr(), is_float,
                                                      temperature.numel(), _stream_handle(stream)))

    # ------------------------------------------------------------------ multi-GPU replication
    def pack_box_device(self, L, abskg, sigmaT4, cellType, src_lo, src_hi, box_lo, box_hi, send, stream=None):
        import torch
        is_float = 1 if abskg.dtype == torch.float32 else 0
        check(self.ctx, self.lib.rmcrt_pack_box_device(self.ctx, L, abskg.data_ptr(), sigmaT4.data_ptr(), cellType.data_ptr(),
                                                       is_float, i3(src_lo), i3(src_hi), i3(box_lo), i3(box_hi),
                                                       send.data_ptr(), _stream_handle(stream)))

    def assemble_level(self, L, gathered, slot_bytes, boxes_lo, boxes_hi, stream=None):
        n = len(boxes_lo)
        lo = (C.c_int32 * (3 * n))(*[int(x) for b in boxes_lo for x in b])
        hi = (C.c_int32 * (3 * n))(*[int(x) for b in boxes_hi for x in b])
        check(self.ctx, self.lib.rmcrt_assemble_level(self.ctx, L, gathered.data_ptr(), slot_bytes, n, lo, hi,
                                                      _stream_handle(stream)))

    # ------------------------------------------------------------------ the hot path
    def _window(self, L, out_lo, out_hi):
        lv = self.levels.get(L)
        if out_lo is None:
            out_lo, out_hi = lv.lo, lv.hi
        return tuple(out_lo), tuple(out_hi)

    def _trace_host(self, fn, Larg, L, patch_lo, patch_hi, out_lo, out_hi, timestep, zero_init, outputs, stream):
        lv = self.levels[L]
        patch_lo = lv.interior_lo if patch_lo is None else tuple(patch_lo)
        patch_hi = lv.interior_hi if patch_hi is None else tuple(patch_hi)
        out_lo, out_hi = self._window(L, out_lo, out_hi)
        shape = (out_hi[2] - out_lo[2], out_hi[1] - out_lo[1], out_hi[0] - out_lo[0])
        if outputs is None:
            divQ = np.zeros(shape, np.float64)
            volq = np.zeros(shape, np.float64)
            bflux = np.zeros(shape + (7,), np.float64) if self.params.solveBoundaryFlux else None
        else:
            divQ, volq, bflux = outputs
        check(self.ctx, fn(self.ctx, Larg, i3(patch_lo), i3(patch_hi), divQ.ctypes.data, volq.ctypes.data,
                           None if bflux is None else bflux.ctypes.data, i3(out_lo), i3(out_hi), 1 if zero_init else 0,
                           C.c_uint64(timestep), _stream_handle(stream)))
        return dict(divQ=divQ, radiationVolq=volq, boundFlux=bflux)

    def trace_divq(self, L=0, patch_lo=None, patch_hi=None, out_lo=None, out_hi=None, timestep=1, zero_init=True,
                   outputs=None, stream=None):
        """Host-buffer solve of one patch (default: the whole interior into a level-sized window)."""
        return self._trace_host(self.lib.rmcrt_trace_divq, L, L, patch_lo, patch_hi, out_lo, out_hi, timestep, zero_init,
                                outputs, stream)

    def trace_dataonion(self, max_levels, patch_lo=None, patch_hi=None, out_lo=None, out_hi=None, timestep=1,
                        zero_init=True, outputs=None, stream=None):
        return self._trace_host(self.lib.rmcrt_trace_dataonion, max_levels, max_levels - 1, patch_lo, patch_hi, out_lo,
                                out_hi, timestep, zero_init, outputs, stream)

    def trace_divq_device(self, L, patch_lo, patch_hi, divQ, volq, bflux, out_lo, out_hi, timestep=1, zero_init=True,
                          stream=None):
        """zero_init: True / 1 = the whole window is zeroed first; 2 = only the part of the window outside the patch (the kernels write
        every cell of the patch: for output buffers in mapped page-locked host memory); False / 0 = modifies_divQ"""
        check(self.ctx, self.lib.rmcrt_trace_divq_device(self.ctx, L, i3(patch_lo), i3(patch_hi), divQ.data_ptr(),
                                                         volq.data_ptr(), None if bflux is None else bflux.data_ptr(),
                                                         i3(out_lo), i3(out_hi), int(zero_init),
                                                         C.c_uint64(timestep), _stream_handle(stream)))

    def trace_dataonion_device(self, max_levels, patch_lo, patch_hi, divQ, volq, bflux, out_lo, out_hi, timestep=1,
                               zero_init=True, stream=None):
        check(self.ctx, self.lib.rmcrt_trace_dataonion_device(self.ctx, max_levels, i3(patch_lo), i3(patch_hi), divQ.data_ptr(),
                                                              volq.data_ptr(), None if bflux is None else bflux.data_ptr(),
                                                              i3(out_lo), i3(out_hi), int(zero_init),
                                                              C.c_uint64(timestep), _stream_handle(stream)))

    # ------------------------------------------------------------------ virtual radiometer
    def radiometer(self, view_angle, vec, cell_lo, cell_hi, n_rays, is_normal_vector=False) -> Radiometer:
        """one <Radiometer> block (Radiometer::problemSetup): angles from the library, cell box and nRays from the caller"""
        r = Radiometer()
        v = (C.c_double * 3)(*[float(x) for x in vec])
        rc = self.lib.rmcrt_radiometer_setup(float(view_angle), v, 1 if is_normal_vector else 0, C.byref(r))
        if rc != 0:
            msg = self.lib.rmcrt_last_error(None)
            raise RmcrtError(rc, msg.decode() if msg else "")
        r.cellLo = i3(cell_lo); r.cellHi = i3(cell_hi); r.nRays = int(n_rays)
        return r

    def trace_radiometer(self, rads, L=0, patch_lo=None, patch_hi=None, out_lo=None, out_hi=None, dtype=np.float64, timestep=1,
                         zero_init=True, outputs=None, stream=None):
        """Radiometer::radiometerFlux for one patch, host buffers; returns (VRFlux, VRIntensity) over the output window"""
        lv = self.levels[L]
        patch_lo = lv.interior_lo if patch_lo is None else tuple(patch_lo)
        patch_hi = lv.interior_hi if patch_hi is None else tuple(patch_hi)
        out_lo, out_hi = self._window(L, out_lo, out_hi)
        shape = (out_hi[2] - out_lo[2], out_hi[1] - out_lo[1], out_hi[0] - out_lo[0])
        if outputs is None:
            flux = np.zeros(shape, dtype); inten = np.zeros(shape, dtype)
        else:
            flux, inten = outputs
        arr = (Radiometer * max(len(rads), 1))(*rads)
        check(self.ctx, self.lib.rmcrt_trace_radiometer(self.ctx, L, arr, len(rads), i3(patch_lo), i3(patch_hi), flux.ctypes.data,
                                                        inten.ctypes.data, 1 if flux.dtype == np.float32 else 0, i3(out_lo),
                                                        i3(out_hi), 1 if zero_init else 0, C.c_uint64(timestep),
                                                        _stream_handle(stream)))
        return flux, inten

    def compute_extents(self, max_levels, patch_lo, patch_hi):
        roiLo = (C.c_int32 * 3)(); roiHi = (C.c_int32 * 3)()
        regLo = (C.c_int32 * (3 * max_levels))(); regHi = (C.c_int32 * (3 * max_levels))()
        check(self.ctx, self.lib.rmcrt_compute_extents(self.ctx, max_levels, i3(patch_lo), i3(patch_hi), roiLo, roiHi, regLo, regHi))
        return (tuple(roiLo), tuple(roiHi), np.array(regLo).reshape(max_levels, 3), np.array(regHi).reshape(max_levels, 3))

    def roi_extents(self, L, patch_lo, patch_hi, abskg_thld, sigmaT4_thld, out_lo=None, out_hi=None, dtype=np.float32, stream=None):
        """Ray::ROI_Extents for one patch: (mag_grad_abskg, mag_grad_sigmaT4, flaggedCells, anyFlagged) over the output window"""
        out_lo, out_hi = self._window(L, out_lo, out_hi)
        shape = (out_hi[2] - out_lo[2], out_hi[1] - out_lo[1], out_hi[0] - out_lo[0])
        ga = np.zeros(shape, dtype); gs = np.zeros(shape, dtype); fl = np.zeros(shape, np.int32)
        flag = C.c_int32(0)
        self.lib.rmcrt_roi_extents.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_double, C.c_double, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]
        check(self.ctx, self.lib.rmcrt_roi_extents(self.ctx, L, i3(patch_lo), i3(patch_hi), float(abskg_thld), float(sigmaT4_thld), ga.ctypes.data,
                                                   gs.ctypes.data, fl.ctypes.data, 1 if dtype == np.float32 else 0, i3(out_lo), i3(out_hi),
                                                   C.byref(flag), _stream_handle(stream)))
        return ga, gs, fl, bool(flag.value)

    # ------------------------------------------------------------------ callers either side of the path
    def refine_Q(self, coarseQ, c_lo, c_hi, grid_lo, grid_hi, rr, order, fl, fh, out_lo, out_hi, stream=None):
        """Ray::refine_Q for one scalar of one fine patch, host buffers; returns the fine window (zero outside [fl,fh))"""
        q = np.ascontiguousarray(coarseQ, np.float64)
        assert q.shape == (c_hi[2] - c_lo[2], c_hi[1] - c_lo[1], c_hi[0] - c_lo[0])
        out = np.empty((out_hi[2] - out_lo[2], out_hi[1] - out_lo[1], out_hi[0] - out_lo[0]), np.float64)
        check(self.ctx, self.lib.rmcrt_refine_Q(self.ctx, q.ctypes.data, i3(c_lo), i3(c_hi), i3(grid_lo), i3(grid_hi), i3(rr), int(order),
                                                i3(fl), i3(fh), out.ctypes.data, i3(out_lo), i3(out_hi), _stream_handle(stream)))
        return out

    def refine_Q_device(self, coarseQ, c_lo, c_hi, grid_lo, grid_hi, rr, order, fl, fh, fineQ, out_lo, out_hi, zero_init=True,
                        stream=None):
        check(self.ctx, self.lib.rmcrt_refine_Q_device(self.ctx, coarseQ.data_ptr(), i3(c_lo), i3(c_hi), i3(grid_lo), i3(grid_hi),
                                                       i3(rr), int(order), i3(fl), i3(fh), fineQ.data_ptr(), i3(out_lo), i3(out_hi),
                                                       1 if zero_init else 0, _stream_handle(stream)))

    @staticmethod
    def _ptr_array(arrays):
        return (C.c_void_p * len(arrays))(*[_ptr(a) for a in arrays])

    def arches_sigmaT4(self, gasVolFrac, absk, temp, do_part_radiation, dtype=np.float64, stream=None):
        """RMCRT_Radiation::sigmaT4<T> over one patch (host numpy arrays of equal shape); absk/temp: lists, [0] = gas"""
        temp = [np.ascontiguousarray(t, np.float64) for t in temp]
        absk = [np.ascontiguousarray(a, np.float64) for a in absk] if absk is not None else None
        gv = np.ascontiguousarray(gasVolFrac, np.float64) if gasVolFrac is not None else None
        out = np.empty(temp[0].shape, dtype)
        check(self.ctx, self.lib.rmcrt_arches_sigmaT4(self.ctx, _ptr(gv), self._ptr_array(absk) if absk else None, self._ptr_array(temp),
                                                      len(temp), 1 if do_part_radiation else 0, out.ctypes.data,
                                                      1 if dtype == np.float32 else 0, out.size, _stream_handle(stream)))
        return out

    def arches_sumAbsk(self, gasVolFrac, absk, lo, hi, int_lo, int_hi, dtype=np.float64, stream=None):
        absk = [np.ascontiguousarray(a, np.float64) for a in absk]
        gv = np.ascontiguousarray(gasVolFrac, np.float64)
        out = np.empty(gv.shape, dtype)
        check(self.ctx, self.lib.rmcrt_arches_sumAbsk(self.ctx, gv.ctypes.data, self._ptr_array(absk), len(absk), i3(lo), i3(hi), i3(int_lo),
                                                      i3(int_hi), out.ctypes.data, 1 if dtype == np.float32 else 0, _stream_handle(stream)))
        return out

    def stencil_to_dbls(self, boundFlux, stream=None):
        """RMCRT_Radiation::stencilToDBLs: Stencil7 AoS [...,7] -> dict of the six radiationFlux fields"""
        bf = np.ascontiguousarray(boundFlux, np.float64)
        assert bf.shape[-1] == 7
        outs = [np.empty(bf.shape[:-1], np.float64) for _ in range(6)]
        check(self.ctx, self.lib.rmcrt_stencil_to_dbls(self.ctx, bf.ctypes.data, outs[0].size, *[o.ctypes.data for o in outs],
                                                       _stream_handle(stream)))
        return dict(zip(("E", "W", "N", "S", "T", "B"), outs))

    # ------------------------------------------------------------------ counters / debug
    def reset_counters(self, stream=None):
        check(self.ctx, self.lib.rmcrt_reset_counters(self.ctx, _stream_handle(stream)))

    def counters(self, stream=None) -> Counters:
        c = Counters()
        check(self.ctx, self.lib.rmcrt_get_counters(self.ctx, C.byref(c), _stream_handle(stream)))
        return c

    def measure_read_bandwidth(self, nbytes, repeats=20, pattern=0):
        """(GB/s, Gaccess/s) of the roofline microbenchmark (pattern 0 streaming, 1 per-lane 32-B sector gathers)"""
        g = C.c_double(); a = C.c_double()
        check(self.ctx, self.lib.rmcrt_measure_read_bandwidth(self.ctx, nbytes, repeats, pattern, C.byref(g), C.byref(a)))
        return g.value, a.value

    def debug_trace_rays(self, origin_cells, dirs, ray_origins, L=0, max_levels=1, max_path=0):
        cells = np.ascontiguousarray(origin_cells, np.int32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        o = np.ascontiguousarray(ray_origins, np.float64).reshape(-1, 3)
        n = cells.shape[0]
        sumI = np.zeros(n, np.float64)
        steps = np.zeros(n, np.uint32)
        path = np.zeros((n, max(max_path, 1), 4), np.int32)
        P = C.POINTER
        check(self.ctx, self.lib.rmcrt_debug_trace_rays(
            self.ctx, L, max_levels, None, None, n, cells.ctypes.data_as(P(C.c_int32)), d.ctypes.data_as(P(C.c_double)),
            o.ctypes.data_as(P(C.c_double)), sumI.ctypes.data_as(P(C.c_double)), steps.ctypes.data_as(P(C.c_uint32)),
            path.ctypes.data_as(P(C.c_int32)) if max_path else None, max_path))
        return sumI, steps, path
