"""C-ABI surface checks that need no GPU: the library builds for sm_100a, loads, exports every symbol the header
declares, reports the reference's .ups defaults, and refuses to work (loudly) without a CUDA device."""
import ctypes as C
import math
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from uintah_b200 import build, capi
    build.build()
    return capi.load_library()


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "rmcrt_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rmcrt_[a-z0-9_A-Z]+)\s*\(", txt)))


def test_header_and_binding_agree(lib):
    from uintah_b200 import capi
    assert _header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for s in _header_symbols():
        assert hasattr(lib, s), s
    assert lib.rmcrt_abi_version() == 5


def test_library_is_sm100a_and_has_no_torch_dependency():
    from uintah_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
    ldd = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "python" not in ldd


def test_default_params_are_the_ups_defaults(lib):
    """Ray::problemSetup getWithDefault values (Ray.cc:162-173) and haloCells [10,10,10] (Ray.cc:254)"""
    from uintah_b200 import capi
    p = capi.default_params()
    assert (p.nDivQRays, p.nFluxRays, p.threshold, p.sigma) == (10, 1, 0.01, 5.67051e-8)
    assert (p.allowReflect, p.CCRays, p.solveDivQ, p.solveBoundaryFlux, p.isSeedRandom) == (1, 0, 1, 0, 1)
    assert p.flowCell == -1 and list(p.haloCells) == [10, 10, 10] and p.cellTypeCoarsenLogic == 0
    assert lib.rmcrt_struct_size(0) == C.sizeof(capi.Params)
    assert lib.rmcrt_struct_size(1) == C.sizeof(capi.LevelDesc)
    assert lib.rmcrt_struct_size(2) == C.sizeof(capi.Counters)
    assert lib.rmcrt_struct_size(3) == C.sizeof(capi.Radiometer)
    assert p.rayDirSampleAlgo == 0 and p.maxRayLength == 0.0      # "naive" (Ray.cc:174), unbounded (RMCRTCommon.cc:117)
    # implementation knobs: the reference's sampling (independent rays per cell) on the length-sorted blocked kernel
    assert p.rayCoherence == 0 and p.blockedMarch == 3


def test_radiometer_setup_is_host_only(lib):
    """Radiometer::problemSetup's angle derivation (Radiometer.cc:214-260) needs no device"""
    from uintah_b200 import capi
    r = capi.Radiometer()
    v = (C.c_double * 3)(0.0, 0.0, 1.0)
    assert lib.rmcrt_radiometer_setup(30.0, v, 0, C.byref(r)) == 0
    assert r.theta_viewAngle == pytest.approx(math.radians(15.0))
    assert r.solidAngle == pytest.approx(2 * math.pi * (1 - math.cos(math.radians(15.0))))
    assert r.theta_rotate == 0.0 and r.phi_rotate == 0.0
    assert r.xi_rotate == pytest.approx(math.pi / 4)      # the reference's 1e-16 zero guard: acos(1e-16 / (sqrt(2) 1e-16))
    assert lib.rmcrt_radiometer_setup(361.0, v, 0, C.byref(r)) != 0


def test_no_cpu_fallback(lib):
    """without a device the product path fails loudly; it never routes to a CPU implementation"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from uintah_b200 import RMCRT, RmcrtError
    with pytest.raises(RmcrtError) as e:
        RMCRT(device=0)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/"""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "uintah_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                src = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+[\"<].*oracle", src, flags=re.M):
                    bad.append(f)
    assert not bad, bad
