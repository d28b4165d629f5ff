"""uintah_b200 -- B200-native RMCRT divQ / boundary-flux solve behind Uintah's `Ray` plugin surface.

Layout
  csrc/   hand-written CUDA (sm_100a) kernels + the C ABI of include/rmcrt_b200.h  -> lib/librmcrt_b200.so
  host/   C++ mirror of the reference plugin surface (Ray, RMCRTCommon, RMCRT_Test driver) over the C ABI
  capi.py ctypes binding of the C ABI (what a cgo/JNI-style stub would be for a C++ host: see INTEGRATION.md)
  rmcrt.py thin python harness used by tests/ and bench.py (device memory + torch.distributed plumbing)

There is no CPU fallback: importing works anywhere, computing requires the CUDA library and a GPU.
"""
from .capi import RmcrtError, Params, LevelDesc, Counters, load_library  # noqa: F401
from .rmcrt import RMCRT, LevelData  # noqa: F401

__all__ = ["RMCRT", "LevelData", "Params", "LevelDesc", "Counters", "RmcrtError", "load_library"]
