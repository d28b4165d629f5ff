"""Two or more GPUs in one box (skipped on a single-GPU box): replication of a level over the devices of one process with ONE
ncclAllGather (rmcrt_replicate_level, the C++ host path), and the raw-communicator entry point rmcrt_allgather_records."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_replicate_level_equals_per_device_uploads(gpu_lib, po, dtype):
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    from uintah_b200 import RMCRT, LevelData
    from uintah_b200.capi import check, i3
    n = min(n_gpus(), 4)
    rng = np.random.default_rng(3)
    lv = po.burns_christon(37, dtype=dtype)
    lv.abskg[1:-1, 1:-1, 1:-1] *= rng.uniform(0.5, 1.5, lv.abskg[1:-1, 1:-1, 1:-1].shape).astype(dtype)
    prec = 1 if dtype == np.float64 else 0
    gs = [RMCRT(device=d, marchPrecision=prec, nDivQRays=4, threshold=0.05, isSeedRandom=0, skipZeroWeightLegs=0) for d in range(n)]
    for g in gs:
        g.define_level(0, LevelData.of(lv).desc())
        g.levels[0] = LevelData.of(lv)
    ctxs = (C.c_void_p * n)(*[g.ctx for g in gs])
    ab = np.ascontiguousarray(lv.abskg); s4 = np.ascontiguousarray(lv.sigmaT4); ct = np.ascontiguousarray(lv.cellType, np.int32)
    lib = gs[0].lib
    lib.rmcrt_replicate_level.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    check(gs[0].ctx, lib.rmcrt_replicate_level(ctxs, n, 0, ab.ctypes.data, s4.ctypes.data, ct.ctypes.data, 1 if dtype == np.float32 else 0,
                                               i3(lv.lo), i3(lv.hi)))
    ref = RMCRT(device=0, marchPrecision=prec)
    ref.upload_level(0, lv)
    want = ref.level_records(0)[0].cpu().numpy()
    for g in gs:
        g.finalize_level(0)
        got = g.level_records(0)[0].cpu().numpy()
        assert np.array_equal(got, want)
    # and a solve on the last device equals the single-device solve bit for bit (deterministic rays)
    a = gs[-1].trace_divq(0)["divQ"]
    ref.set_params(nDivQRays=4, threshold=0.05, isSeedRandom=0, skipZeroWeightLegs=0)
    b = ref.trace_divq(0)["divQ"]
    assert np.array_equal(a, b)


def test_cpp_host_on_two_devices(gpu_lib, tmp_path):
    """rmcrt_test --devices 2: the 8 patches are dealt round robin over two contexts, the level is replicated through NCCL; with the
    default sampling (keyed on the cell) divQ equals the one-device run bit for bit"""
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    import json
    exe = os.path.join(ROOT, "uintah_b200", "host", "rmcrt_test")
    deck = os.path.join(ROOT, "tests", "decks", "bm1_1L_float_8patches.ups")
    outs = []
    for nd in (1, 2):
        d = tmp_path / f"dev{nd}"
        r = subprocess.run([exe, deck, "--dump", str(d), "--devices", str(nd)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert f"on {nd} GPU(s)" in r.stdout
        meta = json.load(open(d / "meta.json"))
        m = meta["divQ"]
        outs.append(np.fromfile(d / m["file"], dtype=np.dtype("<" + m["dtype"])))
    assert np.array_equal(outs[0], outs[1]) and np.abs(outs[0]).max() > 0
