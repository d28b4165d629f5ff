"""world_size-2 (and 8-way, single process) checks of the multi-GPU bookkeeping on CPU: patch -> rank map, owner boxes,
all-gather slot layout over a real `gloo` process group, and re-assembly of the replicated level.
The device kernels that do the packing / assembly are covered on the GPU by test_replication_pack_gather_assemble."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from uintah_b200 import partition  # noqa: E402


def numpy_records(ab, s4, ct):
    """test-side restatement of the packed record: {abskg with sign bit = not a flow cell, sigmaT4/pi} as 8 bytes"""
    k = np.abs(ab).astype(np.float32)
    k = np.where(ct != -1, -k, k).astype(np.float32)
    return np.stack([k, s4.astype(np.float32)], axis=-1)


@pytest.mark.parametrize("n_ranks", [1, 2, 4, 8])
def test_owner_boxes_tile_the_level_exactly_once(n_ranks):
    P = 6
    pg = partition.patch_grid(n_ranks)
    assert int(np.prod(pg)) == n_ranks
    shape = tuple(P * g for g in pg)
    cover = np.zeros(tuple(s + 2 for s in shape[::-1]), np.int32)
    traced = np.zeros_like(cover)
    los, his = partition.all_owner_boxes(n_ranks, pg, P, shape)
    for r in range(n_ranks):
        lo, hi = los[r], his[r]
        cover[lo[2] + 1:hi[2] + 1, lo[1] + 1:hi[1] + 1, lo[0] + 1:hi[0] + 1] += 1
        plo, phi = partition.rank_patch(r, pg, P)
        traced[plo[2] + 1:phi[2] + 1, plo[1] + 1:phi[1] + 1, plo[0] + 1:phi[0] + 1] += 1
        assert partition.box_cells(lo, hi) * 8 <= partition.slot_bytes(los, his)
    assert (cover == 1).all()                              # every cell incl. extra cells is packed by exactly one rank
    assert (traced[1:-1, 1:-1, 1:-1] == 1).all() and traced.sum() == int(np.prod(shape))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from bench import burns_christon_fields
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = 8
        pg = partition.patch_grid(world)
        shape = tuple(P * g for g in pg)
        plo, phi = partition.rank_patch(rank, pg, P)
        blo, bhi = partition.owner_box(plo, phi, shape)
        los, his = partition.all_owner_boxes(world, pg, P, shape)
        slot = partition.slot_bytes(los, his)
        ab, s4, ct = burns_christon_fields(shape, blo, bhi, np.float32)
        rec = numpy_records(ab, s4, ct).tobytes()            # x-fastest over the owner box, like rmcrt_pack_box_device
        send = torch.zeros(slot, dtype=torch.uint8)
        send[:len(rec)] = torch.frombuffer(bytearray(rec), dtype=torch.uint8)
        gathered = torch.zeros(slot * world, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, send)
        # what rmcrt_assemble_level does: scatter every rank's box into the level
        lvl = np.full((shape[2] + 2, shape[1] + 2, shape[0] + 2, 2), np.nan, np.float32)
        for r in range(world):
            lo, hi = los[r], his[r]
            n = partition.box_cells(lo, hi)
            box = np.frombuffer(gathered[r * slot:r * slot + n * 8].numpy().tobytes(), np.float32).reshape(
                hi[2] - lo[2], hi[1] - lo[1], hi[0] - lo[0], 2)
            lvl[lo[2] + 1:hi[2] + 1, lo[1] + 1:hi[1] + 1, lo[0] + 1:hi[0] + 1] = box
        fab, fs4, fct = burns_christon_fields(shape, (-1, -1, -1), tuple(s + 1 for s in shape), np.float32)
        whole = numpy_records(fab, fs4, fct)
        ok = bool(np.array_equal(lvl.view(np.uint32), whole.view(np.uint32)))
        # the weak-scaling reduction bench.py does: max over ranks of time, sum of work
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        w = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        q.put((rank, ok, float(t), float(w)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allgather_reassembles_the_level():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert all(r[2] == 2.0 and r[3] == 30.0 for r in res)
