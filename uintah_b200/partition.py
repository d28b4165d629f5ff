"""Patch -> rank bookkeeping of the multi-GPU solve: which origin cells a rank traces, which property cells
it owns and packs for the all-gather, and how big an all-gather slot must be.  Pure index arithmetic (no device code),
shared by bench.py and the world_size-2 gloo test.

Mirrors how Uintah lays a level out: `<patches> [px,py,pz]` equal boxes, x fastest (Examples/RMCRT_Test decks), each patch
owning the extra cells on the domain faces it touches (Patch::getExtraCellLowIndex / HighIndex)."""
from __future__ import annotations

import numpy as np


def patch_grid(n_ranks: int):
    """ranks -> (px,py,pz): octants at 8 (the load-symmetric split of SURVEY 8e), slabs of octants below"""
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(n_ranks, (n_ranks, 1, 1))


def rank_patch(rank: int, pg, patch_cells):
    """interior cell box [lo,hi) of the patch traced by `rank` (x-fastest rank order)"""
    P = (patch_cells,) * 3 if np.isscalar(patch_cells) else tuple(patch_cells)
    idx = (rank % pg[0], (rank // pg[0]) % pg[1], rank // (pg[0] * pg[1]))
    lo = tuple(idx[d] * P[d] for d in range(3))
    return lo, tuple(lo[d] + P[d] for d in range(3))


def owner_box(patch_lo, patch_hi, level_shape, extra=1):
    """the patch plus the extra cells of the domain faces it touches: the property cells this rank packs"""
    lo = tuple(patch_lo[d] - (extra if patch_lo[d] == 0 else 0) for d in range(3))
    hi = tuple(patch_hi[d] + (extra if patch_hi[d] == level_shape[d] else 0) for d in range(3))
    return lo, hi


def all_owner_boxes(n_ranks: int, pg, patch_cells, level_shape, extra=1):
    los, his = [], []
    for r in range(n_ranks):
        lo, hi = owner_box(*rank_patch(r, pg, patch_cells), level_shape, extra)
        los.append(lo)
        his.append(hi)
    return los, his


def box_cells(lo, hi) -> int:
    return int(np.prod([hi[d] - lo[d] for d in range(3)]))


def slot_bytes(boxes_lo, boxes_hi, record_bytes=8) -> int:
    """all-gather slots are equal sized (NCCL all-gather): the largest owner box, in records"""
    return max(box_cells(l, h) for l, h in zip(boxes_lo, boxes_hi)) * record_bytes
