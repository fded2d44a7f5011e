"""Haplotype sharding across the GPUs of one box.

Every haplotype (Base, Smooth) and every haplotype PAIR (Gnofix; rows 2i, 2i+1 are
individual i -- src/model.py:205, src/utils.py:123) is independent, so the hot path needs
no collective: rank r owns a contiguous block of whole individuals and runs the same
kernels on it.  torch.distributed (NCCL on the GPUs, gloo in the CPU tests) is used only
around the hot path: scatter of the int8 SNP block and gather of the [N, W] labels.
"""
from __future__ import annotations

import numpy as np


def partition(n_rows: int, world: int):
    """[(row_lo, row_hi)] per rank: contiguous blocks of whole individuals (even row counts,
    a trailing odd haplotype stays with the last non-empty block), sizes differing by at most one
    individual."""
    n_ind = (n_rows + 1) // 2
    per, extra = divmod(n_ind, world)
    out, lo = [], 0
    for r in range(world):
        k = per + (1 if r < extra else 0)
        hi = min(n_rows, lo + 2 * k)
        out.append((lo, hi))
        lo = hi
    return out


def local_rows(n_rows: int, world: int, rank: int):
    return partition(n_rows, world)[rank]


def scatter_rows(X_root, n_rows: int, n_cols: int, dtype, device, group=None, src: int = 0):
    """Rank `src` holds X_root [n_rows, n_cols]; every rank gets its block (a tensor on `device`)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition(n_rows, world)
    lo, hi = parts[rank]
    mine = torch.empty((hi - lo, n_cols), dtype=dtype, device=device)
    if rank == src:
        reqs = []
        for r, (a, b) in enumerate(parts):
            if r == src:
                mine.copy_(X_root[a:b])
            elif b > a:
                reqs.append(dist.isend(X_root[a:b].contiguous(), dst=r, group=group))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(mine, src=src, group=group)
    return mine


def gather_rows(local, n_rows: int, group=None, dst: int = 0):
    """Inverse of scatter_rows for per-row outputs (labels [n, W], proba [n, W, A]):
    rank `dst` returns the concatenated [n_rows, ...] tensor, the others None."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = partition(n_rows, world)
    if rank == dst:
        out = torch.empty((n_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        for r, (a, b) in enumerate(parts):
            if r == dst:
                out[a:b].copy_(local)
            elif b > a:
                dist.recv(out[a:b], src=r, group=group)
        return out
    if local.shape[0] > 0:
        dist.send(local.contiguous(), dst=dst, group=group)
    return None
