"""Multi-GPU orchestration: one process per GPU, torch.distributed (NCCL on GPUs, gloo in
the CPU tests) for the only three exchanges the path has:

  1. the earliest-TOI all-reduce(min) -- 8 bytes;
  2. an all-gather of per-rank pair counts -- 8 bytes per rank;
  3. an order-preserving all-to-all that evens out the candidate pairs before the narrow
     phase (the sweep shards OWNERS by window work, which does not equalise pair counts).

Broad phase: every rank builds and sorts all boxes (replica; the reference's dead
_multigpu code did the same, _multigpu/broad_phase.cu:113-116) and sweeps only its own
owner slice (sccd_set_shard), so the pair lists are disjoint and their concatenation in rank
order IS the single-GPU deterministic list.  No collective is needed for the sweep itself.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------ pure host logic
def balance_plan(counts: Sequence[int], rank: int) -> Tuple[List[int], List[int]]:
    """Order-preserving even redistribution of a distributed list.

    counts[r] = items rank r holds (global order = rank order).  After the exchange rank r
    holds the global items [r*T/G, (r+1)*T/G).  Returns (send_splits, recv_splits) for
    all_to_all_single on `rank`."""
    world = len(counts)
    total = int(sum(counts))
    have_lo = [0] * (world + 1)
    for r in range(world):
        have_lo[r + 1] = have_lo[r] + int(counts[r])
    want_lo = [(total * r) // world for r in range(world + 1)]

    def overlap(a0, a1, b0, b1):
        return max(0, min(a1, b1) - max(a0, b0))

    send = [overlap(have_lo[rank], have_lo[rank + 1], want_lo[d], want_lo[d + 1])
            for d in range(world)]
    recv = [overlap(have_lo[s], have_lo[s + 1], want_lo[rank], want_lo[rank + 1])
            for s in range(world)]
    return send, recv


def rebalance(items, group=None):
    """Even out a distributed (n_r, k) tensor across the ranks, preserving global order.
    Works on CPU tensors (gloo) and CUDA tensors (NCCL)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([items.shape[0]], dtype=torch.int64, device=items.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    send, recv = balance_plan(counts, rank)
    out = torch.empty((sum(recv),) + tuple(items.shape[1:]), dtype=items.dtype, device=items.device)
    dist.all_to_all_single(out, items.contiguous(), recv, send, group=group)
    return out, counts


def allreduce_min(value: float, device, group=None) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return float(t.item())


class _DevArray:
    """Zero-copy torch view of a raw device pointer (CUDA array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


# ------------------------------------------------------------------ the sharded pipeline
class ShardedCCD:
    """ccd() over the GPUs of one box.  `ctx` is this rank's Context with the mesh uploaded."""

    def __init__(self, ctx, group=None, rebalance_pairs: bool = True):
        import torch.distributed as dist
        self.ctx = ctx
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rebalance_pairs = rebalance_pairs and self.world > 1
        self.last = {}

    def ccd(self, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True) -> float:
        import torch
        ctx = self.ctx
        dev = torch.device("cuda", ctx.device)
        ctx.set_shard(self.rank, self.world)
        ctx.build_boxes(ms)
        toi = 1.0
        info = {"pairs_local": [], "pairs_after": []}
        for kind in (0, 1):
            ctx.broad_phase_begin(kind)
            parts = []
            while not ctx.broad_phase_is_complete():
                ptr, n = ctx.broad_phase_partial()
                if n:
                    view = torch.as_tensor(_DevArray(ptr, (n, 2), "<i4"), device=dev)
                    parts.append(view.clone() if self.rebalance_pairs else view)
                    if not self.rebalance_pairs:
                        toi = ctx.narrow_phase(kind, ptr, n, ms, max_iter, tol, allow_zero_toi, toi)
            n_local = sum(int(p.shape[0]) for p in parts)
            info["pairs_local"].append(n_local)
            if self.rebalance_pairs:
                mine = torch.cat(parts) if parts else torch.empty((0, 2), dtype=torch.int32, device=dev)
                mine, _ = rebalance(mine, self.group)
                info["pairs_after"].append(int(mine.shape[0]))
                if mine.shape[0]:
                    toi = ctx.narrow_phase(kind, mine.data_ptr(), int(mine.shape[0]), ms, max_iter,
                                           tol, allow_zero_toi, toi)
            if self.world > 1:   # the next pass prunes with the global bound
                toi = allreduce_min(toi, dev, self.group)
        self.last = info
        ctx.set_shard(0, 1)
        return toi
