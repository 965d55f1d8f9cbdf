"""Multi-GPU orchestration: one process per GPU, torch.distributed (NCCL on GPUs, gloo in
the CPU tests) for the only exchanges the path has:

  1. the earliest-TOI all-reduce(min) at the end of the step -- 8 bytes;
  2. an all-gather of per-rank pair counts -- 8 bytes per rank;
  3. an order-preserving all-to-all that evens out the candidate pairs before the narrow
     phase (the sweep is sharded by estimated sweep work, which does not equalise pairs);
  4. (host-buffer entry only) an all-gather of the mesh: every rank copies 1/world of the
     host buffers over its own PCIe link and the ranks exchange the slices over NVLink,
     instead of every rank pulling the whole mesh through PCIe.

Broad phase: every rank builds the (cheap, 64 B/box) exact boxes of the whole mesh, derives
the same (y, z) cell grid and the same `world` contiguous cell ranges of ~equal sweep work from
them, and then makes, radix-sorts and sweeps ONLY the records of its own cell range
(sccd_set_shard; csrc/grid.cu).  Cells are independent sweep domains -- a pair is reported in
its home cell only -- so the ranks' pair lists are disjoint and their concatenation in rank
order IS the single-GPU deterministic list: no halo and no collective for the sweep.  (Lists
with too few cells fall back to owner slices of a replicated sorted list, which is what the
reference's dead _multigpu code did, _multigpu/broad_phase.cu:113-116.)
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
    if items.is_cuda:   # one collective, one host sync
        allc = torch.empty(world, dtype=torch.int64, device=items.device)
        dist.all_gather_into_tensor(allc, n, group=group)
        counts = [int(c) for c in allc.tolist()]
    else:
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


def gather_min_and_loads(value: float, loads, device, group=None):
    """The step's single collective: every rank contributes (toi, its per-list query counts).
    Returns (min toi, [per-rank load lists]); the loads steer the NEXT step's decision to
    rebalance pairs (frame-to-frame coherence), so no extra collective sits inside a step."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor([value] + [float(x) for x in loads], dtype=torch.float64, device=device)
    out = torch.empty(world * mine.numel(), dtype=torch.float64, device=device)
    if out.is_cuda:
        dist.all_gather_into_tensor(out, mine, group=group)
    else:
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        out = torch.cat(parts)
    rows = out.view(world, -1).tolist()
    return min(r[0] for r in rows), [[int(x) for x in r[1:]] for r in rows]


def imbalance(loads) -> float:
    """max / mean of the ranks' total loads (1.0 = perfectly even)."""
    tot = [sum(r) for r in loads]
    mean = sum(tot) / max(len(tot), 1)
    return max(tot) / mean if mean > 0 else 1.0


def pack_mesh(V0, V1, E, F, world: int, pin: bool = True):
    """[V0 | V1 | E | F] (column-major, as the C ABI takes them) in ONE flat byte tensor whose
    length is a multiple of 16 * world, plus the byte offset of each array."""
    import torch
    arrs = [np.asfortranarray(V0, dtype=np.float64), np.asfortranarray(V1, dtype=np.float64),
            np.asfortranarray(E, dtype=np.int32), np.asfortranarray(F, dtype=np.int32)]
    offs, total = [], 0
    for a in arrs:
        offs.append(total)
        total += (a.nbytes + 15) // 16 * 16
    unit = 16 * world
    total = (total + unit - 1) // unit * unit
    flat = torch.zeros(max(total, unit), dtype=torch.uint8)
    if pin and torch.cuda.is_available():
        flat = flat.pin_memory()
    view = flat.numpy()
    for a, o in zip(arrs, offs):
        view[o:o + a.nbytes] = np.frombuffer(a.tobytes(order="F"), dtype=np.uint8)
    return flat, offs


def gather_mesh(flat_host, out, group=None):
    """Every rank moves ITS 1/world slice of the packed host mesh to `out` (device tensor of
    the same length; a CPU tensor under gloo) and the slices are all-gathered in place of a
    full per-rank host->device copy."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = flat_host.numel()
    assert n % world == 0 and out.numel() == n
    chunk = n // world
    mine = out[rank * chunk:(rank + 1) * chunk]
    mine.copy_(flat_host[rank * chunk:(rank + 1) * chunk], non_blocking=True)
    if out.is_cuda:
        dist.all_gather_into_tensor(out, mine, group=group)   # in place (NCCL semantics)
    else:
        parts = [out[r * chunk:(r + 1) * chunk] for r in range(world)]
        got = [p.clone() for p in parts]
        dist.all_gather(got, mine.clone(), group=group)
        for p, g in zip(parts, got):
            p.copy_(g)
    return out


class _DevArray:
    """Zero-copy torch view of a raw device pointer (CUDA array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


# ------------------------------------------------------------------ the sharded pipeline
class ShardedCCD:
    """ccd() over the GPUs of one box.  `ctx` is this rank's Context with the mesh uploaded."""

    # pairs are redistributed before the narrow phase only when the previous step left one
    # rank with more than this many times the mean number of queries
    REBALANCE_ABOVE = 1.25

    def __init__(self, ctx, group=None, rebalance_pairs="auto"):
        import torch.distributed as dist
        self.ctx = ctx
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # True / False, or "auto": decided from the load imbalance of the previous step.  The
        # cell ranges are balanced by records, which usually balances the pairs well enough
        # (config 4: within 8 %), and every mid-step collective is also a barrier that adds the
        # ranks' sweep skew to the step.
        self.mode = rebalance_pairs
        self.rebalance_pairs = (rebalance_pairs is True) and self.world > 1
        self.last = {}
        self.profile = False      # per-stage device times of the last ccd() in self.last["ms"]
        self._mesh = None

    def upload_mesh_host(self, flat_host, offs, sizes):
        """Host-buffer entry: pack_mesh() output -> sliced H2D + NVLink all-gather -> the
        context's mesh (device pointers)."""
        import torch
        dev = torch.device("cuda", self.ctx.device)
        if self._mesh is None or self._mesh.numel() != flat_host.numel():
            self._mesh = torch.empty(flat_host.numel(), dtype=torch.uint8, device=dev)
        if self.world > 1:
            gather_mesh(flat_host, self._mesh, self.group)
        else:
            self._mesh.copy_(flat_host, non_blocking=True)
        base = self._mesh.data_ptr()
        self.ctx.upload_mesh(base + offs[0], base + offs[1], base + offs[2], base + offs[3],
                             sizes=sizes)

    def ccd(self, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True) -> float:
        import torch
        ctx = self.ctx
        dev = torch.device("cuda", ctx.device)
        marks = []

        def mark(name):
            if self.profile:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        ctx.set_shard(self.rank, self.world)
        ctx.build_boxes(ms)
        mark("build+sort")
        toi = 1.0
        info = {"pairs_local": [], "pairs_after": []}
        for kind in (0, 1):
            tag = "vf" if kind == 0 else "ee"
            ctx.broad_phase_begin(kind)
            parts = []
            while not ctx.broad_phase_is_complete():
                ptr, n = ctx.broad_phase_partial()
                if n:
                    view = torch.as_tensor(_DevArray(ptr, (n, 2), "<i4"), device=dev)
                    if self.rebalance_pairs and not ctx.broad_phase_is_complete():
                        view = view.clone()   # the context reuses its pair buffer per chunk
                    parts.append(view)
                    if not self.rebalance_pairs:
                        toi = ctx.narrow_phase(kind, ptr, n, ms, max_iter, tol, allow_zero_toi, toi)
            n_local = sum(int(p.shape[0]) for p in parts)
            info["pairs_local"].append(n_local)
            mark("sweep_" + tag)
            if self.rebalance_pairs:
                mine = (parts[0] if len(parts) == 1 else torch.cat(parts)) if parts else \
                    torch.empty((0, 2), dtype=torch.int32, device=dev)
                mine, _ = rebalance(mine, self.group)
                info["pairs_after"].append(int(mine.shape[0]))
                mark("rebalance_" + tag)
                if mine.shape[0]:
                    toi = ctx.narrow_phase(kind, mine.data_ptr(), int(mine.shape[0]), ms, max_iter,
                                           tol, allow_zero_toi, toi)
            mark("narrow_" + tag)
        # ONE all-reduce at the end: the edge-edge pass prunes with this rank's own vertex-face
        # bound, which changes no result (the minimum is order-independent) and saves a
        # collective + host sync in the middle of the step
        if self.world > 1:
            toi, loads = gather_min_and_loads(toi, info["pairs_local"], dev, self.group)
            info["imbalance"] = imbalance(loads)
            if self.mode == "auto":
                self.rebalance_pairs = info["imbalance"] > self.REBALANCE_ABOVE
        mark("allreduce")
        if marks:
            torch.cuda.synchronize()
            info["ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])}
        self.last = info
        ctx.set_shard(0, 1)
        return toi
