"""Multi-GPU launch plumbing.  The multi-GPU pipeline itself lives behind the C ABI
(include/sccd.h "multi-GPU", csrc/shard.cu): slices of the elements, 8-byte records exchanged by
owning cell range over NCCL, local sort / sweep / narrow phase, the earliest TOI published to
every rank over NVLink and closed by one all-reduce(min).  What is left to the host language is
what any launcher does:

  * one process per GPU (torchrun / mpirun), each with one Context on its LOCAL_RANK device;
  * rank 0 makes the NCCL id (Context.comm_unique_id) and the others receive it -- here through
    a torch.distributed group (gloo is enough), in tests/cpp/sharded_ccd.cpp through a file;
  * Context.comm_create(id, rank, world), then Context.ccd_sharded() on every rank.
"""
from __future__ import annotations


class _DevArray:
    """Zero-copy torch / cupy view of a raw device pointer (CUDA array interface v2)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def attach(ctx, group=None):
    """Attach `ctx` (this rank's Context) to a communicator spanning the ranks of a
    torch.distributed process group: rank 0's NCCL id is broadcast through the group."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = [type(ctx).comm_unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=dist.get_global_rank(group, 0) if group else 0,
                                   group=group)
    ctx.comm_create(uid[0], rank, world)
    return rank, world


def simulate_exchange(counts, payloads, group=None):
    """The record exchange of csrc/shard.cu replayed on the host with a torch.distributed
    all-to-all (gloo in the CPU tests): `payloads[d]` is what this rank holds for rank d
    (1-D int64 tensors, already grouped by destination); `counts[src][dst]` is the all-gathered
    count matrix.  Returns what this rank receives, laid out by sccd_exchange_plan -- sources in
    rank order -- so the tests check the plan the C++ side uses against a real exchange."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import capi
    rank = dist.get_rank(group)
    send_off, recv_cnt, recv_off, recv_total = capi.exchange_plan(np.asarray(counts), rank)
    send = torch.cat(list(payloads)) if payloads else torch.zeros(0, dtype=torch.int64)
    assert [int(x) for x in send_off] == [int(x) for x in np.cumsum([0] + [len(p) for p in payloads])[:-1]]
    out = torch.empty(recv_total, dtype=torch.int64)
    dist.all_to_all_single(out, send, [int(x) for x in recv_cnt], [len(p) for p in payloads],
                           group=group)
    return out, [int(x) for x in recv_off], [int(x) for x in recv_cnt]
