"""Host-side logic of the N>1 path on CPU (no GPU needed): the exchange plan the C++ side derives
from the all-gathered send-count matrix (sccd_exchange_plan, csrc/shard.cu) -- checked for
consistency between ranks and replayed with real world_size-2 and -3 gloo process groups."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [1, 2, 3, 8, 16])
def test_exchange_plan_is_consistent_between_ranks(sccd, world):
    rng = np.random.default_rng(world)
    counts = rng.integers(0, 1000, (world, world)).astype(np.uint64)
    counts[rng.random((world, world)) < 0.2] = 0
    plans = [sccd.capi.exchange_plan(counts, r) for r in range(world)]
    for r, (send_off, recv_cnt, recv_off, recv_total) in enumerate(plans):
        # what r receives from s is what s holds for r; sources land in rank order, gap-free
        assert np.array_equal(recv_cnt, counts[:, r])
        assert np.array_equal(recv_off, np.concatenate([[0], np.cumsum(counts[:, r])[:-1]]))
        assert recv_total == int(counts[:, r].sum())
        # the parts for the destinations tile this rank's dest-grouped send buffer
        assert np.array_equal(send_off, np.concatenate([[0], np.cumsum(counts[r])[:-1]]))
    with pytest.raises(sccd.SccdError):
        sccd.capi.exchange_plan(np.zeros((17, 17), np.uint64), 0)      # world > 16


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import load_package
    mg = load_package().multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank holds a slice of "records"; record v belongs to rank (v * 7) % world.
        # A record carries (source rank, position in the source's slice): the receiver must see
        # sources in rank order and, within a source, slice order -- what makes the sharded
        # sorted list equal to the single-GPU one.
        n = 50 + 13 * rank
        vals = torch.arange(n, dtype=torch.int64) * 3 + rank
        dest = (vals * 7) % world
        payloads = [(vals[dest == d] * 1000 + rank) for d in range(world)]
        mine = torch.tensor([len(p) for p in payloads], dtype=torch.int64)
        rows = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(rows, mine)                     # the count matrix, as ncclAllGather does
        counts = torch.stack(rows).numpy()
        got, recv_off, recv_cnt = mg.simulate_exchange(counts, payloads)
        q.put((rank, got.numpy().copy(), recv_off, recv_cnt))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_replayed_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = []
    for rank, got, recv_off, recv_cnt in res:
        src = got % 1000
        val = got // 1000
        assert np.all((val * 7) % world == rank)                     # only this rank's records
        assert np.all(np.diff(src) >= 0)                             # sources in rank order
        for s in range(world):
            part = val[recv_off[s]:recv_off[s] + recv_cnt[s]]
            assert np.all(src[recv_off[s]:recv_off[s] + recv_cnt[s]] == s)
            assert np.all(np.diff(part) > 0)                         # slice order within a source
        seen.append(val * 1000 + src)
    allv = np.sort(np.concatenate(seen))
    want = np.sort(np.concatenate([(np.arange(50 + 13 * r) * 3 + r) * 1000 + r for r in range(world)]))
    assert np.array_equal(allv, want)                                # nothing lost, nothing twice


def test_unique_id_and_world_one_comm_without_gpu(sccd):
    """NCCL is loaded at run time: the library itself must load and export the multi-GPU entry
    points on a box without NCCL or GPUs; the id call either works or says NCCL is missing."""
    L = sccd.capi.load()
    for name in ("sccd_comm_get_unique_id", "sccd_comm_create", "sccd_ccd_sharded",
                 "sccd_ccd_sharded_host", "sccd_comm_destroy", "sccd_exchange_plan"):
        assert hasattr(L, name)
    try:
        uid = sccd.Context.comm_unique_id()
        assert len(uid) == sccd.capi.UNIQUE_ID_BYTES and any(uid)
    except sccd.SccdError as e:
        assert e.code in (sccd.capi.ERR_STATE, sccd.capi.ERR_CUDA)
