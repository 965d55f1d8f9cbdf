"""Host-side logic of the N>1 path on CPU: world_size-2 (and 3) gloo process groups."""
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


def _worker(rank, world, port, counts, q):
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import load_package
    mg = load_package().multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo = sum(counts[:rank])
        mine = torch.arange(lo, lo + counts[rank], dtype=torch.int32)
        items = torch.stack([mine, -mine], dim=1)            # (n_r, 2), global order = value
        out, seen = mg.rebalance(items)
        toi = mg.allreduce_min(0.25 + rank, torch.device("cpu"))
        toi2, loads = mg.gather_min_and_loads(0.5 + rank, [10 * rank, 7], torch.device("cpu"))
        assert toi2 == 0.5 and loads == [[10 * r, 7] for r in range(world)]
        assert abs(mg.imbalance(loads) - (10 * (world - 1) + 7) /
                   (sum(10 * r + 7 for r in range(world)) / world)) < 1e-12
        q.put((rank, out.numpy().copy(), seen, toi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[10, 0], [3, 17], [1000, 1], [5, 5, 90], [0, 0]])
def test_rebalance_preserves_order_and_evens_out(counts):
    world = len(counts)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, counts, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(counts)
    cat = np.concatenate([r[1] for r in res]) if total else np.zeros((0, 2), np.int32)
    assert np.array_equal(cat[:, 0], np.arange(total, dtype=np.int32))   # order preserved
    assert np.array_equal(cat[:, 1], -np.arange(total, dtype=np.int32))
    sizes = [len(r[1]) for r in res]
    assert max(sizes) - min(sizes) <= 1                                   # balanced
    assert all(r[2] == counts for r in res)
    assert all(r[3] == 0.25 for r in res)                                 # min over ranks


def test_balance_plan_is_consistent(sccd):
    mg = sccd.multigpu
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for _ in range(20):
            counts = rng.integers(0, 1000, world).tolist()
            plans = [mg.balance_plan(counts, r) for r in range(world)]
            for s in range(world):
                assert sum(plans[s][0]) == counts[s]
                for d in range(world):
                    assert plans[s][0][d] == plans[d][1][s]      # what s sends d, d expects
            got = [sum(p[1]) for p in plans]
            assert sum(got) == sum(counts) and max(got) - min(got) <= 1


def _mesh_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from _pkg import load_package
    sccd = load_package()
    mg = sccd.multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = sccd.scenes.cloth_on_sphere(11, seed=3, sphere="uv")
        flat, offs = mg.pack_mesh(s["V0"], s["V1"], s["E"], s["F"], world, pin=False)
        out = torch.full((flat.numel(),), 255, dtype=torch.uint8)
        mg.gather_mesh(flat, out)
        q.put((rank, bool(torch.equal(out, flat)), offs, flat.numel()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_mesh_slices_all_gather_to_the_whole_mesh(world, sccd):
    """Host-buffer entry at N > 1: every rank moves 1/N of the packed mesh, the all-gather
    rebuilds the whole mesh on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mesh_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert all(r[3] % (16 * world) == 0 for r in res)
    # the packed layout is the C ABI's: column-major arrays at 16-byte aligned offsets
    s = sccd.scenes.cloth_on_sphere(11, seed=3, sphere="uv")
    flat, offs = sccd.multigpu.pack_mesh(s["V0"], s["V1"], s["E"], s["F"], world, pin=False)
    raw = flat.numpy()
    nV, nE = s["V0"].shape[0], s["E"].shape[0]
    v1 = np.frombuffer(raw[offs[1]:offs[1] + 24 * nV].tobytes(), np.float64).reshape(3, nV).T
    e = np.frombuffer(raw[offs[2]:offs[2] + 8 * nE].tobytes(), np.int32).reshape(2, nE).T
    assert np.array_equal(v1, s["V1"]) and np.array_equal(e, s["E"])
    assert all(o % 16 == 0 for o in offs)
