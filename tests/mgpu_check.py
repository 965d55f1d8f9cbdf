"""torchrun --nproc-per-node N tests/mgpu_check.py [workload]: the sharded pipeline returns the
single-GPU answer, its pair shards partition the single-GPU list, and rank 0 prints timings."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from _pkg import load_package
import bench

sccd = load_package()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
scene, desc = bench.make_scene(sccd.scenes, name)
ctx = sccd.Context(local, torch.cuda.current_stream().cuda_stream)
ctx.upload_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"])
single = ctx.ccd(**bench.PARAMS)                       # every rank: whole problem
full = [ctx.broad_phase(0), ctx.broad_phase(1)] if name in ("c1", "small") else None
sh = sccd.multigpu.ShardedCCD(ctx)
out = {}
for reb in (True, False, "auto"):
    sh.mode = reb; sh.rebalance_pairs = reb is True
    toi = sh.ccd(**bench.PARAMS)
    assert toi == single, (toi, single)
    torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter()
    for _ in range(5):
        sh.ccd(**bench.PARAMS)
    torch.cuda.synchronize(); dist.barrier()
    out[f"ms_rebalance_{reb}"] = (time.perf_counter() - t) / 5 * 1e3
    out[f"pairs_{reb}"] = sh.last
if full is not None:
    ctx.set_shard(rank, world)
    for k in (0, 1):
        mine = torch.from_numpy(ctx.broad_phase(k)).cuda()      # this rank's shard
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.shape[0]], device="cuda"))
        sizes = [int(s.item()) for s in sizes]
        lo = sum(sizes[:rank])
        assert np.array_equal(mine.cpu().numpy(), full[k][lo:lo + sizes[rank]]), "not a partition"
        assert sum(sizes) == len(full[k])
if rank == 0:
    print(json.dumps({"workload": name, "world": world, "toi": single, **out}))
dist.destroy_process_group()
