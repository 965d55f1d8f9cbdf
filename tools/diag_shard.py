"""Emulates rank R of a W-rank job on ONE GPU (no NCCL): build+sort and the two sweeps of that
rank's cell range, wall-clock with a device sync on both sides.  Not a benchmark."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _pkg import load_package  # noqa: E402
import bench  # noqa: E402

sccd = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ranks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
scene, desc = bench.make_scene(sccd.scenes, name)
ctx = sccd.Context(0)
ctx.upload_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"])
for r in ranks:
    ctx.set_shard(r, world)
    for rep in range(3):
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.build_boxes(0.0)
        ctx.synchronize()
        t1 = time.perf_counter()
        n_vf = ctx.broad_phase(0, want_pairs=False)
        ctx.synchronize()
        t2 = time.perf_counter()
        n_ee = ctx.broad_phase(1, want_pairs=False)
        ctx.synchronize()
        t3 = time.perf_counter()
    st = ctx.stats()
    print(json.dumps({"rank": r, "world": world, "build_sort_ms": round((t1 - t0) * 1e3, 3),
                      "sweep_vf_ms": round((t2 - t1) * 1e3, 3), "sweep_ee_ms": round((t3 - t2) * 1e3, 3),
                      "pairs": [n_vf, n_ee], "n_records": st["n_records"], "grid": st["grid_cells"]}))
