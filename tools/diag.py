"""Quick per-stage diagnostics for one workload on the GPU (not a benchmark)."""
import json
import sys
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _pkg import load_package  # noqa: E402
import bench  # noqa: E402

sccd = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene, desc = bench.make_scene(sccd.scenes, name)
ctx = sccd.Context(0)
ctx.upload_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"])
for i in range(reps):
    t = time.perf_counter()
    toi = ctx.ccd(**bench.PARAMS)
    wall = (time.perf_counter() - t) * 1e3
    st = ctx.stats()
    keep = {k: st[k] for k in ("n_pairs", "n_candidates", "n_box_checks", "n_donated", "queue_overflow",
                               "ms_k_sweep_count", "ms_k_sweep_fill", "ms_k_narrow", "ms_k_boxes",
                               "ms_k_gather", "ms_build", "ms_sort", "ms_total", "n_launches")}
    print(json.dumps({"rep": i, "toi": toi, "wall_ms": round(wall, 3), **keep}))
