"""Narrow-phase only diagnostics: run the broad phase once, then time narrow-phase variants."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _pkg import load_package
import bench
import torch
sccd = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
scene, _ = bench.make_scene(sccd.scenes, name)
ctx = sccd.Context(0)
ctx.upload_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"])
ctx.build_boxes(0.0)
for kind in (0, 1):
    pairs = torch.from_numpy(ctx.broad_phase(kind)).cuda()
    n = pairs.shape[0]
    tq = torch.empty(n, dtype=torch.float64, device="cuda")
    for mode in ("global", "per_query"):
        for rep in range(2):
            torch.cuda.synchronize(); t = time.perf_counter()
            toi = ctx.narrow_phase(kind, pairs.data_ptr(), n, toi=1.0,
                                   d_toi_per_query=tq.data_ptr() if mode == "per_query" else 0)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
            st = ctx.stats()
            print(json.dumps({"kind": kind, "mode": mode, "n": n, "toi": toi, "wall_ms": round(dt, 3),
                              "k_ms": st["ms_k_narrow"][kind], "checks": st["n_box_checks"][kind],
                              "donated": st["n_donated"][kind], "ovf": st["queue_overflow"],
                              "flags": os.environ.get("SCCD_NP_FLAGS", "0")}))
