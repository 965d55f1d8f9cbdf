"""Config 5 (adversarial narrow phase): N VF + N EE direct queries (near-parallel edges, grazing
vertex-face), the four parameter sets of SURVEY.md 8d, bounded item lists.  Not a benchmark."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _pkg import load_package  # noqa: E402
import torch  # noqa: E402

sccd = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 22
ee, vf = sccd.scenes.queries_c5(n, seed=4)
ctx = sccd.Context(0)
ctx.set_queue_capacity(cap)          # 2^22 items: forces the bounded-list behaviour
CASES = [dict(ms=0.0, max_iter=10_000, tol=1e-6), dict(ms=1e-8, max_iter=10_000, tol=1e-6),
         dict(ms=0.0, max_iter=10_000, tol=1e-9), dict(ms=0.0, max_iter=-1, tol=1e-6)]
for kind, q in ((0, vf), (1, ee)):
    dq = torch.from_numpy(q).cuda()
    for kw in CASES:
        for rep in range(2):
            ctx.reset_stats()
            torch.cuda.synchronize()
            t = time.perf_counter()
            toi = ctx.narrow_phase_queries(kind, dq.data_ptr(), n=len(q), **kw)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t) * 1e3
        st = ctx.stats()
        print(json.dumps({"kind": "vf" if kind == 0 else "ee", **kw, "n": len(q), "toi": toi,
                          "wall_ms": round(dt, 3), "queries_per_s": round(len(q) / dt * 1e3),
                          "checks": st["n_box_checks"][kind], "handed_on": st["n_donated"][kind],
                          "capped": st["n_capped"][kind], "list_full": st["queue_overflow"]}))
