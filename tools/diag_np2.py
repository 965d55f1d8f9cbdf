"""Narrow-phase diagnostics of one scene: per solver setting, the statistics of a ccd() step."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from _pkg import load_package
sccd = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
gen = {"c1": sccd.scenes.scene_c1, "c2": sccd.scenes.scene_c2, "c3": sccd.scenes.scene_c3}[name]
s = gen()
ctx = sccd.Context(0)
ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
ctx.set_option(sccd.capi.OPT_PROFILE, int(os.environ.get("PROF", "1")))
for solver, flags in [(1, 0), (1, 1 << 6), (0, 0), (1, 0), (0, 1 << 7)]:
    ctx.set_option(sccd.capi.OPT_NARROW_SOLVER, solver)
    ctx.set_option(sccd.capi.OPT_NARROW_FLAGS, flags)
    for _ in range(2):
        toi = ctx.ccd()
    torch.cuda.synchronize()
    st = ctx.stats()
    print(json.dumps({"solver": solver, "flags": flags, "toi": toi, "checks": st["n_box_checks"],
                      "skipped": st["n_skipped"], "culled": st["n_culled"], "items": st["n_round_items"],
                      "rchecks": st["n_round_checks"], "ms_narrow": st["ms_narrow"],
                      "donated": st["n_donated"], "ms_total": st["ms_total"]}))
ctx.close()
