"""A/B of one sccd_set_option on one scene: ccd() step time, device stage times.
  python tools/time_opts.py c2 SWEEP_STAGED 0 1"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from _pkg import load_package
sccd = load_package()
name, opt = sys.argv[1], getattr(sccd.capi, "OPT_" + sys.argv[2])
values = [int(v, 0) for v in sys.argv[3:]]
gen = {"c1": sccd.scenes.scene_c1, "c2": sccd.scenes.scene_c2, "c3": sccd.scenes.scene_c3,
       "c4": sccd.scenes.scene_c4}[name]
s = gen()
ctx = sccd.Context(0)
ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
out = {"workload": name, "option": sys.argv[2]}
for rep in range(2):
    for v in values:
        ctx.set_option(opt, v)
        ctx.set_option(sccd.capi.OPT_PROFILE, 0)
        for _ in range(3):
            toi = ctx.ccd()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(10):
            toi = ctx.ccd()
        b.record(); torch.cuda.synchronize()
        ctx.set_option(sccd.capi.OPT_PROFILE, 2)
        ctx.ccd(); ctx.ccd()
        st = ctx.stats()
        out[f"{v}#{rep}"] = {"ms_per_step": a.elapsed_time(b) / 10, "toi": toi, "ms_sweep": st["ms_sweep"],
                             "ms_k_sweep_count": st["ms_k_sweep_count"], "n_pairs": st["n_pairs"],
                             "ms_narrow": st["ms_narrow"], "ms_k_cull": st["ms_k_cull"],
                             "ms_k_round": st["ms_k_round"], "n_culled": st["n_culled"],
                             "n_box_checks": st["n_box_checks"], "n_skipped": st["n_skipped"],
                             "n_records": st["n_records"], "n_candidates": st["n_candidates"],
                             "ms_k_sort": st["ms_k_sort"], "ms_k_gather": st["ms_k_gather"],
                             "grid_cells": st["grid_cells"]}
ctx.close()
print(json.dumps(out))
