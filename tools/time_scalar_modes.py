"""ccd() step time of one scene in the two scalar modes (SCCD_F64 / SCCD_F32), one JSON line.

  python tools/time_scalar_modes.py [workload] [steps]

The float mode (the reference's SCALABLE_CCD_USE_DOUBLE=OFF build) runs the separating-axis cull
(float filters) and the lane-per-tree float solver; the warp-cooperative kernel is double only.
SCCD_NP_CULL=0 in the environment switches the cull off in both modes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from _pkg import load_package  # noqa: E402

sccd = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gen = {"small": lambda: sccd.scenes.cloth_on_sphere(31, seed=7, sphere="uv"),
       "c1": sccd.scenes.scene_c1, "c2": sccd.scenes.scene_c2}[name]
s = gen()
ctx = sccd.Context(0)
ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
out = {"workload": name, "steps": steps}
for mode, label in ((sccd.capi.F64, "f64"), (sccd.capi.F32, "f32")):
    ctx.set_scalar_type(mode)
    for _ in range(2):
        toi = ctx.ccd()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        toi = ctx.ccd()
    b.record()
    torch.cuda.synchronize()
    st = ctx.stats()
    out[label] = {"ms_per_step": a.elapsed_time(b) / steps, "toi": toi, "n_pairs": st["n_pairs"],
                  "n_box_checks": st["n_box_checks"], "n_culled": st["n_culled"],
                  "ms_narrow": st["ms_narrow"], "ms_sweep": st["ms_sweep"],
                  "ms_total_device": st["ms_total"]}
ctx.close()
print(json.dumps(out))
