"""Per-step times of consecutive ccd() calls on one scene, for a list of NARROW_FLAGS values:
shows frame-to-frame state (lower-bound pause, launch speculation, grid reuse) at work.
  python tools/time_steps.py c3 0 0x80"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from _pkg import load_package
sccd = load_package()
name = sys.argv[1]
flags = [int(v, 0) for v in sys.argv[2:]] or [0]
gen = {"c1": sccd.scenes.scene_c1, "c2": sccd.scenes.scene_c2, "c3": sccd.scenes.scene_c3,
       "c4": sccd.scenes.scene_c4}[name]
s = gen()
ctx = sccd.Context(0)
ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
out = {"workload": name}
for fv in flags:
    ctx.set_option(sccd.capi.OPT_NARROW_FLAGS, fv)
    steps = []
    for i in range(int(os.environ.get("STEPS", "24"))):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        toi = ctx.ccd()
        b.record(); torch.cuda.synchronize()
        st = ctx.stats()
        steps.append((round(a.elapsed_time(b), 3), st["n_box_checks"], st["n_skipped"], st["n_relaunched"], st["n_culled"]))
    out[hex(fv)] = {"toi": toi, "steps": steps}
ctx.close()
print(json.dumps(out))
