"""A few ccd() steps of one scene for `ncu --metrics gpu__time_duration.sum` (launch list):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      python tools/launch_list.py c2 [solver] [flags]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _pkg import load_package  # noqa: E402

sccd = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
gen = {"small": lambda: sccd.scenes.cloth_on_sphere(31, seed=7, sphere="uv"), "c1": sccd.scenes.scene_c1,
       "c2": sccd.scenes.scene_c2, "c3": sccd.scenes.scene_c3, "c4": sccd.scenes.scene_c4}[name]
s = gen()
ctx = sccd.Context(0)
if len(sys.argv) > 2:
    ctx.set_option(sccd.capi.OPT_NARROW_SOLVER, int(sys.argv[2]))
if len(sys.argv) > 3:
    ctx.set_option(sccd.capi.OPT_NARROW_FLAGS, int(sys.argv[3], 0))
ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
for _ in range(3):
    toi = ctx.ccd()
print("toi", toi, ctx.stats()["n_launches"], "launches per step")
ctx.close()
