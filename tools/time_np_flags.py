"""A/B of the narrow-phase scheduling knobs (SCCD_NP_FLAGS, csrc/narrow.cu) on one scene:
ccd() step time and per-stage device times for each flag value given on the command line.

  python tools/time_np_flags.py c2 0 $((2<<28)) $((1<<28))
  SCCD_NP_FLAGS_EE=$((24<<8)) python tools/time_np_flags.py c2 0     # edge-edge pass alone

Bits 28..30 = log2(cooperative limit) - 13: round 0 goes to the warp-cooperative kernel when at
most limit/2 queries survive the cull.  The edge-edge pass of a cloth scene runs after the
vertex-face pass has set the earliest toi, so its ~30 K surviving trees are ~6 checks each --
a throughput problem that one lane per tree should suit better than one warp per tree."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from _pkg import load_package  # noqa: E402

sccd = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
# an argument is FLAGS or SOLVER:FLAGS (solver = 0, 4 or 8 lanes per tree, SCCD_OPT_NARROW_SOLVER)
def _parse(v):
    sv, _, fv = v.rpartition(":")
    return (int(sv) if sv else 0, int(fv, 0))


flag_values = [_parse(v) for v in sys.argv[2:]] or [(0, 0)]
gen = {"small": lambda: sccd.scenes.cloth_on_sphere(31, seed=7, sphere="uv"),
       "c1": sccd.scenes.scene_c1, "c2": sccd.scenes.scene_c2, "c3": sccd.scenes.scene_c3}[name]
s = gen()
ctx = sccd.Context(0)
ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
ctx.set_option(sccd.capi.OPT_PROFILE, int(os.environ.get("PROF", "1")))
ctx.set_option(sccd.capi.OPT_CONCURRENT_PASSES, int(os.environ.get("CONC", "0")))
out = {"workload": name, "concurrent_passes": int(os.environ.get("CONC", "0"))}
for sv, fv in flag_values:
    ctx.set_option(sccd.capi.OPT_NARROW_SOLVER, sv)
    ctx.set_option(sccd.capi.OPT_NARROW_FLAGS, fv)
    for _ in range(2):
        toi = ctx.ccd()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        toi = ctx.ccd()
    b.record()
    torch.cuda.synchronize()
    st = ctx.stats()
    out[f"{sv}:{hex(fv)}"] = {"ms_per_step": a.elapsed_time(b) / 5, "toi": toi,
                    "round_items": st["n_round_items"], "round_checks": st["n_round_checks"],
                    "n_box_checks": st["n_box_checks"], "ms_narrow": st["ms_narrow"],
                    "ms_k_narrow": st["ms_k_narrow"], "ms_total_device": st["ms_total"]}
ctx.close()
print(json.dumps(out))
