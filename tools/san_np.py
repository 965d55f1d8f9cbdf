import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _pkg import load_package
sccd = load_package()
s = sccd.scenes.scene_c1()
ctx = sccd.Context(0)
ctx.set_option(sccd.capi.OPT_NARROW_SOLVER, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
print("toi", ctx.ccd(), ctx.stats()["n_box_checks"])
ctx.close()
