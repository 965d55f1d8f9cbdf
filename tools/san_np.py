"""Small pipeline runs for compute-sanitizer (memcheck / racecheck / initcheck):
  compute-sanitizer --tool memcheck python tools/san_np.py [solver]
Covers: config 1 (work queue, pair step, frame-to-frame second call), a small pile (lane-per-tree
rounds), the per-query list, and the adversarial direct queries."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _pkg import load_package
sccd = load_package()
ctx = sccd.Context(0)
ctx.set_option(sccd.capi.OPT_NARROW_SOLVER, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for name, s in (("c1", sccd.scenes.scene_c1()), ("pile", sccd.scenes.blob_pile(150, seed=2))):
    ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    for _ in range(2):                       # second call: grid reuse + launch speculation
        toi = ctx.ccd()
    print(name, "toi", toi, ctx.stats()["n_box_checks"], ctx.stats()["n_relaunched"])
    t, vf, ee = ctx.ccd_collisions()
    print(name, "collisions", t, len(vf[1]), len(ee[1]))
ee, vf = sccd.scenes.queries_c5(2000, seed=4)
for kind, q in ((0, vf), (1, ee)):
    print("c5", kind, ctx.narrow_phase_queries(kind, q, max_iter=2000))
ctx.close()
