"""Freeze golden vectors from the UNMODIFIED reference (oracle/_ref, see oracle/Makefile).

  python tests/golden/make_golden.py cpu  [outdir]   # reference CPU broad phase (no GPU)
  python tests/golden/make_golden.py cuda [outdir]   # reference CUDA path (needs a GPU)
  python tests/golden/make_golden.py cpu_f32  [outdir]  # the same three for the reference's float
  python tests/golden/make_golden.py prep_f32 [outdir]  #   build (SCALABLE_CCD_USE_DOUBLE off);
  python tests/golden/make_golden.py cuda_f32 [outdir]  #   prep_f32 = CPU-side query selection

Inputs are the deterministic generators in scalable-ccd_b200/scenes.py; each fixture
stores a hash of its inputs so generator drift is detected.  Outputs committed under
tests/golden/ are what tests/test_oracle.py and tests/test_gpu_parity.py check against.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from _pkg import load_package  # noqa: E402
from oracle import orc  # noqa: E402

sccd = load_package()
scenes = sccd.scenes


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def scene_hash(s):
    return sha(s["V0"], s["V1"], s["E"], s["F"])


def small_scene():
    return scenes.cloth_on_sphere(31, seed=7, sphere="uv")


def c5_small(n=3000):
    return scenes.queries_c5(n, seed=4)


NARROW_CASES = [  # (name, ms, max_iter, tol, allow_zero_toi)
    ("default", 0.0, -1, 1e-6, True),
    ("tight", 0.0, -1, 1e-9, True),
    ("ms", 1e-8, -1, 1e-6, True),
    ("nozero", 0.0, -1, 1e-6, False),
]


def make_cpu(out, f32=False):
    sfx = "_f32" if f32 else ""
    fix = {}
    for name, s in (("c1", scenes.scene_c1()), ("small", small_scene()),
                    ("pile", scenes.blob_pile(60, seed=5)),               # configs 3 / 4 in small
                    ("slab", scenes.blob_pile(60, seed=6, slab=True))):
        r = orc.ref_cpu_broad_phase(s, f32=f32)
        vf, ee = orc.canonical(r["vf"]), orc.canonical(r["ee"])
        assert len(vf) == r["n_vf"] and len(ee) == r["n_ee"], "reference emitted duplicates"
        fix[name] = {
            "scene_sha256": scene_hash(s), "n_vf": int(r["n_vf"]), "n_ee": int(r["n_ee"]),
            "next_axes": list(r["axes"]), "vf_sha256": sha(vf), "ee_sha256": sha(ee),
            "sizes": [int(s["V0"].shape[0]), int(s["E"].shape[0]), int(s["F"].shape[0])],
        }
        vb, eb, fb = orc.ref_cpu_build_boxes(s, f32=f32)
        fix[name]["boxes_sha256"] = sha(vb, eb, fb)
        if f32:  # inflated boxes too: the radius goes through nextafterf((float)r)
            vb, eb, fb = orc.ref_cpu_build_boxes(s, 1e-3, f32=True)
            fix[name]["boxes_r1e-3_sha256"] = sha(vb, eb, fb)
        if name == "small":
            np.savez_compressed(os.path.join(out, f"broad_small_ref_cpu{sfx}.npz"), vf=vf, ee=ee)
    with open(os.path.join(out, f"broad_ref_cpu{sfx}.json"), "w") as f:
        json.dump(fix, f, indent=1, sort_keys=True)
    print("wrote", out, json.dumps(fix)[:300])


def make_cuda(out):
    meta = {}
    # 1. whole pipeline, TOI_PER_QUERY build: toi + collisions
    # Only scenes whose breadth-first front fits the reference's ring queue (2x the query
    # count, memory_handler.cpp:116): on the 31x31 "small" scene the edge-edge front is
    # 32,308 boxes against 14,940 slots, the racy full-check (ccd_buffer.cuh:25-34) only
    # sometimes fires, and the reference's collision list changes from run to run.
    for name, s in (("c1", scenes.scene_c1()),):
        want = orc.ccd(s)
        for pairs, is_vf in ((want["vf"], True), (want["ee"], False)):
            _, lv, _, front = orc.narrow_phase_bfs(orc.gather_queries(s, pairs, is_vf), is_vf)
            assert lv > 0 and front < 2 * len(pairs), (name, front, len(pairs))
            meta[f"front_{name}_{'vf' if is_vf else 'ee'}"] = [int(front), 2 * len(pairs)]
        r = orc.ref_cuda_ccd(s, per_query=True, coll_cap=1 << 22)
        r0 = orc.ref_cuda_ccd(s, per_query=False)
        order = np.lexsort((r["coll_ids"][:, 1], r["coll_ids"][:, 0]))
        meta[f"ccd_{name}"] = {"scene_sha256": scene_hash(s), "toi_pq": r["toi"],
                               "toi": r0["toi"], "n_coll": int(r["n_coll"])}
        np.savez_compressed(os.path.join(out, f"ccd_{name}_ref_cuda.npz"),
                            coll_ids=r["coll_ids"][order], coll_toi=r["coll_toi"][order],
                            toi=np.float64(r0["toi"]), toi_pq=np.float64(r["toi"]))
        b = orc.ref_cuda_broad_phase(s)
        vf, ee = orc.canonical(b["vf"]), orc.canonical(b["ee"])
        meta[f"broad_{name}"] = {"n_vf": int(b["n_vf"]), "n_ee": int(b["n_ee"]),
                                 "n_vf_unique": len(vf), "n_ee_unique": len(ee),
                                 "vf_sha256": sha(vf), "ee_sha256": sha(ee)}
        # IPC strategy
        meta[f"ipc_{name}"] = {}
        L = orc.ref_cuda(False)
        import ctypes as C
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        for md, mi in ((0.0, -1), (1e-4, 200)):
            el = C.c_double(0)
            t = L.ref_cuda_ipc_ccd_strategy(
                p(s["V0"]), p(s["V1"]), C.c_int64(s["V0"].shape[0]), p(s["E"]),
                C.c_int64(s["E"].shape[0]), p(s["F"]), C.c_int64(s["F"].shape[0]),
                C.c_double(md), C.c_int(mi), C.c_double(1e-6), C.byref(el))
            meta[f"ipc_{name}"][f"md{md}_mi{mi}"] = t
    # 2. root finder on adversarial query arrays
    ee_q, vf_q = c5_small()
    arrays = {}
    for cname, ms, mi, tol, az in NARROW_CASES:
        for kind, q in (("vf", vf_q), ("ee", ee_q)):
            # uncapped runs only on queries the solver finishes (see orc.tractable)
            mask = orc.tractable(q, kind == "vf", ms, tol, az)
            # ... and whose breadth-first front fits the reference's ring queue (it wraps
            # silently otherwise and loses hits, see orc.tractable_bfs)
            mask &= orc.tractable_bfs(q, kind == "vf", ms, tol, az)
            arrays[f"{cname}_{kind}_idx"] = np.flatnonzero(mask).astype(np.int32)
            q = q[mask]
            r = orc.ref_cuda_narrow_queries(q, kind == "vf", ms, mi, tol, az, 1.0, True)
            g = orc.ref_cuda_narrow_queries(q, kind == "vf", ms, mi, tol, az, 1.0, False)
            arrays[f"{cname}_{kind}_tpq"] = r["toi_per_query"]
            arrays[f"{cname}_{kind}_toi"] = np.float64(g["toi"])
            meta[f"narrow_{cname}_{kind}"] = {
                "toi_pq_build": r["toi"], "toi": g["toi"], "reruns": [r["reruns"], g["reruns"]],
                "hits": int((r["toi_per_query"] < 1).sum())}
    meta["c5_sha256"] = sha(ee_q, vf_q)
    np.savez_compressed(os.path.join(out, "narrow_c5_ref_cuda.npz"), **arrays)
    with open(os.path.join(out, "ref_cuda_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1)[:3000])


# ---- the reference's float build ------------------------------------------------------
# Float queries are far cheaper to select than to solve on the GPU box, so the selection
# (which queries the reference can solve without its ring queue wrapping) is made on the CPU
# with the float oracle and committed; cuda_f32 then only runs the reference.
NARROW_CASES_F32 = [  # (name, ms, max_iter, tol, allow_zero_toi)
    ("default", 0.0, -1, 1e-6, True),
    ("loose", 0.0, -1, 1e-4, True),
    ("ms", 1e-5, -1, 1e-6, True),
    ("nozero", 0.0, -1, 1e-6, False),
]


def prep_f32(out):
    ee_q, vf_q = c5_small()
    arrays = {}
    for cname, ms, mi, tol, az in NARROW_CASES_F32:
        for kind, q in (("vf", vf_q), ("ee", ee_q)):
            mask = orc.tractable(q, kind == "vf", ms, tol, az, f32=True)
            mask &= orc.tractable_bfs(q, kind == "vf", ms, tol, az, f32=True)
            arrays[f"{cname}_{kind}_idx"] = np.flatnonzero(mask).astype(np.int32)
            print(cname, kind, int(mask.sum()))
    # (The float build's error filters are ~1e9 times the double build's, so breadth-first
    # fronts are much wider: on config 1 the vertex-face front is 132,722 boxes against the
    # 43,424 slots the reference's pipeline gives its ring queue, which then wraps and loses
    # hits at random -- see make_cuda().  The pipeline golden of the float build is therefore
    # composed from the reference's broad phase and its root finder run with a large queue.)
    arrays["c5_sha256"] = np.frombuffer(bytes.fromhex(sha(ee_q, vf_q)), np.uint8)
    np.savez_compressed(os.path.join(out, "narrow_c5_f32_idx.npz"), **arrays)


def make_cuda_f32(out):
    here = os.path.dirname(os.path.abspath(__file__))
    sel = np.load(os.path.join(here, "narrow_c5_f32_idx.npz"))
    meta = {}
    s = scenes.scene_c1()
    b = orc.ref_cuda_broad_phase(s, f32=True)
    vf, ee = orc.canonical(b["vf"]), orc.canonical(b["ee"])
    meta["broad_c1"] = {"n_vf": int(b["n_vf"]), "n_ee": int(b["n_ee"]), "n_vf_unique": len(vf),
                        "n_ee_unique": len(ee), "vf_sha256": sha(vf), "ee_sha256": sha(ee),
                        "scene_sha256": scene_hash(s)}
    # pipeline = reference broad phase + reference root finder (large queue) on its pairs
    tv = orc.ref_cuda_narrow_queries(orc.gather_queries(s, vf, True), True, f32=True)
    te = orc.ref_cuda_narrow_queries(orc.gather_queries(s, ee, False), False, f32=True)
    toi = min(1.0, tv["toi"], te["toi"])
    hv, he = tv["toi_per_query"] < 1, te["toi_per_query"] < 1
    ids = np.concatenate([vf[hv], ee[he]])
    tq = np.concatenate([tv["toi_per_query"][hv], te["toi_per_query"][he]])
    np.savez_compressed(os.path.join(out, "ccd_c1_ref_cuda_f32.npz"), coll_ids=ids, coll_toi=tq,
                        n_vf_hits=np.int64(hv.sum()), toi=np.float64(toi))
    meta["ccd_c1"] = {"toi": toi, "n_coll": int(len(ids)), "reruns": [tv["reruns"], te["reruns"]]}
    ee_q, vf_q = c5_small()
    arrays = {}
    for cname, ms, mi, tol, az in NARROW_CASES_F32:
        for kind, q in (("vf", vf_q), ("ee", ee_q)):
            idx = sel[f"{cname}_{kind}_idx"]
            qq = q[idx]
            r = orc.ref_cuda_narrow_queries(qq, kind == "vf", ms, mi, tol, az, 1.0, True, f32=True)
            g = orc.ref_cuda_narrow_queries(qq, kind == "vf", ms, mi, tol, az, 1.0, False, f32=True)
            arrays[f"{cname}_{kind}_tpq"] = r["toi_per_query"]
            arrays[f"{cname}_{kind}_toi"] = np.float64(g["toi"])
            meta[f"narrow_{cname}_{kind}"] = {
                "toi_pq_build": r["toi"], "toi": g["toi"], "reruns": [r["reruns"], g["reruns"]],
                "hits": int((r["toi_per_query"] < 1).sum()), "n": int(len(idx))}
    meta["c5_sha256"] = sha(ee_q, vf_q)
    np.savez_compressed(os.path.join(out, "narrow_c5_ref_cuda_f32.npz"), **arrays)
    with open(os.path.join(out, "ref_cuda_f32_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1)[:3000])


if __name__ == "__main__":
    mode = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.dirname(os.path.abspath(__file__))
    os.makedirs(out, exist_ok=True)
    {"cpu": make_cpu, "cuda": make_cuda, "cpu_f32": lambda o: make_cpu(o, True),
     "prep_f32": prep_f32, "cuda_f32": make_cuda_f32}[mode](out)
