"""CPU tests: the oracle against the reference's golden vectors, brute force and known
answers.  (The reference's own fixtures need downloaded sample data)"""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def scene_hash(s):
    return sha(s["V0"], s["V1"], s["E"], s["F"])


@pytest.fixture(scope="module")
def gold_cpu():
    return json.load(open(os.path.join(GOLD, "broad_ref_cpu.json")))


def test_scene_generators_are_stable(scene_c1, scene_small, gold_cpu):
    assert scene_hash(scene_c1) == gold_cpu["c1"]["scene_sha256"]
    assert scene_hash(scene_small) == gold_cpu["small"]["scene_sha256"]
    assert [scene_c1["V0"].shape[0], scene_c1["E"].shape[0], scene_c1["F"].shape[0]] \
        == gold_cpu["c1"]["sizes"]


def golden_scene(sccd, scene_c1, scene_small, name):
    """The scenes the reference-CPU goldens were frozen on (tests/golden/make_golden.py)."""
    if name == "pile":      # configs 3 / 4 in small: rigid blob instances, heavy-tailed pile
        return sccd.scenes.blob_pile(60, seed=5)
    if name == "slab":
        return sccd.scenes.blob_pile(60, seed=6, slab=True)
    return {"small": scene_small, "c1": scene_c1}[name]


@pytest.mark.parametrize("name", ["small", "c1", "pile", "slab"])
def test_oracle_broad_phase_matches_reference_golden(orc, sccd, scene_c1, scene_small, gold_cpu,
                                                     name):
    s = golden_scene(sccd, scene_c1, scene_small, name)
    g = gold_cpu[name]
    assert scene_hash(s) == g["scene_sha256"]
    vb, eb, fb = orc.build_boxes(s)
    assert sha(vb, eb, fb) == g["boxes_sha256"]          # boxes bit-exact
    vf, ax_vf = orc.sort_and_sweep_two_lists(vb, fb, 0)
    ee, ax_ee = orc.sort_and_sweep(eb, 0)
    assert (len(vf), len(ee)) == (g["n_vf"], g["n_ee"])
    assert sha(orc.canonical(vf)) == g["vf_sha256"]
    assert sha(orc.canonical(ee)) == g["ee_sha256"]
    assert [ax_vf, ax_ee] == g["next_axes"]


def test_oracle_equals_committed_reference_pairs(orc, scene_small):
    z = np.load(os.path.join(GOLD, "broad_small_ref_cpu.npz"))
    vb, eb, fb = orc.build_boxes(scene_small)
    vf, _ = orc.sort_and_sweep_two_lists(vb, fb, 0)
    ee, _ = orc.sort_and_sweep(eb, 0)
    assert np.array_equal(orc.canonical(vf), z["vf"])
    assert np.array_equal(orc.canonical(ee), z["ee"])


def test_oracle_sweep_equals_brute_force_all_axes(orc, scene_small):
    vb, eb, fb = orc.build_boxes(scene_small, r=1e-3)
    bf_vf = orc.canonical(orc.brute_force(vb, fb))
    bf_ee = orc.canonical(orc.brute_force(eb))
    for axis in (0, 1, 2):
        vf, _ = orc.sort_and_sweep_two_lists(vb, fb, axis)
        ee, _ = orc.sort_and_sweep(eb, axis)
        assert len(vf) == len(bf_vf) and np.array_equal(orc.canonical(vf), bf_vf)
        assert len(ee) == len(bf_ee) and np.array_equal(orc.canonical(ee), bf_ee)


def test_oracle_against_live_reference_cpu(orc, scene_small):
    if orc.ref_cpu() is None:
        pytest.skip("oracle/_ref not built")
    for r in (0.0, 2e-3):
        ref = orc.ref_cpu_broad_phase(scene_small, r=r)
        vb, eb, fb = orc.build_boxes(scene_small, r=r)
        rb = orc.ref_cpu_build_boxes(scene_small, r=r)
        for a, b in zip((vb, eb, fb), rb):
            assert a.tobytes() == b.tobytes()
        vf, _ = orc.sort_and_sweep_two_lists(vb, fb, 0)
        ee, _ = orc.sort_and_sweep(eb, 0)
        assert np.array_equal(orc.canonical(vf), orc.canonical(ref["vf"]))
        assert np.array_equal(orc.canonical(ee), orc.canonical(ref["ee"]))


def test_empty_inputs(orc):
    empty = np.zeros(0, orc.AABB_DTYPE)
    p, _ = orc.sort_and_sweep(empty, 0)
    assert len(p) == 0
    t, tpq, _ = orc.narrow_phase(np.zeros((0, 24)), True)
    assert t == 1.0 and len(tpq) == 0


# ---- narrow phase known answers ---------------------------------------------------
def vf_query(p0, p1, tri0, tri1):
    return np.concatenate([p0, *tri0, p1, *tri1]).astype(np.float64)


TRI = [np.array([0., 0., 0.]), np.array([1., 0., 0.]), np.array([0., 1., 0.])]


def test_vertex_hits_static_triangle_at_half(orc):
    q = vf_query([0.25, 0.25, 1.0], [0.25, 0.25, -1.0], TRI, TRI)
    toi, tpq, st = orc.narrow_phase(q[None], True, tol=1e-6)
    assert tpq[0] == toi and toi <= 0.5 and 0.5 - toi < 1e-5   # conservative, tight
    assert st["box_checks"] > 10


def test_vertex_misses_triangle(orc):
    q = vf_query([2.0, 2.0, 1.0], [2.0, 2.0, -1.0], TRI, TRI)
    toi, tpq, _ = orc.narrow_phase(q[None], True)
    assert toi == 1.0 and np.isinf(tpq[0])


def test_edge_edge_crossing(orc):
    a0, a1 = np.array([-1., 0., 1.]), np.array([1., 0., 1.])
    b0, b1 = np.array([0., -1., 0.]), np.array([0., 1., 0.])
    q = np.concatenate([a0, a1, b0, b1, a0 - [0, 0, 4], a1 - [0, 0, 4], b0, b1])
    toi, tpq, _ = orc.narrow_phase(q[None], False, tol=1e-6)
    assert toi <= 0.25 and 0.25 - toi < 1e-5 and tpq[0] == toi


def test_min_separation_makes_it_earlier(orc):
    q = vf_query([0.25, 0.25, 1.0], [0.25, 0.25, -1.0], TRI, TRI)
    t0, _, _ = orc.narrow_phase(q[None], True, ms=0.0)
    # (a face-parallel approach with a large ms makes the solver resolve the whole contact
    # patch at tolerance -- inherent to Tight-Inclusion -- so keep ms small)
    t1, _, _ = orc.narrow_phase(q[None], True, ms=1e-3, tol=1e-4)
    assert t1 < t0 and abs(t1 - 0.4995) < 1e-3


def test_allow_zero_toi_flag(orc):
    # vertex resting exactly on the triangle the whole time
    q = vf_query([0.25, 0.25, 0.0], [0.3, 0.25, 0.0], TRI, TRI)
    t_allow, _, _ = orc.narrow_phase(q[None], True, allow_zero_toi=True)
    assert t_allow == 0.0


def test_global_and_per_query_modes_agree(orc, sccd):
    ee, vf = sccd.scenes.queries_c5(400, seed=11)
    for q, is_vf in ((vf, True), (ee, False)):
        tg, _, _ = orc.narrow_phase(q, is_vf, per_query=False)
        tp, tpq, _ = orc.narrow_phase(q, is_vf, per_query=True)
        assert tg == tp == min(1.0, tpq.min())


def test_iteration_cap_is_conservative(orc, sccd):
    ee, vf = sccd.scenes.queries_c5(300, seed=12)
    for q, is_vf in ((vf, True), (ee, False)):
        _, full, _ = orc.narrow_phase(q, is_vf, max_iter=-1)
        _, capped, st = orc.narrow_phase(q, is_vf, max_iter=50, cap_mode=1)
        assert st["capped_queries"] > 0
        assert np.all(capped <= full)          # never later than the uncapped answer
        _, dropped, _ = orc.narrow_phase(q, is_vf, max_iter=50, cap_mode=0)
        assert np.all(capped <= dropped)       # and never later than the reference's drop rule


def test_golden_reference_cuda_narrow(orc, sccd):
    """Oracle vs per-query TOIs frozen from the unmodified reference CUDA kernels."""
    path = os.path.join(GOLD, "narrow_c5_ref_cuda.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet (tests/golden/make_golden.py cuda)")
    z = np.load(path)
    meta = json.load(open(os.path.join(GOLD, "ref_cuda_meta.json")))
    ee, vf = sccd.scenes.queries_c5(3000, seed=4)
    assert sha(ee, vf) == meta["c5_sha256"]
    cases = {"default": (0.0, -1, 1e-6, True), "tight": (0.0, -1, 1e-9, True),
             "ms": (1e-8, -1, 1e-6, True), "nozero": (0.0, -1, 1e-6, False)}
    for cname, (ms, mi, tol, az) in cases.items():
        for kind, q in (("vf", vf), ("ee", ee)):
            idx = z[f"{cname}_{kind}_idx"]
            # the subset the reference can solve without its ring queue wrapping
            # (make_golden.py: orc.tractable & orc.tractable_bfs)
            assert len(idx) > 800 and orc.tractable(q, kind == "vf", ms, tol, az)[idx].all()
            toi, tpq, _ = orc.narrow_phase(q[idx], kind == "vf", ms, mi, tol, az)
            ref = z[f"{cname}_{kind}_tpq"]
            assert np.array_equal(tpq < 1, ref < 1), (cname, kind)        # hit/miss
            assert np.array_equal(tpq, ref), (cname, kind)                 # bit-exact toi
            assert toi == float(z[f"{cname}_{kind}_toi"])


def test_golden_reference_cuda_ccd(orc, scene_c1):
    scene_small = scene_c1   # the smallest scene whose BFS front fits the reference's queue
    path = os.path.join(GOLD, "ccd_c1_ref_cuda.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet (tests/golden/make_golden.py cuda)")
    z = np.load(path)
    r = orc.ccd(scene_small)
    assert r["toi"] == float(z["toi"]) == float(z["toi_pq"])
    hits_vf = r["vf"][r["toi_vf"] < 1]
    hits_ee = r["ee"][r["toi_ee"] < 1]
    mine = np.concatenate([hits_vf, hits_ee])
    mine_t = np.concatenate([r["toi_vf"][r["toi_vf"] < 1], r["toi_ee"][r["toi_ee"] < 1]])
    order = np.lexsort((mine[:, 1], mine[:, 0]))
    assert np.array_equal(mine[order], z["coll_ids"])
    assert np.array_equal(mine_t[order], z["coll_toi"])
