"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, and against golden vectors frozen from the unmodified reference.

Bar: boxes bit-exact, overlap sets bit-exact after canonical sort, per-query hit/miss equal,
per-query and global TOI bit-equal for max_iter < 0 (tolerance 0.0: the arithmetic contract
reproduces the reference's rounding), conservative (<=) for capped queries."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch


def run_broad(ctx, scene, r=0.0):
    ctx.upload_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"])
    ctx.build_boxes(r)
    return ctx.broad_phase(0), ctx.broad_phase(1)


@pytest.mark.parametrize("r", [0.0, 1e-3])
def test_boxes_bit_exact(ctx, orc, scene_c1, r):
    ctx.upload_mesh(scene_c1["V0"], scene_c1["V1"], scene_c1["E"], scene_c1["F"])
    ctx.build_boxes(r)
    got = ctx.get_boxes()
    want = orc.build_boxes(scene_c1, r)
    for g, w in zip(got, want):
        assert g.tobytes() == w.tobytes()


@pytest.mark.parametrize("which", ["small", "c1"])
def test_overlap_sets_bit_exact(ctx, orc, scene_c1, scene_small, which):
    s = {"small": scene_small, "c1": scene_c1}[which]
    vf, ee = run_broad(ctx, s)
    vb, eb, fb = orc.build_boxes(s)
    ovf, _ = orc.sort_and_sweep_two_lists(vb, fb, 0)
    oee, _ = orc.sort_and_sweep(eb, 0)
    assert len(vf) == len(ovf) and len(ee) == len(oee)           # no duplicates
    assert np.array_equal(orc.canonical(vf), orc.canonical(ovf))
    assert np.array_equal(orc.canonical(ee), orc.canonical(oee))
    gold = json.load(open(os.path.join(GOLD, "broad_ref_cpu.json")))[which]
    assert (len(vf), len(ee)) == (gold["n_vf"], gold["n_ee"])      # reference CPU golden


def test_overlap_list_is_deterministic_and_chunking_is_lossless(ctx, sccd, scene_c1):
    vf1, ee1 = run_broad(ctx, scene_c1)
    vf2, ee2 = run_broad(ctx, scene_c1)
    assert np.array_equal(vf1, vf2) and np.array_equal(ee1, ee2)   # same ORDER, not just set
    ctx.set_max_pairs_per_chunk(5000)                               # force ~15 chunks
    try:
        vf3, ee3 = run_broad(ctx, scene_c1)
    finally:
        ctx.set_max_pairs_per_chunk(0)
    assert np.array_equal(vf1, vf3) and np.array_equal(ee1, ee3)


def test_inflated_boxes_and_brute_force(ctx, orc, scene_small):
    vf, ee = run_broad(ctx, scene_small, r=5e-3)
    vb, eb, fb = orc.build_boxes(scene_small, 5e-3)
    assert np.array_equal(orc.canonical(vf), orc.canonical(orc.brute_force(vb, fb)))
    assert np.array_equal(orc.canonical(ee), orc.canonical(orc.brute_force(eb)))


def test_ragged_and_empty_inputs(ctx, orc, sccd):
    # a single triangle: 3 V / 3 E / 1 F -> no admissible pairs at all
    V = np.asfortranarray(np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0]]))
    F = np.asfortranarray(np.array([[0, 1, 2]], dtype=np.int32))
    E = np.asfortranarray(sccd.scenes.edges_from_faces(F))
    ctx.upload_mesh(V, V.copy(order="F"), E, F)
    assert ctx.ccd() == 1.0
    assert len(ctx.broad_phase(0)) == 0 and len(ctx.broad_phase(1)) == 0
    # no faces / no edges at all
    E0 = np.zeros((0, 2), np.int32, order="F")
    F0 = np.zeros((0, 3), np.int32, order="F")
    ctx.upload_mesh(V, V.copy(order="F"), E0, F0)
    assert ctx.ccd() == 1.0
    # identical boxes (all ties on the sort key): 64 coincident, disconnected edges
    n = 64
    Vt = np.asfortranarray(np.tile(np.array([[0., 0, 0], [1, 1, 1]]), (n, 1)))
    Et = np.asfortranarray(np.arange(2 * n, dtype=np.int32).reshape(n, 2))
    ctx.upload_mesh(Vt, Vt.copy(order="F"), Et, F0)
    ctx.build_boxes(0.0)
    assert len(ctx.broad_phase(1)) == n * (n - 1) // 2


def test_broad_phase_before_build_is_an_error(sccd):
    c = sccd.Context(0)
    with pytest.raises(sccd.SccdError) as e:
        c.broad_phase_begin(0)
    assert e.value.code == sccd.capi.ERR_STATE
    c.close()


def _narrow_gpu(ctx, torch, kind, q, **kw):
    tq = torch.empty(len(q), dtype=torch.float64, device="cuda")
    toi = ctx.narrow_phase_queries(kind, q, d_toi_per_query=tq.data_ptr(), **kw)
    return toi, tq.cpu().numpy()


CASES = {"default": dict(ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True),
         "tight": dict(ms=0.0, max_iter=-1, tol=1e-9, allow_zero_toi=True),
         "ms": dict(ms=1e-8, max_iter=-1, tol=1e-6, allow_zero_toi=True),
         "nozero": dict(ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=False)}


@pytest.mark.parametrize("case", list(CASES))
def test_adversarial_queries_match_oracle(ctx, orc, sccd, torch_cuda, case):
    ee, vf = sccd.scenes.queries_c5(3000, seed=4)
    kw = CASES[case]
    for kind, q in ((0, vf), (1, ee)):
        # uncapped parity only on queries the solver can finish (grazing ms > 0 queries
        # need >1e8 boxes, for the reference too); the capped test covers the rest
        q = q[orc.tractable(q, kind == 0, kw["ms"], kw["tol"], kw["allow_zero_toi"])]
        assert len(q) > 2500
        toi, tpq = _narrow_gpu(ctx, torch_cuda, kind, q, **kw)
        otoi, otpq, _ = orc.narrow_phase(q, kind == 0, kw["ms"], kw["max_iter"], kw["tol"],
                                         kw["allow_zero_toi"])
        assert np.array_equal(tpq < 1, otpq < 1)      # hit / miss
        assert np.array_equal(tpq, otpq)              # bit-exact per-query toi (tolerance 0)
        assert toi == otoi
        # shared-bound mode (default reference build) returns the same minimum
        assert ctx.narrow_phase_queries(kind, q, **kw) == otoi


def test_adversarial_queries_match_reference_cuda_golden(ctx, sccd, torch_cuda):
    path = os.path.join(GOLD, "narrow_c5_ref_cuda.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    z = np.load(path)
    ee, vf = sccd.scenes.queries_c5(3000, seed=4)
    for case, kw in CASES.items():
        for kind, name, q in ((0, "vf", vf), (1, "ee", ee)):
            q = q[z[f"{case}_{name}_idx"]]
            toi, tpq = _narrow_gpu(ctx, torch_cuda, kind, q, **kw)
            assert np.array_equal(tpq, z[f"{case}_{name}_tpq"]), (case, name)
            assert toi == float(z[f"{case}_{name}_toi"])


def test_iteration_cap_is_conservative(ctx, orc, sccd, torch_cuda):
    ee, vf = sccd.scenes.queries_c5(2000, seed=12)
    for kind, q in ((0, vf), (1, ee)):
        _, full, _ = orc.narrow_phase(q, kind == 0, max_iter=-1)
        toi, capped = _narrow_gpu(ctx, torch_cuda, kind, q, max_iter=60)
        # with a minimum separation: every query, including the intractable grazing ones
        _, capped_ms = _narrow_gpu(ctx, torch_cuda, kind, q, ms=1e-8, max_iter=2000)
        m = orc.tractable(q, kind == 0, 1e-8, 1e-6)
        _, full_ms, _ = orc.narrow_phase(q[m], kind == 0, ms=1e-8)
        assert np.all(capped_ms[m] <= full_ms)
        assert ctx.stats()["n_capped"][kind] > 0
        assert np.all(capped <= full)                  # never later than the exact answer
        under = ~(capped < full)                       # queries the cap did not touch
        assert np.array_equal(capped[under], full[under])
        assert toi <= min(1.0, full.min())


@pytest.fixture
def np_env(ctx, sccd):
    """Sets / restores the narrow-phase knobs of the shared context (sccd_set_option)."""
    K = sccd.capi

    def set_(flags=None, depth=None, cull=None):
        ctx.set_option(K.OPT_NARROW_FLAGS, 0 if flags is None else flags)
        ctx.set_option(K.OPT_NARROW_MAX_DEPTH, 128 if depth is None else depth)
        ctx.set_option(K.OPT_NARROW_CULL, 1 if cull is None else cull)
    yield set_
    set_()


# refill (6 bits) | no scout << 6 | no skip << 7 | first-round budget << 8 | later budget (7 bits)
# << 16 | no pair step (warp-per-tree walker checks one box per step) << 22 |
# survivors not sorted by toi lower bound << 23 | never-cooperate << 24 |
# log2(coop budget) - 3 << 25 | log2(coop limit) - 13 << 28
VARIANTS = {"unsorted": 1 << 23, "no_scout": 1 << 6, "no_skip": 1 << 7, "lane_only": 1 << 24, "tiny_budgets": 4 | (3 << 8) | (2 << 16) | (1 << 25),
            "refill_every_lane": 1, "refill_all_idle": 32, "coop_big_budget": 5 << 25,
            "coop_small_limit": 1 << 28, "no_pair_step": 1 << 22,
            "no_pair_tiny_budgets": (1 << 22) | 4 | (3 << 8) | (2 << 16) | (1 << 25)}


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_narrow_phase_scheduling_does_not_change_results(ctx, orc, sccd, torch_cuda, np_env,
                                                          variant):
    """How the bisection trees are cut into rounds / items and which kernel (lane-per-tree or
    warp-cooperative) walks them must not change a single bit of the per-query TOIs."""
    ee, vf = sccd.scenes.queries_c5(1500, seed=9)
    for kind, q in ((0, vf), (1, ee)):
        q = q[orc.tractable(q, kind == 0, 0.0, 1e-6)]
        np_env()
        toi0, tpq0 = _narrow_gpu(ctx, torch_cuda, kind, q)
        np_env(flags=VARIANTS[variant])
        toi1, tpq1 = _narrow_gpu(ctx, torch_cuda, kind, q)
        shared1 = ctx.narrow_phase_queries(kind, q)
        np_env()
        assert np.array_equal(tpq0, tpq1) and toi0 == toi1 == shared1
        _, otpq, _ = orc.narrow_phase(q, kind == 0)
        assert np.array_equal(tpq1, otpq)


@pytest.mark.parametrize("flags", [0, 1 << 24, 1 << 22])
def test_paths_deeper_than_the_lane_state_are_handed_on(ctx, orc, sccd, torch_cuda, np_env, flags):
    """With only 6 trackable levels every non-trivial tree outgrows the walk state in every
    round, including the last: the boxes are handed on and the host keeps adding rounds."""
    ee, vf = sccd.scenes.queries_c5(400, seed=21)
    for kind, q in ((0, vf), (1, ee)):
        q = q[orc.tractable(q, kind == 0, 0.0, 1e-6, limit=3000)]
        assert len(q) > 100
        np_env(flags=flags, depth=6)
        toi, tpq = _narrow_gpu(ctx, torch_cuda, kind, q)
        np_env()
        otoi, otpq, _ = orc.narrow_phase(q, kind == 0)
        assert np.array_equal(tpq, otpq) and toi == otoi


@pytest.mark.parametrize("kw", [dict(ms=0.0, tol=1e-6), dict(ms=1e-3, tol=1e-6),
                                dict(ms=0.0, tol=1e-3), dict(ms=1e-8, tol=1e-9)])
def test_separating_axis_cull_changes_no_result(ctx, orc, sccd, torch_cuda, scene_c1, np_env, kw):
    """The cull answers most candidate pairs "no collision" without the solver; every per-query
    TOI must equal the solver's (cull off) and the oracle's, also with a minimum separation or a
    tolerance far larger than the gaps it tests against."""
    ctx.upload_mesh(scene_c1["V0"], scene_c1["V1"], scene_c1["E"], scene_c1["F"])
    ctx.build_boxes(kw["ms"])
    for kind in (0, 1):
        pairs = ctx.broad_phase(kind)
        q = orc.gather_queries(scene_c1, np.ascontiguousarray(pairs), kind == 0)
        np_env(cull=1)
        ctx.reset_stats()
        toi1, tpq1 = _narrow_gpu(ctx, torch_cuda, kind, q, **kw)
        culled = ctx.stats()["n_culled"][kind]
        np_env(cull=2)                            # separating axes only, no root-box check
        ctx.reset_stats()
        toi2, tpq2 = _narrow_gpu(ctx, torch_cuda, kind, q, **kw)
        assert 0 < ctx.stats()["n_culled"][kind] <= culled
        np_env(cull=0)
        ctx.reset_stats()
        toi0, tpq0 = _narrow_gpu(ctx, torch_cuda, kind, q, **kw)
        assert ctx.stats()["n_culled"][kind] == 0
        np_env()
        assert np.array_equal(tpq0, tpq1) and toi0 == toi1
        assert np.array_equal(tpq0, tpq2) and toi0 == toi2
        otoi, otpq, _ = orc.narrow_phase(q, kind == 0, kw["ms"], -1, kw["tol"], True)
        assert np.array_equal(tpq1, otpq) and toi1 == otoi
        if kw["ms"] == 0.0 and kw["tol"] == 1e-6:
            assert culled > 0.8 * len(q)          # it does its job on a cloth scene


def test_dense_tiles_overflow_the_staging_area_and_chunk(ctx, sccd):
    """600 coincident boxes: every tile finds far more pairs than it can park between the count
    and the place pass, so the place pass sweeps again -- also when the list is cut in chunks
    that end inside tiles."""
    n = 600
    F0 = np.zeros((0, 3), np.int32, order="F")
    Vt = np.asfortranarray(np.tile(np.array([[0., 0, 0], [1, 1, 1]]), (n, 1)))
    Et = np.asfortranarray(np.arange(2 * n, dtype=np.int32).reshape(n, 2))
    ctx.upload_mesh(Vt, Vt.copy(order="F"), Et, F0)
    ctx.build_boxes(0.0)
    full = ctx.broad_phase(1)
    assert len(full) == n * (n - 1) // 2
    assert len(np.unique(full, axis=0)) == len(full)
    ctx.set_max_pairs_per_chunk(7001)
    try:
        chunked = ctx.broad_phase(1)
    finally:
        ctx.set_max_pairs_per_chunk(0)
    assert np.array_equal(full, chunked)


def test_bounded_queue_and_donation(sccd, orc, torch_cuda):
    """A handful of very deep queries: the tail must be spread through the work queue and the
    answer must not depend on it."""
    c = sccd.Context(0)
    c.set_queue_capacity(1)            # clamped to the minimum the kernel accepts
    ee, vf = sccd.scenes.queries_c5(64, seed=5)
    for kind, q in ((0, vf), (1, ee)):
        toi, tpq = _narrow_gpu(c, torch_cuda, kind, q, tol=1e-9)
        otoi, otpq, _ = orc.narrow_phase(q, kind == 0, tol=1e-9)
        assert np.array_equal(tpq, otpq) and toi == otoi
    c.close()


@pytest.mark.parametrize("which", ["small", "c1"])
def test_full_pipeline_matches_oracle(ctx, orc, scene_c1, scene_small, which):
    s = {"small": scene_small, "c1": scene_c1}[which]
    ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    want = orc.ccd(s)
    assert ctx.ccd() == want["toi"]
    st = ctx.stats()
    assert st["n_pairs"] == [len(want["vf"]), len(want["ee"])]
    toi, (vf_ids, vf_t), (ee_ids, ee_t) = ctx.ccd_collisions()
    assert toi == want["toi"]
    for ids, t, pairs, tq in ((vf_ids, vf_t, want["vf"], want["toi_vf"]),
                              (ee_ids, ee_t, want["ee"], want["toi_ee"])):
        order = np.lexsort((ids[:, 1], ids[:, 0]))
        hit = tq < 1
        assert np.array_equal(ids[order], pairs[hit])      # per-query hit / miss
        assert np.array_equal(t[order], tq[hit])           # per-query toi
    # host-pointer entry point (upload inside the call)
    assert ctx.ccd_host(s["V0"], s["V1"], s["E"], s["F"]) == want["toi"]


def test_pipeline_with_min_distance_and_chunks(ctx, orc, scene_small):
    s = scene_small
    ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    want = orc.ccd(s, ms=1e-4, tol=1e-5)
    ctx.set_max_pairs_per_chunk(1000)
    try:
        assert ctx.ccd(ms=1e-4, tol=1e-5) == want["toi"]
    finally:
        ctx.set_max_pairs_per_chunk(0)


def test_pipeline_matches_reference_cuda_golden(ctx, scene_small, scene_c1):
    # (the 31x31 scene is not frozen: the reference's ring queue wraps on it, see
    # tests/golden/make_golden.py)
    for name, s in (("c1", scene_c1),):
        path = os.path.join(GOLD, f"ccd_{name}_ref_cuda.npz")
        if not os.path.exists(path):
            pytest.skip("golden not generated yet")
        z = np.load(path)
        ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        toi, (vf_ids, vf_t), (ee_ids, ee_t) = ctx.ccd_collisions()
        assert toi == float(z["toi"])
        ids = np.concatenate([vf_ids, ee_ids])
        t = np.concatenate([vf_t, ee_t])
        order = np.lexsort((ids[:, 1], ids[:, 0]))
        assert np.array_equal(ids[order], z["coll_ids"])
        assert np.array_equal(t[order], z["coll_toi"])


def test_ipc_ccd_strategy(ctx, orc, scene_small):
    """ipc_ccd_strategy.cu:54-92 composed from oracle pieces.  With an iteration cap the
    result must be conservative w.r.t. the uncapped composition; the reference-CUDA value
    for the same call is frozen in tests/golden/ref_cuda_meta.json."""
    s = scene_small
    ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    meta_path = os.path.join(GOLD, "ref_cuda_meta.json")
    gold = None   # reference values are frozen for config 1 only (see below)
    for md, mi in ((0.0, -1), (1e-4, 200)):
        vb, eb, fb = orc.build_boxes(s, md)
        vf = orc.canonical(orc.sort_and_sweep_two_lists(vb, fb, 0)[0])
        ee = orc.canonical(orc.sort_and_sweep(eb, 0)[0])
        toi = 1.0
        for pairs, is_vf in ((vf, True), (ee, False)):
            q = orc.gather_queries(s, pairs, is_vf)
            before = toi
            toi, _, _ = orc.narrow_phase(q, is_vf, md, -1, 1e-6, True, toi, per_query=False)
            if toi < 1e-6:
                toi, _, _ = orc.narrow_phase(q, is_vf, 0.0, -1, 1e-6, False, before,
                                             per_query=False)
                toi *= 0.8
        got = ctx.ipc_ccd_strategy(md, mi, 1e-6)
        if mi < 0:
            assert got == toi
            if gold:
                assert got == gold[f"md{md}_mi{mi}"]
        else:
            assert got <= toi      # capped: never later than the exact answer
            if gold:               # ... nor than the reference's (which drops boxes)
                assert got <= gold[f"md{md}_mi{mi}"]


def test_ipc_ccd_strategy_matches_reference_cuda_golden(ctx, scene_c1):
    meta_path = os.path.join(GOLD, "ref_cuda_meta.json")
    if not os.path.exists(meta_path):
        pytest.skip("golden not generated yet")
    gold = json.load(open(meta_path))["ipc_c1"]
    s = scene_c1
    ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    assert ctx.ipc_ccd_strategy(0.0, -1, 1e-6) == gold["md0.0_mi-1"]
    # capped: the reference drops boxes (later or equal), we accept them (earlier or equal)
    assert ctx.ipc_ccd_strategy(1e-4, 200, 1e-6) <= gold["md0.0001_mi200"]


def test_device_pointer_inputs(ctx, orc, scene_small, torch_cuda):
    torch = torch_cuda
    s = scene_small
    t = {k: torch.from_numpy(np.ascontiguousarray(v.T)).cuda() for k, v in s.items()}
    ctx.upload_mesh(t["V0"].data_ptr(), t["V1"].data_ptr(), t["E"].data_ptr(), t["F"].data_ptr(),
                    sizes=(s["V0"].shape[0], s["E"].shape[0], s["F"].shape[0]))
    assert ctx.ccd() == orc.ccd(s)["toi"]


@pytest.mark.parametrize("world,max_cells", [(3, 0), (8, 0), (3, 1), (4, 16)])
def test_sharded_broad_phase_is_a_partition(sccd, orc, scene_c1, world, max_cells):
    """sccd_set_shard: the ranks' pair lists are disjoint and their rank-order concatenation is
    the single-GPU deterministic list -- by cell range (enough cells), or by owner slices of the
    replicated list (max_cells = 1 / few cells)."""
    s = scene_c1
    c = sccd.Context(0)
    c.set_grid_cells(max_cells)
    c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    c.build_boxes(0.0)
    full = [c.broad_phase(0), c.broad_phase(1)]
    n_full = c.stats()["n_records"]
    parts, recs = [[], []], [0, 0]
    for r in range(world):
        c.set_shard(r, world)
        for k in (0, 1):
            parts[k].append(c.broad_phase(k))
        c.build_boxes(0.0)                           # records this rank makes when sharded
        st = c.stats()["n_records"]
        recs = [recs[0] + st[0], recs[1] + st[1]]
    c.close()
    for k in (0, 1):
        cat = np.concatenate(parts[k])
        assert np.array_equal(cat, full[k])          # rank order == global deterministic order
        if world <= 4:
            assert min(len(p) for p in parts[k]) > 0
    if max_cells == 0:      # cell-range sharding: the records are partitioned, not replicated
        assert recs == n_full
    if max_cells == 1:      # owner slices: every rank holds the whole list
        assert recs == [world * n_full[0], world * n_full[1]]


def test_sharded_caller_made_boxes(sccd, orc, scene_c1):
    vb, eb, fb = orc.build_boxes(scene_c1, 1e-3)
    c = sccd.Context(0)
    c.set_boxes(vb, fb, 0)
    full = c.broad_phase(sccd.capi.BOXES)
    parts = []
    for r in range(4):
        c.set_shard(r, 4)
        parts.append(c.broad_phase(sccd.capi.BOXES))
    c.close()
    assert np.array_equal(np.concatenate(parts), full)
    assert min(len(p) for p in parts) > 0


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_caller_made_boxes_any_axis(ctx, sccd, orc, scene_small, axis):
    """BroadPhase::build(boxes[, boxesB]) / sort_and_sweep(boxes..., axis): the overlap set does
    not depend on the sweep axis and the next-axis report matches the reference CPU path."""
    vb, eb, fb = orc.build_boxes(scene_small, 1e-3)
    B = sccd.capi.BOXES
    nxt = ctx.set_boxes(eb, None, axis)
    ee = ctx.broad_phase(B)
    oee, oax = orc.sort_and_sweep(eb, axis)
    assert nxt == oax and len(ee) == len(oee)
    assert np.array_equal(orc.canonical(ee), orc.canonical(oee))
    nxt = ctx.set_boxes(vb, fb, axis)
    vf = ctx.broad_phase(B)
    ovf, oax = orc.sort_and_sweep_two_lists(vb, fb, axis)
    assert nxt == oax and len(vf) == len(ovf)
    assert np.array_equal(orc.canonical(vf), orc.canonical(ovf))
    if orc.ref_cpu() is not None and axis == 0:   # live unmodified reference, when it travels
        ref = orc.ref_cpu_broad_phase(scene_small, r=1e-3, axis=axis)
        assert np.array_equal(orc.canonical(vf), orc.canonical(ref["vf"]))
        assert np.array_equal(orc.canonical(ee), orc.canonical(ref["ee"]))


@pytest.mark.parametrize("max_cells", [1, 2, 7, 64, 0])
def test_cell_grid_does_not_change_the_overlap_set(sccd, orc, scene_c1, max_cells):
    """(y, z) cell striping only prunes candidates: same set for any grid, no duplicates even
    though boxes are replicated into every cell they touch."""
    c = sccd.Context(0)
    c.set_grid_cells(max_cells)
    s = scene_c1
    c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    c.build_boxes(2e-3)
    vf, ee = c.broad_phase(0), c.broad_phase(1)
    st = c.stats()
    c.close()
    vb, eb, fb = orc.build_boxes(s, 2e-3)
    ovf = orc.sort_and_sweep_two_lists(vb, fb, 0)[0]
    oee = orc.sort_and_sweep(eb, 0)[0]
    assert len(vf) == len(ovf) and len(ee) == len(oee)
    assert np.array_equal(orc.canonical(vf), orc.canonical(ovf))
    assert np.array_equal(orc.canonical(ee), orc.canonical(oee))
    cells = [a * b for a, b in st["grid_cells"]]
    if max_cells == 1:
        assert cells == [1, 1] and st["n_records"] == st["n_boxes"]
    elif max_cells == 0:
        assert min(cells) > 16 and st["n_records"][0] <= 2.5 * st["n_boxes"][0] + 1024
    else:
        assert max(cells) <= max_cells


def test_update_vertices_equals_a_fresh_upload(ctx, sccd, orc, scene_small):
    """Frame-to-frame reuse: new positions for the uploaded topology give exactly what a
    fresh upload of the whole mesh gives (host arrays and device pointers)."""
    import torch
    s = scene_small
    s2 = sccd.scenes.cloth_on_sphere(31, seed=8, sphere="uv")      # same topology, other frame
    assert np.array_equal(s["E"], s2["E"]) and np.array_equal(s["F"], s2["F"])
    want = orc.ccd(s2)
    ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    assert ctx.ccd() == orc.ccd(s)["toi"]
    ctx.update_vertices(s2["V0"], s2["V1"])
    assert ctx.ccd() == want["toi"]
    assert ctx.stats()["n_pairs"] == [len(want["vf"]), len(want["ee"])]
    assert np.array_equal(orc.canonical(ctx.broad_phase(0)), want["vf"])
    # device pointers: back to the first frame
    d0 = torch.from_numpy(np.ascontiguousarray(s["V0"].T)).cuda()   # column-major bytes
    d1 = torch.from_numpy(np.ascontiguousarray(s["V1"].T)).cuda()
    ctx.update_vertices(d0.data_ptr(), d1.data_ptr(), nV=s["V0"].shape[0])
    assert ctx.ccd() == orc.ccd(s)["toi"]
    with pytest.raises(sccd.SccdError) as e:
        ctx.update_vertices(s["V0"][:-1].copy(order="F"), s["V1"][:-1].copy(order="F"))
    assert e.value.code == sccd.capi.ERR_ARG
    fresh = sccd.Context(0)
    with pytest.raises(sccd.SccdError) as e:
        fresh.update_vertices(s["V0"], s["V1"])
    assert e.value.code == sccd.capi.ERR_STATE
    fresh.close()
    ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])               # leave host-owned buffers behind


@pytest.mark.parametrize("axis", [1, 2, -1])
def test_pipeline_sweep_axis_changes_no_result(sccd, orc, scene_c1, axis):
    """SCCD_OPT_SWEEP_AXIS (sort_and_sweep.cpp:176-195): boxes, overlap sets and
    every TOI are those of the x sweep."""
    s = scene_c1
    want = orc.ccd(s)
    c = sccd.Context(0)
    try:
        c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        c.set_option(sccd.capi.OPT_SWEEP_AXIS, axis)
        assert c.get_option(sccd.capi.OPT_SWEEP_AXIS) == axis
        for _ in range(2):       # -1: the second call sweeps along the axis the first handed back
            assert c.ccd() == want["toi"]
            assert c.stats()["n_pairs"] == [len(want["vf"]), len(want["ee"])]
        vb, eb, fb = orc.build_boxes(s, 0.0)
        for got, exp in zip(c.get_boxes(), (vb, eb, fb)):
            assert got.tobytes() == exp.tobytes()
        assert np.array_equal(orc.canonical(c.broad_phase(0)), want["vf"])
        assert np.array_equal(orc.canonical(c.broad_phase(1)), want["ee"])
        if axis == -1:
            # the cloth lies in the xy plane: the variance argmax is x or y, never z
            c.build_boxes(0.0)
            assert c.broad_phase(1, want_pairs=False) == len(want["ee"])
    finally:
        c.close()


def test_bad_indices_are_errors_not_faults(sccd, scene_small, torch_cuda):
    """Out-of-range E / F entries and pair ids are SCCD_ERR_ARG (the reference would read out of
    bounds, aabb.cu:199-226 / narrow_phase.cu:38-60); the context stays usable."""
    s = scene_small
    c = sccd.Context(0)
    try:
        E = s["E"].copy(order="F")
        E[5, 1] = s["V0"].shape[0] + 7
        c.upload_mesh(s["V0"], s["V1"], E, s["F"])
        with pytest.raises(sccd.SccdError) as e:
            c.build_boxes(0.0)
        assert e.value.code == sccd.capi.ERR_ARG
        F = s["F"].copy(order="F")
        F[0, 2] = -1
        c.upload_mesh(s["V0"], s["V1"], s["E"], F)
        with pytest.raises(sccd.SccdError) as e:
            c.ccd()
        assert e.value.code == sccd.capi.ERR_ARG
        c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        good = c.ccd()
        pairs = torch_cuda.tensor([[0, 1], [3, s["F"].shape[0]]], dtype=torch_cuda.int32, device="cuda")
        with pytest.raises(sccd.SccdError) as e:
            c.narrow_phase(0, pairs.data_ptr(), 2)
        assert e.value.code == sccd.capi.ERR_ARG
        assert c.ccd() == good
    finally:
        c.close()


def test_max_iter_mode_drop_is_the_reference_rule(sccd, orc, torch_cuda):
    """SCCD_OPT_MAX_ITER_MODE: 0 accepts the boxes of a query that reached max_iter at their t_lo
    (never later than the exact answer); 1 drops them like the reference (root_finder.cu:303-305)
    -- never earlier than the exact answer, may miss; queries under the cap are exact in both."""
    ee, vf = sccd.scenes.queries_c5(4000, seed=12)
    c = sccd.Context(0)
    try:
        for kind, q in ((0, vf), (1, ee)):
            m = orc.tractable(q, kind == 0, 0.0, 1e-6)
            q = np.ascontiguousarray(q[m])
            _, full, _ = orc.narrow_phase(q, kind == 0)
            out = {}
            for mode in (0, 1):
                c.set_option(sccd.capi.OPT_MAX_ITER_MODE, mode)
                c.reset_stats()
                _, tpq = _narrow_gpu(c, torch_cuda, kind, q, max_iter=40)
                ptr, n = c.narrow_phase_checks()
                checks = torch_cuda.as_tensor(
                    sccd.multigpu._DevArray(ptr, (n,), "<u4"), device="cuda").cpu().numpy()
                out[mode] = (tpq, checks <= 41, c.stats()["n_capped"][kind])
            c.set_option(sccd.capi.OPT_MAX_ITER_MODE, 0)
            for mode in (0, 1):
                tpq, under, n_capped = out[mode]
                assert n_capped > 0 and (~under).sum() >= n_capped
                assert np.array_equal(tpq[under], full[under])
            # (depth-first, earliest-t-first order usually reaches the earliest root before the
            # cap, so dropping the rest rarely loses it -- but it can; accepting never does)
            assert np.all(out[0][0] <= full) and np.all(out[1][0] >= full)
    finally:
        c.close()


@pytest.fixture(params=[1, 4, 8])
def group_ctx(request, sccd):
    """A context whose solver is not the default (work queue for short lists): rounds for short
    lists too (1), or the group solver with 4 / 8 lanes per tree."""
    c = sccd.Context(0)
    c.set_option(sccd.capi.OPT_NARROW_SOLVER, request.param)
    yield c
    c.close()


@pytest.mark.parametrize("case", list(CASES))
def test_group_solver_matches_oracle_on_adversarial_queries(group_ctx, orc, sccd, torch_cuda, case):
    """SCCD_OPT_NARROW_SOLVER = 4 / 8 lanes per tree: per-query hit / miss and TOI bit-equal to the
    oracle, shared-bound mode returns the same minimum -- cull on and off, tiny budgets, paths
    deeper than the tracked depth, a 64-entry item list."""
    K = sccd.capi
    c = group_ctx
    ee, vf = sccd.scenes.queries_c5(3000, seed=4)
    kw = CASES[case]
    for kind, q in ((0, vf), (1, ee)):
        q = q[orc.tractable(q, kind == 0, kw["ms"], kw["tol"], kw["allow_zero_toi"])]
        otoi, otpq, _ = orc.narrow_phase(q, kind == 0, kw["ms"], kw["max_iter"], kw["tol"],
                                         kw["allow_zero_toi"])
        variants = [(1, 0, 128, 0), (0, 0, 128, 0), (1, 4 | (3 << 8) | (2 << 16), 128, 0)]
        if case in ("default", "nozero"):
            # (tol 1e-9 / ms > 0 make trees of 10^4 boxes: cut every 6 levels they outgrow the
            # item list, and a 64-entry list cannot take the hand-on of a path deeper than 128
            # levels -- both documented errors, not wrong answers)
            variants += [(1, 0, 6, 0), (1, 8 | (5 << 8), 128, 64)]
        for cull, flags, depth, qcap in variants:
            c.set_option(K.OPT_NARROW_CULL, cull)
            c.set_option(K.OPT_NARROW_FLAGS, flags)
            c.set_option(K.OPT_NARROW_MAX_DEPTH, depth)
            c.set_queue_capacity(qcap)
            toi, tpq = _narrow_gpu(c, torch_cuda, kind, q, **kw)
            assert np.array_equal(tpq, otpq), (kind, cull, flags, depth, qcap)
            assert toi == otoi
            assert c.narrow_phase_queries(kind, q, **kw) == otoi
    c.set_option(K.OPT_NARROW_CULL, 1)
    c.set_option(K.OPT_NARROW_FLAGS, 0)
    c.set_option(K.OPT_NARROW_MAX_DEPTH, 128)
    c.set_queue_capacity(0)


def test_group_solver_pipeline_and_cap(group_ctx, orc, sccd, torch_cuda, scene_c1):
    c = group_ctx
    s = scene_c1
    want = orc.ccd(s)
    c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    assert c.ccd() == want["toi"]
    toi, (vf_ids, vf_t), (ee_ids, ee_t) = c.ccd_collisions()
    assert toi == want["toi"]
    for ids, t, pairs, tq in ((vf_ids, vf_t, want["vf"], want["toi_vf"]),
                              (ee_ids, ee_t, want["ee"], want["toi_ee"])):
        order = np.lexsort((ids[:, 1], ids[:, 0]))
        hit = tq < 1
        assert np.array_equal(ids[order], pairs[hit]) and np.array_equal(t[order], tq[hit])
    # float build of the reference
    c.set_scalar_type(sccd.capi.F32)
    want32 = orc.ccd(s, f32=True)
    assert c.ccd() == want32["toi"]
    c.set_scalar_type(sccd.capi.F64)
    # iteration cap: conservative, exact under the cap
    ee, vf = sccd.scenes.queries_c5(2000, seed=12)
    for kind, q in ((0, vf), (1, ee)):
        _, full, _ = orc.narrow_phase(q, kind == 0, max_iter=-1)
        _, capped = _narrow_gpu(c, torch_cuda, kind, q, max_iter=60)
        ptr, n = c.narrow_phase_checks()
        checks = torch_cuda.as_tensor(sccd.multigpu._DevArray(ptr, (n,), "<u4"), device="cuda").cpu().numpy()
        under = checks <= 61
        assert np.all(capped <= full) and np.array_equal(capped[under], full[under])
        assert c.stats()["n_capped"][kind] > 0


def test_time_ordered_solve_skips_work_but_changes_no_result(sccd, orc, scene_c1):
    """The cull attaches a lower bound of the time of impact to every surviving query
    (tests/test_cull_math.py: toi_lower_bound); round 0 solves the survivors earliest-possible-
    contact first, a scout launch first of all, and skips queries whose bound is not below the
    running earliest toi.  Same TOI with every part of it switched off; far fewer box checks
    with it on."""
    K = sccd.capi
    s = scene_c1
    want = orc.ccd(s)
    c = sccd.Context(0)
    try:
        c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        c.ccd()      # (the work queue's grid follows the previous batch: same for every variant)
        checks = {}
        for name, flags in (("on", 0), ("no_scout", 1 << 6), ("no_skip", 1 << 7),
                            ("unsorted", 1 << 23), ("all_off", (1 << 6) | (1 << 7) | (1 << 23))):
            c.set_option(K.OPT_NARROW_FLAGS, flags)
            assert c.ccd() == want["toi"], name
            st = c.stats()
            checks[name] = sum(st["n_box_checks"])
            if name == "on":
                assert sum(st["n_skipped"]) > 0
            if name in ("no_skip", "all_off"):
                assert st["n_skipped"] == [0, 0]
        assert checks["on"] < checks["all_off"] and checks["on"] <= checks["no_skip"], checks
        # the per-query list (TOI_PER_QUERY) never skips
        c.set_option(K.OPT_NARROW_FLAGS, 0)
        toi, (vi, vt), (ei, et) = c.ccd_collisions()
        assert toi == want["toi"] and c.stats()["n_skipped"] == [0, 0]
        assert len(vi) == int((want["toi_vf"] < 1).sum()) and len(ei) == int((want["toi_ee"] < 1).sum())
        # ... nor does a run with an iteration cap
        c.ccd(max_iter=1000)
        assert c.stats()["n_skipped"] == [0, 0]
    finally:
        c.close()


@pytest.mark.parametrize("max_cells", [0, 1])
def test_tma_staged_sweep_gives_the_same_list(sccd, scene_c1, scene_small, max_cells):
    """SCCD_OPT_SWEEP_STAGED: the count pass reads its window from shared memory filled by bulk
    async copies -- same pairs in the same order, also for owner slices that start at unaligned
    records (one-axis sweep, 3 shards) and for tiles whose windows outrun the staged records."""
    K = sccd.capi
    c = sccd.Context(0)
    try:
        c.set_grid_cells(max_cells)
        for s in (scene_small, scene_c1):
            c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
            for world in (1, 3):
                lists = {}
                for staged in (0, 1):
                    c.set_option(K.OPT_SWEEP_STAGED, staged)
                    parts = []
                    for r in range(world):
                        c.set_shard(r, world)
                        c.build_boxes(0.0)
                        parts.append([c.broad_phase(0), c.broad_phase(1)])
                    lists[staged] = parts
                c.set_shard(0, 1)
                for r in range(world):
                    for k in (0, 1):
                        assert np.array_equal(lists[0][r][k], lists[1][r][k])
        c.set_option(K.OPT_SWEEP_STAGED, 1)
        toi1 = c.ccd()
        c.set_option(K.OPT_SWEEP_STAGED, 0)
        assert c.ccd() == toi1
    finally:
        c.close()


def test_frame_to_frame_grid_reuse(sccd, orc, scene_small):
    """SCCD_OPT_REUSE_GRID: a build with the list sizes of the previous one takes its
    cell grid and key quantisation from the PREVIOUS build's statistics (one host sync less); the
    mesh may have moved, stretched or left the old bounding box altogether -- same pairs, same
    TOI.  The first build after an upload of another size waits for its own statistics again."""
    K = sccd.capi
    s = scene_small
    c = sccd.Context(0)
    fresh = sccd.Context(0)
    fresh.set_option(K.OPT_REUSE_GRID, 0)
    try:
        assert c.get_option(K.OPT_REUSE_GRID) == 1
        c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        toi = c.ccd()
        first = c.stats()["n_host_syncs"]
        assert toi == orc.ccd(s)["toi"]
        c.ccd()
        again = c.stats()["n_host_syncs"]
        assert again == first - 1, (first, again)
        # frames that leave the statistics of the previous build far behind
        frames = [(1.0, np.zeros(3)), (1.0, np.array([50.0, -20.0, 3.0])), (7.5, np.array([-3.0, 0.5, 9.0])),
                  (0.01, np.array([1e3, 1e3, 1e3])), (1.0, np.zeros(3))]
        for scale, shift in frames:
            V0 = np.asfortranarray(s["V0"] * scale + shift)
            V1 = np.asfortranarray(s["V1"] * scale + shift)
            c.update_vertices(V0, V1)
            got = c.ccd()
            assert c.stats()["n_host_syncs"] == again
            fresh.upload_mesh(V0, V1, s["E"], s["F"])
            assert fresh.ccd() == got
            assert fresh.stats()["n_pairs"] == c.stats()["n_pairs"]
            for k in (0, 1):
                assert np.array_equal(orc.canonical(c.broad_phase(k)), orc.canonical(fresh.broad_phase(k)))
        # another mesh size: no stale statistics
        s2 = sccd.scenes.cloth_on_sphere(17, seed=3, sphere="uv")
        c.upload_mesh(s2["V0"], s2["V1"], s2["E"], s2["F"])
        assert c.ccd() == orc.ccd(s2)["toi"]
        assert c.stats()["n_host_syncs"] == first
        # the option off: every build waits for its own statistics
        c.set_option(K.OPT_REUSE_GRID, 0)
        c.ccd()
        assert c.stats()["n_host_syncs"] == first
    finally:
        c.close()
        fresh.close()


def test_solver_launches_follow_the_previous_batch(sccd, scene_small):
    """Which solver kernels a batch needs depends on the length of its survivor list, known on
    the device only; the host launches what the previous batch of that kind needed and launches
    the rest when the length class changed (sccd_stats.n_relaunched).  Alternating a cloth scene
    (hundreds of survivors: work queue) with a pile (tens of thousands: rounds) must give what a
    context that always launches everything gives, collisions included."""
    K = sccd.capi
    pile = sccd.scenes.blob_pile(1000, seed=2)
    a, b = sccd.Context(0), sccd.Context(0)
    b.set_option(K.OPT_REUSE_GRID, 0)              # no frame-to-frame guesses at all
    for c in (a, b):
        c.set_option(K.OPT_NARROW_CULL, 2)         # (keeps the pile's survivor lists long)
    try:
        relaunched = 0
        for s in (scene_small, scene_small, pile, pile, scene_small, pile, pile):
            for c in (a, b):
                c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
            ta, tb = a.ccd(), b.ccd()
            assert ta == tb
            sa, sb = a.stats(), b.stats()
            assert sa["n_pairs"] == sb["n_pairs"] and sa["n_culled"] == sb["n_culled"]
            assert sb["n_relaunched"] == 0
            relaunched += sa["n_relaunched"]
            ca, cb = a.ccd_collisions(), b.ccd_collisions()
            assert ca[0] == cb[0] == ta
            for k in (1, 2):
                oa, ob = np.lexsort(ca[k][0].T[::-1]), np.lexsort(cb[k][0].T[::-1])
                assert np.array_equal(ca[k][0][oa], cb[k][0][ob])
                assert np.array_equal(ca[k][1][oa], cb[k][1][ob])
        assert relaunched >= 3                      # small -> pile, pile -> small, small -> pile
        # the same scene again: the guess holds
        a.ccd()
        assert a.stats()["n_relaunched"] == 0
    finally:
        a.close()
        b.close()


def test_profile_modes_change_no_result(sccd, orc, scene_c1):
    """SCCD_OPT_PROFILE 1 (event pair per kernel) and 2 (the same with both lists on the caller's
    stream, so that a pair brackets its kernel alone): same TOI and pair counts as the untimed
    pipeline, kernel times reported; an out-of-range value is an error, not an abort."""
    K = sccd.capi
    s = scene_c1
    want = orc.ccd(s)
    c = sccd.Context(0)
    try:
        c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        assert c.ccd() == want["toi"]
        base = c.stats()["n_pairs"]
        for mode in (1, 2, 0, 2, 1, 0):
            c.set_option(K.OPT_PROFILE, mode)
            for _ in range(2):
                assert c.ccd() == want["toi"]
            st = c.stats()
            assert st["n_pairs"] == base
            if mode:
                assert st["ms_k_boxes"] > 0 and min(st["ms_k_cull"]) > 0 and st["ms_k_gather"] > 0
                assert st["ms_k_round"][0][0] > 0 and st["ms_k_round"][1][0] > 0
        for bad_opt, bad in ((K.OPT_PROFILE, 3), (K.OPT_NARROW_CULL, 5), (K.OPT_NARROW_SOLVER, 3)):
            with pytest.raises(sccd.SccdError) as e:
                c.set_option(bad_opt, bad)
            assert e.value.code == K.ERR_ARG
        assert c.ccd() == want["toi"]
    finally:
        c.close()
