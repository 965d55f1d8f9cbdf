"""The reference's float build (SCALABLE_CCD_USE_DOUBLE off, scalar.hpp:16-18) as a run-time
mode of the library: sccd_set_scalar_type(SCCD_F32).

CPU part: the float oracle (oracle/liborc_f32.so) against goldens of the UNMODIFIED reference
CPU sources compiled without SCALABLE_CCD_USE_DOUBLE (oracle/Makefile ref_f32) and, live, against
that library when oracle/_ref is present.
GPU part: the CUDA path in SCCD_F32 mode against the float oracle (boxes and overlap sets
bit-exact) and against per-query TOIs frozen from the unmodified reference CUDA sources in
their float build (tests/golden/*_f32*, generator tests/golden/make_golden.py cuda_f32)."""
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


@pytest.fixture(scope="module")
def gold_f32():
    return json.load(open(os.path.join(GOLD, "broad_ref_cpu_f32.json")))


def is_float_valued(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array_equal(a.astype(np.float32).astype(np.float64), a)


# ------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("name", ["small", "c1", "pile", "slab"])
def test_float_oracle_broad_phase_matches_reference_golden(orc, sccd, scene_c1, scene_small,
                                                           gold_f32, name):
    from test_oracle import golden_scene
    s = golden_scene(sccd, scene_c1, scene_small, name)
    g = gold_f32[name]
    vb, eb, fb = orc.build_boxes(s, f32=True)
    assert sha(vb, eb, fb) == g["boxes_sha256"]                      # boxes bit-exact
    assert all(is_float_valued(b[f]) for b in (vb, eb, fb) for f in ("min", "max"))
    assert sha(*orc.build_boxes(s, 1e-3, f32=True)) == g["boxes_r1e-3_sha256"]
    vf, ax_vf = orc.sort_and_sweep_two_lists(vb, fb, 0, f32=True)
    ee, ax_ee = orc.sort_and_sweep(eb, 0, f32=True)
    assert (len(vf), len(ee)) == (g["n_vf"], g["n_ee"])
    assert sha(orc.canonical(vf)) == g["vf_sha256"]
    assert sha(orc.canonical(ee)) == g["ee_sha256"]
    assert [ax_vf, ax_ee] == g["next_axes"]


def test_float_boxes_differ_from_double_and_contain_the_motion(orc, scene_small):
    d = orc.build_boxes(scene_small)[0]
    f = orc.build_boxes(scene_small, f32=True)[0]
    assert d.tobytes() != f.tobytes()
    V0, V1 = scene_small["V0"], scene_small["V1"]
    lo, hi = np.minimum(V0, V1), np.maximum(V0, V1)
    # conservative in float as well: the cast moves a coordinate by <= 1/2 ulp, nextafterf by 1
    assert np.all(f["min"] < lo) and np.all(f["max"] > hi)


def test_float_oracle_against_live_reference_cpu(orc, scene_small):
    if orc.ref_cpu(f32=True) is None:
        pytest.skip("oracle/_ref float build not present (make -C oracle ref_f32)")
    for r in (0.0, 2e-3):
        ref = orc.ref_cpu_broad_phase(scene_small, r=r, f32=True)
        mine = orc.build_boxes(scene_small, r=r, f32=True)
        for a, b in zip(mine, orc.ref_cpu_build_boxes(scene_small, r=r, f32=True)):
            assert a.tobytes() == b.tobytes()
        vf, a1 = orc.sort_and_sweep_two_lists(mine[0], mine[2], 0, f32=True)
        ee, a2 = orc.sort_and_sweep(mine[1], 0, f32=True)
        assert np.array_equal(orc.canonical(vf), orc.canonical(ref["vf"]))
        assert np.array_equal(orc.canonical(ee), orc.canonical(ref["ee"]))
        assert (a1, a2) == tuple(ref["axes"])


def vf_query(p0, p1, tri):
    return np.concatenate([p0, *tri, p1, *tri]).astype(np.float64)


TRI = [np.array([0., 0., 0.]), np.array([1., 0., 0.]), np.array([0., 1., 0.])]


def test_float_oracle_narrow_phase_known_answers(orc):
    q = vf_query([0.25, 0.25, 1.0], [0.25, 0.25, -1.0], TRI)
    toi, tpq, _ = orc.narrow_phase(q[None], True, tol=1e-6, f32=True)
    assert tpq[0] == toi and toi <= 0.5 and 0.5 - toi < 1e-4 and is_float_valued(toi)
    miss = vf_query([2.0, 2.0, 1.0], [2.0, 2.0, -1.0], TRI)
    toi, tpq, _ = orc.narrow_phase(miss[None], True, f32=True)
    assert toi == 1.0 and np.isinf(tpq[0])
    # resting contact: allow_zero_toi
    rest = vf_query([0.25, 0.25, 0.0], [0.3, 0.25, 0.0], TRI)
    assert orc.narrow_phase(rest[None], True, f32=True)[0] == 0.0


def test_float_and_double_oracles_agree_within_the_float_error_filter(orc, scene_small):
    d = orc.ccd(scene_small)
    f = orc.ccd(scene_small, f32=True)
    # the float build resolves contact up to its error filter (3.6e-6 max^3, root_finder.cu:102-119)
    # and its coarser parameter grid: earlier (more conservative) or equal within 1e-3
    assert f["toi"] <= d["toi"] + 1e-6 and d["toi"] - f["toi"] < 1e-3
    assert is_float_valued(f["toi"])


def test_float_flush_to_zero_and_tiny_inputs(orc):
    # subnormal-in-float coordinates are flushed by the first device operation: the query
    # behaves as if they were exact zeros
    q = vf_query([0.25, 0.25, 1.0], [0.25, 0.25, -1.0], TRI)
    q2 = q.copy()
    q2[3:6] += 1e-41                       # below FLT_MIN (1.18e-38): subnormal as float
    a = orc.narrow_phase(q[None], True, f32=True)[0]
    b = orc.narrow_phase(q2[None], True, f32=True)[0]
    assert a == b


CASES_F32 = {"default": dict(ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True),
             "loose": dict(ms=0.0, max_iter=-1, tol=1e-4, allow_zero_toi=True),
             "ms": dict(ms=1e-5, max_iter=-1, tol=1e-6, allow_zero_toi=True),
             "nozero": dict(ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=False)}


def test_float_oracle_matches_reference_cuda_float_golden_queries(orc, sccd):
    """Float oracle vs per-query TOIs frozen from the unmodified reference CUDA float build.
    The oracle's reciprocal is the correctly rounded one, not MUFU.RCP (see its header), so
    equality is not guaranteed in general -- on these fixtures it holds bit for bit."""
    z = np.load(os.path.join(GOLD, "narrow_c5_ref_cuda_f32.npz"))
    sel = np.load(os.path.join(GOLD, "narrow_c5_f32_idx.npz"))
    ee, vf = sccd.scenes.queries_c5(3000, seed=4)
    assert sha(ee, vf) == bytes(sel["c5_sha256"]).hex()
    n_hits = 0
    for case, kw in CASES_F32.items():
        for name, q in (("vf", vf), ("ee", ee)):
            idx = sel[f"{case}_{name}_idx"]
            toi, tpq, _ = orc.narrow_phase(q[idx], name == "vf", kw["ms"], kw["max_iter"],
                                           kw["tol"], kw["allow_zero_toi"], f32=True)
            ref = z[f"{case}_{name}_tpq"]
            assert np.array_equal(tpq < 1, ref < 1), (case, name)
            assert np.array_equal(tpq, ref), (case, name)
            assert toi == float(z[f"{case}_{name}_toi"])
            n_hits += int((ref < 1).sum())
    assert n_hits > 150


def test_float_oracle_matches_reference_cuda_float_golden_pipeline(orc, scene_c1):
    """config 1 in float: 1,489 collisions with their TOIs, from the reference's float broad
    phase + float root finder."""
    z = np.load(os.path.join(GOLD, "ccd_c1_ref_cuda_f32.npz"))
    r = orc.ccd(scene_c1, f32=True)
    assert r["toi"] == float(z["toi"])
    hv, he = r["toi_vf"] < 1, r["toi_ee"] < 1
    nv = int(z["n_vf_hits"])
    assert (int(hv.sum()), int(he.sum())) == (nv, len(z["coll_ids"]) - nv) and nv > 100
    assert np.array_equal(r["vf"][hv], z["coll_ids"][:nv])       # both in canonical order
    assert np.array_equal(r["ee"][he], z["coll_ids"][nv:])
    assert np.array_equal(r["toi_vf"][hv], z["coll_toi"][:nv])
    assert np.array_equal(r["toi_ee"][he], z["coll_toi"][nv:])


# ------------------------------------------------------------------------------- GPU
@pytest.fixture()
def ctx32(ctx, sccd):
    ctx.set_scalar_type(sccd.capi.F32)
    yield ctx
    ctx.set_scalar_type(sccd.capi.F64)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch


@pytest.mark.gpu
@pytest.mark.parametrize("r", [0.0, 1e-3])
def test_gpu_float_boxes_bit_exact(ctx32, orc, scene_c1, gold_f32, r):
    s = scene_c1
    ctx32.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    ctx32.build_boxes(r)
    got = ctx32.get_boxes()
    for g, w in zip(got, orc.build_boxes(s, r, f32=True)):
        assert g.tobytes() == w.tobytes()
    key = "boxes_sha256" if r == 0.0 else "boxes_r1e-3_sha256"
    assert sha(*got) == gold_f32["c1"][key]          # unmodified reference CPU, float build


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["small", "c1"])
def test_gpu_float_overlap_sets_bit_exact(ctx32, orc, scene_c1, scene_small, gold_f32, which):
    s = {"small": scene_small, "c1": scene_c1}[which]
    ctx32.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    ctx32.build_boxes(0.0)
    vf, ee = ctx32.broad_phase(0), ctx32.broad_phase(1)
    g = gold_f32[which]
    assert (len(vf), len(ee)) == (g["n_vf"], g["n_ee"])
    assert sha(orc.canonical(vf)) == g["vf_sha256"] and sha(orc.canonical(ee)) == g["ee_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("f32", [False, True])
def test_gpu_named_box_builders(ctx, sccd, orc, scene_small, f32):
    """build_vertex_boxes / build_edge_boxes / build_face_boxes (aabb.cuh:150-188) by name."""
    s = scene_small
    ctx.set_scalar_type(sccd.capi.F32 if f32 else sccd.capi.F64)
    try:
        for r in (0.0, 2e-3):
            vb = ctx.build_vertex_boxes(s["V0"], s["V1"], r)
            eb = ctx.build_element_boxes(vb, s["E"])
            fb = ctx.build_element_boxes(vb, s["F"])
            for g, w in zip((vb, eb, fb), orc.build_boxes(s, r, f32=f32)):
                assert g.tobytes() == w.tobytes()
        # single-frame overload = the box of a point
        one = ctx.build_vertex_boxes(s["V0"], None, 0.0)
        two = ctx.build_vertex_boxes(s["V0"], s["V0"], 0.0)
        assert one.tobytes() == two.tobytes()
        # caller-made boxes go straight into the sweep (tests/test_broad_phase.cu:93-104)
        ctx.set_boxes(vb, fb)
        vf = ctx.broad_phase(sccd.capi.BOXES)
        ovf, _ = orc.sort_and_sweep_two_lists(vb, fb, 0, f32=f32)
        assert np.array_equal(orc.canonical(vf), orc.canonical(ovf))
        with pytest.raises(sccd.SccdError) as e:
            bad = s["E"].copy(order="F")
            bad[0, 0] = len(vb)
            ctx.build_element_boxes(vb, bad)
        assert e.value.code == sccd.capi.ERR_ARG
        assert len(ctx.build_element_boxes(vb, np.zeros((0, 2), np.int32, order="F"))) == 0
    finally:
        ctx.set_scalar_type(sccd.capi.F64)


def _narrow_gpu(ctx, torch, kind, q, **kw):
    tq = torch.empty(len(q), dtype=torch.float64, device="cuda")
    toi = ctx.narrow_phase_queries(kind, q, d_toi_per_query=tq.data_ptr(), **kw)
    return toi, tq.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES_F32))
def test_gpu_float_queries_match_reference_cuda_float_build(ctx32, sccd, torch_cuda, case):
    """Per-query hit / miss and TOI bit-equal to the UNMODIFIED reference CUDA kernels compiled
    with SCALABLE_CCD_USE_DOUBLE off (and --use_fast_math, as its CMake does), run on a B200."""
    path = os.path.join(GOLD, "narrow_c5_ref_cuda_f32.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet (tests/golden/make_golden.py cuda_f32)")
    z = np.load(path)
    sel = np.load(os.path.join(GOLD, "narrow_c5_f32_idx.npz"))
    ee, vf = sccd.scenes.queries_c5(3000, seed=4)
    assert sha(ee, vf) == bytes(sel["c5_sha256"]).hex()
    kw = CASES_F32[case]
    for kind, name, q in ((0, "vf", vf), (1, "ee", ee)):
        idx = sel[f"{case}_{name}_idx"]
        assert len(idx) > 100
        toi, tpq = _narrow_gpu(ctx32, torch_cuda, kind, q[idx], **kw)
        ref = z[f"{case}_{name}_tpq"]
        n_bad = int((tpq != ref).sum())
        assert np.array_equal(tpq < 1, ref < 1), (case, name, int(((tpq < 1) != (ref < 1)).sum()))
        assert n_bad == 0, (case, name, n_bad, float(np.nanmax(np.abs(
            np.where(np.isfinite(tpq) & np.isfinite(ref), tpq - ref, 0.0)))))
        assert toi == float(z[f"{case}_{name}_toi"])
        assert ctx32.narrow_phase_queries(kind, q[idx], **kw) == toi      # shared-bound build


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["default", "ms"])
def test_gpu_float_queries_against_float_oracle(ctx32, orc, sccd, torch_cuda, case):
    """The float oracle models the hardware reciprocal with the correctly rounded one (its
    header says so): hit / miss must agree, TOIs may differ where a tolerance or a split
    choice moved by an ulp -- by no more than the query's own resolution."""
    sel = np.load(os.path.join(GOLD, "narrow_c5_f32_idx.npz"))
    ee, vf = sccd.scenes.queries_c5(3000, seed=4)
    kw = CASES_F32[case]
    for kind, name, q in ((0, "vf", vf), (1, "ee", ee)):
        q = q[sel[f"{case}_{name}_idx"]]
        toi, tpq = _narrow_gpu(ctx32, torch_cuda, kind, q, **kw)
        _, otpq, _ = orc.narrow_phase(q, kind == 0, kw["ms"], kw["max_iter"], kw["tol"],
                                      kw["allow_zero_toi"], f32=True)
        assert is_float_valued(tpq[np.isfinite(tpq)])
        same_hit = (tpq < 1) == (otpq < 1)
        assert same_hit.mean() >= 0.99, (case, name, float(same_hit.mean()))
        both = (tpq < 1) & (otpq < 1)
        if both.any():
            assert np.abs(tpq[both] - otpq[both]).max() <= 1e-3, (case, name)
            assert (tpq[both] == otpq[both]).mean() >= 0.9, (case, name)


@pytest.mark.gpu
def test_gpu_float_pipeline_matches_reference_cuda_float_build(ctx32, orc, scene_c1):
    """ccd() in float: the candidate sets are the reference's, every collision and its TOI equal
    what the reference's float root finder returns for that pair."""
    path = os.path.join(GOLD, "ccd_c1_ref_cuda_f32.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet (tests/golden/make_golden.py cuda_f32)")
    z = np.load(path)
    s = scene_c1
    ctx32.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    toi = ctx32.ccd()
    assert toi == float(z["toi"]) and is_float_valued(toi)
    st = ctx32.stats()
    meta = json.load(open(os.path.join(GOLD, "ref_cuda_f32_meta.json")))
    assert st["n_pairs"] == [meta["broad_c1"]["n_vf"], meta["broad_c1"]["n_ee"]]
    assert min(st["n_culled"]) > 0.8 * min(st["n_pairs"])     # the float cull does its job
    t2, (vf_ids, vf_t), (ee_ids, ee_t) = ctx32.ccd_collisions()
    assert t2 == toi
    nv = int(z["n_vf_hits"])
    for ids, t, rid, rt in ((vf_ids, vf_t, z["coll_ids"][:nv], z["coll_toi"][:nv]),
                            (ee_ids, ee_t, z["coll_ids"][nv:], z["coll_toi"][nv:])):
        o1, o2 = np.lexsort((ids[:, 1], ids[:, 0])), np.lexsort((rid[:, 1], rid[:, 0]))
        assert np.array_equal(ids[o1], rid[o2])
        assert np.array_equal(t[o1], rt[o2])
    # the IPC wrapper stays consistent with it
    t3 = ctx32.ipc_ccd_strategy()
    assert is_float_valued(t3) and (t3 == toi or toi < 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(ms=0.0, tol=1e-6), dict(ms=1e-4, tol=1e-6),
                                dict(ms=0.0, tol=1e-3), dict(ms=1e-6, tol=1e-5)])
def test_gpu_float_cull_changes_no_result(ctx32, sccd, orc, torch_cuda, scene_c1, kw):
    """The float variant of the separating-axis cull: every per-query TOI equals the float
    solver's own (cull off), also with a minimum separation or a tolerance far larger than the
    gaps it tests against."""
    s = scene_c1
    ctx32.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    ctx32.build_boxes(kw["ms"])
    OPT = sccd.capi.OPT_NARROW_CULL
    try:
        for kind in (0, 1):
            pairs = ctx32.broad_phase(kind)
            q = orc.gather_queries(s, np.ascontiguousarray(pairs), kind == 0)
            ctx32.set_option(OPT, 1)
            ctx32.reset_stats()
            toi1, tpq1 = _narrow_gpu(ctx32, torch_cuda, kind, q, max_iter=-1,
                                     allow_zero_toi=True, **kw)
            culled = ctx32.stats()["n_culled"][kind]
            ctx32.set_option(OPT, 0)
            ctx32.reset_stats()
            toi0, tpq0 = _narrow_gpu(ctx32, torch_cuda, kind, q, max_iter=-1,
                                     allow_zero_toi=True, **kw)
            assert ctx32.stats()["n_culled"][kind] == 0
            assert np.array_equal(tpq0, tpq1) and toi0 == toi1
            if kw["ms"] == 0.0 and kw["tol"] == 1e-6:
                assert culled > 0.8 * len(q)
    finally:
        ctx32.set_option(OPT, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [1 << 24, 4 | (3 << 8) | (2 << 16) | (1 << 25), 1 << 28])
def test_gpu_float_scheduling_variants_change_no_result(ctx32, orc, sccd, torch_cuda, scene_c1,
                                                         flags):
    """Lane-per-tree only / tiny budgets / small cooperative limit (SCCD_OPT_NARROW_FLAGS, see
    tests/test_gpu_parity.py VARIANTS): the float kernels return the same per-query TOIs
    however the trees are cut and whichever of the two kernels walks them."""
    s = scene_c1
    ctx32.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    ctx32.build_boxes(0.0)
    ee_q, vf_q = sccd.scenes.queries_c5(1500, seed=4)
    sel = orc.tractable(vf_q, True, 0.0, 1e-6, f32=True)
    OPT = sccd.capi.OPT_NARROW_FLAGS
    try:
        for kind in (0, 1):
            pairs = ctx32.broad_phase(kind)
            qs = [orc.gather_queries(s, np.ascontiguousarray(pairs), kind == 0)]
            if kind == 0:
                qs.append(vf_q[sel])
            for q in qs:
                ctx32.set_option(OPT, 0)
                toi1, tpq1 = _narrow_gpu(ctx32, torch_cuda, kind, q)
                ctx32.set_option(OPT, flags)
                toi2, tpq2 = _narrow_gpu(ctx32, torch_cuda, kind, q)
                assert np.array_equal(tpq1, tpq2) and toi1 == toi2
    finally:
        ctx32.set_option(OPT, 0)


@pytest.mark.gpu
def test_gpu_scalar_type_switch_rebuilds_and_double_is_untouched(ctx, sccd, orc, scene_small):
    s = scene_small
    ctx.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
    t64 = ctx.ccd()
    ctx.set_scalar_type(sccd.capi.F32)
    try:
        with pytest.raises(sccd.SccdError):      # boxes belong to the scalar type
            ctx.broad_phase_begin(0)
        t32 = ctx.ccd()
    finally:
        ctx.set_scalar_type(sccd.capi.F64)
    assert ctx.ccd() == t64 == orc.ccd(s)["toi"]
    assert is_float_valued(t32) and abs(t32 - t64) < 1e-3
    with pytest.raises(sccd.SccdError):
        ctx.set_scalar_type(7)
