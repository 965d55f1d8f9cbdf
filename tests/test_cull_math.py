"""The separating-axis cull of the narrow phase (csrc/narrow.cu: narrow_cull_kernel), restated
in numpy and checked against the oracle on the CPU: whatever it culls, the reference's root
finder answers "no collision" -- for every tolerance / minimum separation, including the ones
far larger than the gaps it tests against, and for the adversarial queries of config 5."""
import numpy as np
import pytest

DIAGS = np.array([[1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1], [0, 1, 1], [0, 1, -1]], float)


def absmax(*pairs):
    return np.max(np.stack([np.abs(b - a) for a, b in pairs]), axis=0)


def cull_mask(q, is_vf, tol, ms, f32=False):
    """numpy mirror of narrow_cull_kernel: True = culled (answered "no collision").
    f32: the variant for the reference's float build (SCCD_F32): inputs as the float solver
    sees them, the float error filters, and a scale test that keeps every tol_k above the
    float resolution of the parameters (so that condition 4 cannot fire)."""
    if f32:
        q = q.astype(np.float32).astype(np.float64)
        tol, ms = float(np.float32(tol)), float(np.float32(ms))
    p = q.reshape(-1, 2, 4, 3)                       # [query, time, vertex, xyz]
    s, e = p[:, 0], p[:, 1]
    L = np.zeros((3, len(p)))
    for k in range(3):
        s0, s1, s2, s3 = (s[:, j, k] for j in range(4))
        e0, e1, e2, e3 = (e[:, j, k] for j in range(4))
        if is_vf:   # root_finder.cu:50-59
            p000, p001, p011, p010 = s0 - s1, s0 - s3, s0 - (s2 + s3 - s1), s0 - s2
            p100, p101, p111, p110 = e0 - e1, e0 - e3, e0 - (e2 + e3 - e1), e0 - e2
        else:       # root_finder.cu:73-80
            p000, p001, p010, p011 = s0 - s2, s0 - s3, s1 - s2, s1 - s3
            p100, p101, p110, p111 = e0 - e2, e0 - e3, e1 - e2, e1 - e3
        L[0] = np.maximum(L[0], absmax((p000, p100), (p001, p101), (p011, p111), (p010, p110)))
        L[1] = np.maximum(L[1], absmax((p000, p010), (p100, p110), (p101, p111), (p001, p011)))
        L[2] = np.maximum(L[2], absmax((p000, p001), (p100, p101), (p110, p111), (p010, p011)))
    if is_vf:
        A = np.stack([s[:, 0], e[:, 0]], axis=1)
        B = np.stack([s[:, 1], s[:, 2], s[:, 3], e[:, 1], e[:, 2], e[:, 3],
                      s[:, 2] + s[:, 3] - s[:, 1], e[:, 2] + e[:, 3] - e[:, 1]], axis=1)
        width = np.full(len(p), tol)
    else:
        A = np.stack([s[:, 0], s[:, 1], e[:, 0], e[:, 1]], axis=1)
        B = np.stack([s[:, 2], s[:, 3], e[:, 2], e[:, 3]], axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            width = np.where((L[0] > 0) & (L[1] > 0),
                             tol * (1 + L[1] / L[0] + L[2] / L[1]) / 3 * 1.000001, np.inf)
        width = np.maximum(width, tol)
    pa = A @ DIAGS.T
    pb = B @ DIAGS.T
    sep = np.maximum(pa.min(1) - pb.max(1), pb.min(1) - pa.max(1)).max(1)
    maxabs = np.maximum(1.0, np.abs(p).reshape(len(p), -1).max(1))
    extent = p.reshape(len(p), -1).max(1) - p.reshape(len(p), -1).min(1)
    if f32:     # 8e-6 >= every float error filter (root_finder.cu:102-119)
        bound = 2.0 * (width + ms + 2.0 * maxabs ** 3 * 8e-6 + 1e-6 * maxabs)
        sane = (extent <= tol * 1e12) & (L.max(0) <= tol * 1e6)     # tol_k >= 3e-7 > 2^-23
    else:
        bound = 2.0 * (width + ms + 2.0 * maxabs ** 3 * 8e-15 + 1e-12 * maxabs)
        sane = extent <= tol * 1e12
    return sane & (0.5 * sep > bound)


def toi_lower_bound(q, is_vf, tol, ms, f32=False):
    """numpy mirror of the time-of-impact lower bound the cull kernel attaches to every
    surviving query (csrc/narrow.cu): no box the root finder ACCEPTS for this query starts before
    it.  Along a coordinate axis k the two primitives are `gap_k` apart at t = 0 and close in by
    at most `D_k` per unit time (largest end-point displacement of either); an accepted box has
    every corner within ms + err + W of the origin in every coordinate (the same W the cull
    uses), so its t_lo is at least (gap_k - B) / D_k.  Used to ORDER the solver's work
    (earliest possible contact first) and, in shared-bound mode without an iteration cap, to
    skip queries that cannot lower the earliest toi."""
    if f32:
        q = q.astype(np.float32).astype(np.float64)
        tol, ms = float(np.float32(tol)), float(np.float32(ms))
    p = q.reshape(-1, 2, 4, 3)
    s, e = p[:, 0], p[:, 1]
    L = np.zeros((3, len(p)))
    for k in range(3):
        s0, s1, s2, s3 = (s[:, j, k] for j in range(4))
        e0, e1, e2, e3 = (e[:, j, k] for j in range(4))
        if is_vf:
            p000, p001, p011, p010 = s0 - s1, s0 - s3, s0 - (s2 + s3 - s1), s0 - s2
            p100, p101, p111, p110 = e0 - e1, e0 - e3, e0 - (e2 + e3 - e1), e0 - e2
        else:
            p000, p001, p010, p011 = s0 - s2, s0 - s3, s1 - s2, s1 - s3
            p100, p101, p110, p111 = e0 - e2, e0 - e3, e1 - e2, e1 - e3
        L[0] = np.maximum(L[0], absmax((p000, p100), (p001, p101), (p011, p111), (p010, p110)))
        L[1] = np.maximum(L[1], absmax((p000, p010), (p100, p110), (p101, p111), (p001, p011)))
        L[2] = np.maximum(L[2], absmax((p000, p001), (p100, p101), (p110, p111), (p010, p011)))
    if is_vf:
        A0, A1 = s[:, 0:1], e[:, 0:1]
        B0 = np.concatenate([s[:, 1:4], (s[:, 2] + s[:, 3] - s[:, 1])[:, None]], axis=1)
        B1 = np.concatenate([e[:, 1:4], (e[:, 2] + e[:, 3] - e[:, 1])[:, None]], axis=1)
        width = np.full(len(p), tol)
    else:
        A0, A1 = s[:, 0:2], e[:, 0:2]
        B0, B1 = s[:, 2:4], e[:, 2:4]
        with np.errstate(divide="ignore", invalid="ignore"):
            width = np.where((L[0] > 0) & (L[1] > 0),
                             tol * (1 + L[1] / L[0] + L[2] / L[1]) / 3 * 1.000001, np.inf)
        width = np.maximum(width, tol)
    maxabs = np.maximum(1.0, np.abs(p).reshape(len(p), -1).max(1))
    extent = p.reshape(len(p), -1).max(1) - p.reshape(len(p), -1).min(1)
    if f32:
        bound = 2.0 * (width + ms + 2.0 * maxabs ** 3 * 8e-6 + 1e-6 * maxabs)
        sane = (extent <= tol * 1e12) & (L.max(0) <= tol * 1e6)
    else:
        bound = 2.0 * (width + ms + 2.0 * maxabs ** 3 * 8e-15 + 1e-12 * maxabs)
        sane = extent <= tol * 1e12
    gap = np.maximum(B0.min(1) - A0.max(1), A0.min(1) - B0.max(1))          # per axis, t = 0
    D = np.abs(A1 - A0).max(1) + np.abs(B1 - B0).max(1)                    # closing speed bound
    with np.errstate(divide="ignore", invalid="ignore"):
        tk = np.where(gap > bound[:, None], (gap - bound[:, None]) / D, 0.0)   # D = 0 -> inf
    t = np.where(sane, tk.max(1), 0.0)
    return np.minimum(t * (1.0 - 1e-9), 2.0)


@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("tol,ms", [(1e-6, 0.0), (1e-3, 0.0), (1e-6, 1e-3), (1e-9, 1e-8), (1e-2, 1e-2)])
def test_toi_lower_bound_never_exceeds_the_solver(orc, sccd, scene_c1, tol, ms, f32):
    """Every per-query TOI of the oracle (and every "no collision") respects the bound, on a
    cloth scene, a rigid-body pile and the adversarial queries -- so ordering by it and skipping
    queries whose bound is not below the running earliest toi changes no result."""
    sets = []
    vb, eb, fb = orc.build_boxes(scene_c1, ms, f32)
    pile = sccd.scenes.blob_pile(60, seed=11)
    pvb, peb, pfb = orc.build_boxes(pile, ms, f32)
    for scene, (v, ed, f) in ((scene_c1, (vb, eb, fb)), (pile, (pvb, peb, pfb))):
        vf = orc.canonical(orc.sort_and_sweep_two_lists(v, f, 0, f32)[0])
        ee = orc.canonical(orc.sort_and_sweep(ed, 0, f32)[0])
        sets.append((orc.gather_queries(scene, vf, True), True))
        sets.append((orc.gather_queries(scene, ee, False), False))
    ee5, vf5 = sccd.scenes.queries_c5(1500, seed=4, parallel=False)
    sets += [(vf5, True), (ee5, False)]
    n_pos = 0
    for q, is_vf in sets:
        m = orc.tractable(q, is_vf, ms, tol, limit=20000, f32=f32)
        q = q[m]
        lb = toi_lower_bound(q, is_vf, tol, ms, f32)
        _, tpq, _ = orc.narrow_phase(q, is_vf, ms, -1, tol, True, f32=f32)
        assert np.all(lb <= tpq)                 # (misses are +inf)
        _, tpq0, _ = orc.narrow_phase(q, is_vf, ms, -1, tol, False, f32=f32)   # allow_zero_toi off
        assert np.all(lb <= tpq0)
        n_pos += int(((lb > 0) & (tpq < 1)).sum())
    if ms == 0.0 and tol == 1e-6:
        assert n_pos > 100                       # it is not vacuous: hits with a positive bound


@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("tol,ms", [(1e-6, 0.0), (1e-3, 0.0), (1e-6, 1e-3), (1e-9, 1e-8), (1e-2, 1e-2)])
def test_culled_mesh_queries_are_misses(orc, scene_c1, tol, ms, f32):
    s = scene_c1
    if f32 and ms == 1e-3:
        pytest.skip("float build with ms = 1e-3: >1e8 box checks on the CPU")
    if f32 and tol == 1e-9:
        pytest.skip("tol = 1e-9 is below the float resolution: the float cull keeps every query "
                    "(scale test) and the solver needs 90 s on the CPU")
    r = orc.ccd(s, f32=f32, per_query=False)        # the candidate pairs of that scalar type
    for pairs, is_vf in ((r["vf"], True), (r["ee"], False)):
        q = orc.gather_queries(s, np.ascontiguousarray(pairs), is_vf)
        _, tpq, _ = orc.narrow_phase(q, is_vf, ms, -1, tol, True, 1.0, per_query=True, f32=f32)
        culled = cull_mask(q, is_vf, tol, ms, f32)
        assert not np.any(culled & (tpq < 1)), "the cull dropped a query the root finder reports"
        if tol == 1e-6 and ms == 0.0:
            assert culled.mean() > (0.85 if f32 else 0.9)      # and it is worth having


@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("slab", [False, True])
@pytest.mark.parametrize("tol,ms", [(1e-6, 0.0), (1e-4, 1e-4)])
def test_culled_rigid_body_queries_are_misses(orc, sccd, tol, ms, slab, f32):
    """Configs 3 / 4 in small: rotating, translating rigid blobs -- other L_t / L_u / L_v ratios
    than a falling cloth (the edge-edge acceptance width depends on them)."""
    s = sccd.scenes.blob_pile(60, seed=6 if slab else 5, slab=slab)
    vb, eb, fb = orc.build_boxes(s, ms, f32=f32)
    vf, _ = orc.sort_and_sweep_two_lists(vb, fb, 0, f32=f32)
    ee, _ = orc.sort_and_sweep(eb, 0, f32=f32)
    n_culled = 0
    for pairs, is_vf in ((orc.canonical(vf), True), (orc.canonical(ee), False)):
        q = orc.gather_queries(s, pairs, is_vf)
        keep = orc.tractable(q, is_vf, ms, tol, f32=f32)
        q = q[keep]
        _, tpq, _ = orc.narrow_phase(q, is_vf, ms, -1, tol, True, 1.0, per_query=True, f32=f32)
        culled = cull_mask(q, is_vf, tol, ms, f32)
        assert not np.any(culled & (tpq < 1)), "the cull dropped a query the root finder reports"
        assert (tpq < 1).any()                    # the scene has real contacts
        n_culled += int(culled.sum())
    assert n_culled > 0


@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("tol,ms", [(1e-6, 0.0), (1e-9, 0.0), (1e-6, 1e-8), (1e-3, 0.0)])
def test_culled_adversarial_queries_are_misses(orc, sccd, tol, ms, f32):
    ee, vf = sccd.scenes.queries_c5(3000 if not f32 else 1200, seed=4)
    for q, is_vf in ((vf, True), (ee, False)):
        q = q[orc.tractable(q, is_vf, ms, tol, f32=f32)]
        _, tpq, _ = orc.narrow_phase(q, is_vf, ms, -1, tol, True, 1.0, per_query=True, f32=f32)
        culled = cull_mask(q, is_vf, tol, ms, f32)
        assert not np.any(culled & (tpq < 1))


def test_static_edges_are_never_culled(sccd):
    """L_t = 0 (no motion): the reference's edge-edge tolerances are infinite, so it can accept
    boxes of any size -- the cull must keep such queries."""
    ee, _ = sccd.scenes.queries_c5(200, seed=1)
    q = ee.copy().reshape(-1, 2, 4, 3)
    q[:, 1] = q[:, 0]                                # end positions = start positions
    q[:, :, 2:, 2] += 5.0                            # far apart along z
    assert not cull_mask(q.reshape(-1, 24), False, 1e-6, 0.0).any()
    assert not cull_mask(q.reshape(-1, 24), False, 1e-6, 0.0, f32=True).any()


def test_float_cull_keeps_queries_whose_tolerance_nears_the_float_resolution(sccd):
    """tol_k = tol / (3 L_k) below ~2^-23: the float solver can stop on condition 4 (interval
    cannot be split) with a hull of any size, so such queries must reach it."""
    ee, vf = sccd.scenes.queries_c5(200, seed=2)
    for q, is_vf in ((vf, True), (ee, False)):
        big = q.reshape(-1, 2, 4, 3).copy()
        big[:, :, 2:] += 1e3                         # far apart: L ~ 1e3 >> tol * 1e6
        assert cull_mask(big.reshape(-1, 24), is_vf, 1e-6, 0.0).any()           # double: culled
        assert not cull_mask(big.reshape(-1, 24), is_vf, 1e-6, 0.0, f32=True).any()


@pytest.mark.parametrize("f32", [False, True])
@pytest.mark.parametrize("is_vf", [True, False])
@pytest.mark.parametrize("tol", [1e-6, 1e-4])
def test_fuzz_near_threshold_separations(orc, is_vf, tol, f32):
    """Random primitives pushed apart along random diagonals by gaps around the acceptance
    width (where edge-edge is much looser than the co-domain tolerance because of the
    reference's tol_u = tol_t): every culled query must be a miss of the root finder."""
    rng = np.random.default_rng(11 if is_vf else 12)
    n = 4000
    base = rng.uniform(-1, 1, (n, 1, 4, 3))
    motion = rng.uniform(-1, 1, (n, 1, 4, 3)) * 10.0 ** rng.uniform(-5, 0, (n, 1, 1, 1))
    p = np.concatenate([base, base + motion], axis=1)          # [n, time, vertex, xyz]
    # collapse primitive B onto A's neighbourhood, then push it away along a diagonal
    centre_a = p[:, :, :1 if is_vf else 2].mean(axis=(1, 2), keepdims=True)
    sl = slice(1, 4) if is_vf else slice(2, 4)
    centre_b = p[:, :, sl].mean(axis=(1, 2), keepdims=True)
    p[:, :, sl] += centre_a - centre_b
    axis = DIAGS[rng.integers(0, 6, n)] * rng.choice([-1.0, 1.0], (n, 1))
    gap = 10.0 ** rng.uniform(np.log10(tol) - 1, np.log10(tol) + 4, n)
    # extent of both primitives along the axis, so that `gap` is the real separation
    proj = p.reshape(n, 8, 3) @ axis[:, :, None]
    pa = proj.reshape(n, 2, 4)[:, :, :1 if is_vf else 2].reshape(n, -1)
    pb = proj.reshape(n, 2, 4)[:, :, sl].reshape(n, -1)
    if is_vf:   # the face's parallelogram corner reaches further than its vertices
        f = p[:, :, 1:4]
        fourth = (f[:, :, 1] + f[:, :, 2] - f[:, :, 0]) @ axis[:, :, None]
        pb = np.concatenate([pb, fourth.reshape(n, -1)], axis=1)
    shift = (pa.max(1) - pb.min(1) + gap) / (axis * axis).sum(1)
    p[:, :, sl] += (shift[:, None] * axis)[:, None, None, :]
    q = np.ascontiguousarray(p.reshape(n, 24))
    q = q[orc.tractable(q, is_vf, 0.0, tol, limit=5000, f32=f32)]
    _, tpq, _ = orc.narrow_phase(q, is_vf, 0.0, -1, tol, True, 1.0, per_query=True, f32=f32)
    culled = cull_mask(q, is_vf, tol, 0.0, f32)
    assert culled.any() and (~culled).any()                    # the sample straddles the bound
    assert not np.any(culled & (tpq < 1))


def cull_mask_float_pretest(q, is_vf, tol, ms):
    """numpy (float32) mirror of cull_query_float in csrc/narrow.cu: the pre-test in front of the
    double test.  True = culled by the pre-test alone."""
    f = q.astype(np.float32).reshape(-1, 2, 4, 3)
    s, e = f[:, 0], f[:, 1]
    n = len(f)
    flat = f.reshape(n, -1)
    mx = np.maximum(np.float32(1), np.abs(flat).max(1))
    mxu = mx.astype(np.float64) * 1.000001
    ok = mx <= np.float32(1e30)
    ok &= (flat.max(1) - flat.min(1)).astype(np.float64) * 1.001 + 1e-6 * mxu <= tol * 1e12
    if is_vf:
        A = np.stack([s[:, 0], e[:, 0]], axis=1)
        B = np.stack([s[:, 1], s[:, 2], s[:, 3], e[:, 1], e[:, 2], e[:, 3],
                      s[:, 2] + s[:, 3] - s[:, 1], e[:, 2] + e[:, 3] - e[:, 1]], axis=1)
        width_up = np.full(n, tol)
    else:
        A = np.stack([s[:, 0], s[:, 1], e[:, 0], e[:, 1]], axis=1)
        B = np.stack([s[:, 2], s[:, 3], e[:, 2], e[:, 3]], axis=1)
        L = np.zeros((3, n), np.float32)
        for k in range(3):
            s0, s1, s2, s3 = (s[:, j, k] for j in range(4))
            e0, e1, e2, e3 = (e[:, j, k] for j in range(4))
            p000, p001, p010, p011 = s0 - s2, s0 - s3, s1 - s2, s1 - s3
            p100, p101, p110, p111 = e0 - e2, e0 - e3, e1 - e2, e1 - e3
            L[0] = np.maximum(L[0], absmax((p000, p100), (p001, p101), (p011, p111), (p010, p110)))
            L[1] = np.maximum(L[1], absmax((p000, p010), (p100, p110), (p101, p111), (p001, p011)))
            L[2] = np.maximum(L[2], absmax((p000, p001), (p100, p101), (p110, p111), (p010, p011)))
        dl = 1e-6 * mxu
        l0, l1 = L[0].astype(np.float64) - dl, L[1].astype(np.float64) - dl
        ok &= (l0 > 0) & (l1 > 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            width_up = tol * (1 + (L[1] + dl) / l0 + (L[2] + dl) / l1) / 3 * 1.00001
        width_up = np.maximum(np.where(ok, width_up, np.inf), tol)
    # float32 projections, one addition each (numpy keeps float32 for float32 operands)
    seps = []
    for i0, i1, sg in ((0, 1, 1), (0, 1, -1), (0, 2, 1), (0, 2, -1), (1, 2, 1), (1, 2, -1)):
        pa = A[:, :, i0] + np.float32(sg) * A[:, :, i1]
        pb = B[:, :, i0] + np.float32(sg) * B[:, :, i1]
        seps.append(np.maximum(pa.min(1) - pb.max(1), pb.min(1) - pa.max(1)))
    sep = np.max(np.stack(seps), axis=0).astype(np.float64)
    bound_up = 2.0 * (width_up + ms + 2.0 * mxu ** 3 * 8e-15 + 1e-12 * mxu) * 1.000001
    return ok & (0.5 * (sep - 1e-5 * mxu) > bound_up)


@pytest.mark.parametrize("tol,ms", [(1e-6, 0.0), (1e-9, 0.0), (1e-6, 1e-8), (1e-3, 0.0), (1e-6, 1e-3)])
def test_float_pretest_culls_a_subset_of_the_double_test(sccd, scene_c1, orc, tol, ms):
    """The float pre-test (double build) may only cull what the double test culls -- its margin
    must cover every float rounding: mesh queries (cloth, piles), adversarial queries, queries
    with separations around the threshold and around the margin, huge coordinates."""
    rng = np.random.default_rng(5)
    sets = []
    for scene in (scene_c1, sccd.scenes.blob_pile(40, seed=3)):
        got = orc.ccd(scene)
        for pairs, is_vf in ((got["vf"], True), (got["ee"], False)):
            sets.append((orc.gather_queries(scene, np.ascontiguousarray(pairs[:60000]), is_vf), is_vf))
    ee, vf = sccd.scenes.queries_c5(4000, seed=6)
    sets += [(vf, True), (ee, False)]
    # random primitives pushed apart by gaps from far below the threshold to far above the margin,
    # at coordinate magnitudes from 1 to 1e6
    for is_vf in (True, False):
        n = 20000
        base = rng.uniform(-1, 1, (n, 1, 4, 3))
        motion = rng.uniform(-1, 1, (n, 1, 4, 3)) * 10.0 ** rng.uniform(-6, 0, (n, 1, 1, 1))
        p = np.concatenate([base, base + motion], axis=1)
        sl = slice(1, 4) if is_vf else slice(2, 4)
        axis = DIAGS[rng.integers(0, 6, n)] * rng.choice([-1.0, 1.0], (n, 1))
        gap = 10.0 ** rng.uniform(-8, 1, n)
        p[:, :, sl] += (gap[:, None] * axis)[:, None, None, :]
        p *= 10.0 ** rng.integers(0, 7, (n, 1, 1, 1))
        p += rng.uniform(-1, 1, (n, 1, 1, 3)) * 10.0 ** rng.integers(0, 5, (n, 1, 1, 1))
        sets.append((np.ascontiguousarray(p.reshape(n, 24)), is_vf))
    n_pre = n_dbl = 0
    for q, is_vf in sets:
        pre = cull_mask_float_pretest(q, is_vf, tol, ms)
        dbl = cull_mask(q, is_vf, tol, ms)
        assert not np.any(pre & ~dbl), "the float pre-test culled a query the double test keeps"
        n_pre += int(pre.sum())
        n_dbl += int(dbl.sum())
    assert n_dbl > 0 and n_pre > 0.5 * n_dbl      # ... and it does decide most of them
