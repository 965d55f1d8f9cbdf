"""Parity at the sizes BASELINE.json names (`-m gpu`, through the C ABI).

The reference's acceptance tests are set equality of the broad phase and the earliest TOI of the
narrow phase on real scenes (tests/test_broad_phase.cu:94-121, tests/test_narrow_phase.cu:41-65).
Here, on the synthetic configs:
  * config 2 (1.06 M primitives): the full vertex-face / edge-edge overlap SETS against the
    unmodified reference CPU build (oracle/_ref/libref_sccd_cpu.so, the oracle port when it is
    absent) and EVERY per-query TOI (the collision list of the TOI_PER_QUERY build) against the
    oracle, separating-axis cull on and off;
  * configs 3 / 4 shape: 1,000-instance blob piles (602 K boxes; the reference's one-axis CPU
    sweep is O(N^5/3) on a dense pile -- 38 s for 2,000 instances on 8 cores -- so the CPU check
    stops there) incl. the concatenation of 8 shards; config 3 at its FULL size (6.0 M boxes,
    18 M queries) through size-independent properties: the set does not depend on the cell
    grid, the sweep axis or the sharding, the TOI not on the cull;
  * config 5 at 10^6 + 10^6 adversarial queries with max_iter = 10^4, a bounded item list
    (2^22) and the four (ms, tol) corners: queries under the cap are bit-equal to the oracle
    (10 % sample), capped ones never later, overflow of the item list is reported, and the
    whole 10^6 is identical with the cull on and off.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def keys(p):
    """(n, 2) int32 pairs -> sorted int64 keys (set comparison without np.unique's copy)."""
    p = np.asarray(p, dtype=np.int32).reshape(-1, 2)
    k = (p[:, 0].astype(np.int64) << 32) | (p[:, 1].astype(np.int64) & 0xFFFFFFFF)
    k.sort()
    return k


def ref_sets(orc, scene, r=0.0):
    """Overlap sets of the unmodified reference CPU build when oracle/_ref holds it (it travels
    to the GPU box as a built library), else of the oracle's restatement."""
    if orc.ref_cpu() is not None:
        b = orc.ref_cpu_broad_phase(scene, r=r)
        return orc.canonical(b["vf"]), orc.canonical(b["ee"])
    vb, eb, fb = orc.build_boxes(scene, r)
    return (orc.canonical(orc.sort_and_sweep_two_lists(vb, fb, 0)[0]),
            orc.canonical(orc.sort_and_sweep(eb, 0)[0]))


def oracle_hits(orc, scene, pairs, is_vf, chunk=2_000_000, **kw):
    """(ids, toi) of every query with toi < 1, per-query mode (narrow_phase.cu:69-82), in the
    order of `pairs`; gathered and solved in chunks to bound host memory."""
    ids, tois = [], []
    for lo in range(0, len(pairs), chunk):
        p = np.ascontiguousarray(pairs[lo:lo + chunk])
        q = orc.gather_queries(scene, p, is_vf)
        _, tpq, _ = orc.narrow_phase(q, is_vf, per_query=True, **kw)
        hit = tpq < 1
        ids.append(p[hit])
        tois.append(tpq[hit])
    return np.concatenate(ids), np.concatenate(tois)


def sort_hits(ids, toi):
    order = np.lexsort((ids[:, 1], ids[:, 0]))
    return ids[order], toi[order]


def check_scene_against_cpu(ctx, sccd, orc, scene, culls=(1, 0)):
    want_vf, want_ee = ref_sets(orc, scene)
    ctx.upload_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"])
    ctx.build_boxes(0.0)
    got_vf, got_ee = ctx.broad_phase(0), ctx.broad_phase(1)
    # no duplicates, and the same set
    assert len(got_vf) == len(want_vf) and len(got_ee) == len(want_ee)
    assert np.array_equal(keys(got_vf), keys(want_vf))
    assert np.array_equal(keys(got_ee), keys(want_ee))
    # every per-query TOI: the collision list of the TOI_PER_QUERY build
    o_vf = sort_hits(*oracle_hits(orc, scene, want_vf, True))
    o_ee = sort_hits(*oracle_hits(orc, scene, want_ee, False))
    o_toi = min([1.0] + list(o_vf[1]) + list(o_ee[1]))
    for cull in culls:
        ctx.set_option(sccd.capi.OPT_NARROW_CULL, cull)
        try:
            toi, (vi, vt), (ei, et) = ctx.ccd_collisions()
            plain = ctx.ccd()
        finally:
            ctx.set_option(sccd.capi.OPT_NARROW_CULL, 1)
        for (gi, gt), (oi, ot) in (((vi, vt), o_vf), ((ei, et), o_ee)):
            gi, gt = sort_hits(gi, gt)
            assert np.array_equal(gi, oi)            # hit / miss of every query
            assert np.array_equal(gt, ot)            # tolerance 0
        assert toi == o_toi == plain
    return len(want_vf), len(want_ee), o_toi


def test_config2_sets_and_every_toi_match_the_cpu_reference(ctx, sccd, orc):
    s = sccd.scenes.scene_c2()
    n_vf, n_ee, toi = check_scene_against_cpu(ctx, sccd, orc, s)
    assert n_vf > 300_000 and n_ee > 1_000_000 and toi < 1.0
    if orc.ref_cuda(False) is not None:      # the reference's own CUDA ccd() on the same box
        assert orc.ref_cuda_ccd(s, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True)["toi"] == toi


@pytest.mark.parametrize("slab", [False, True])
def test_blob_pile_sets_tois_and_eight_shards(sccd, orc, slab):
    """Configs 3 / 4 shape at 1,000 instances (602 K boxes) against the CPU reference; the pair
    lists of 8 shards concatenated in rank order are the single-device list."""
    s = sccd.scenes.blob_pile(1000, seed=3 if slab else 2, slab=slab)
    c = sccd.Context(0)
    try:
        check_scene_against_cpu(c, sccd, orc, s, culls=(1,))
        c.build_boxes(0.0)
        whole = [c.broad_phase(0), c.broad_phase(1)]
        for kind in (0, 1):
            parts = []
            for r in range(8):
                c.set_shard(r, 8)
                c.build_boxes(0.0)
                parts.append(c.broad_phase(kind))
            c.set_shard(0, 1)
            assert sum(len(p) > 0 for p in parts) >= 6            # the work really is spread
            assert np.array_equal(np.concatenate(parts), whole[kind])
    finally:
        c.close()


def test_config3_full_size_properties(sccd):
    """Config 3 at its full size (10,000 blobs, 6.0 M boxes, ~18 M queries): the overlap set
    does not depend on the cell grid, the sweep axis or 8-way sharding; the TOI does not depend
    on the separating-axis cull or the sweep axis."""
    K = sccd.capi
    s = sccd.scenes.scene_c3()
    c = sccd.Context(0)
    try:
        c.upload_mesh(s["V0"], s["V1"], s["E"], s["F"])
        c.build_boxes(0.0)
        base = [c.broad_phase(0), c.broad_phase(1)]
        base_keys = [keys(base[0]), keys(base[1])]
        assert len(base[0]) + len(base[1]) > 15_000_000
        for k in (0, 1):
            assert np.all(base_keys[k][1:] != base_keys[k][:-1])        # duplicate-free
        toi = c.ccd()
        assert 0.0 < toi < 1.0
        c.set_option(K.OPT_NARROW_CULL, 0)
        assert c.ccd() == toi
        c.set_option(K.OPT_NARROW_CULL, 1)
        # another cell grid
        c.set_grid_cells(4096)
        c.build_boxes(0.0)
        for k in (0, 1):
            assert np.array_equal(keys(c.broad_phase(k)), base_keys[k])
        c.set_grid_cells(0)
        # another sweep axis (SCCD_OPT_SWEEP_AXIS), incl. the automatic choice
        for axis in (1, -1):
            c.set_option(K.OPT_SWEEP_AXIS, axis)
            assert c.ccd() == toi
            st = c.stats()
            assert st["n_pairs"] == [len(base[0]), len(base[1])]
            assert np.array_equal(keys(c.broad_phase(1)), base_keys[1])
        c.set_option(K.OPT_SWEEP_AXIS, 0)
        # 8 shards: concatenation in rank order is the single-device list
        parts = []
        for r in range(8):
            c.set_shard(r, 8)
            c.build_boxes(0.0)
            parts.append(c.broad_phase(1))
        c.set_shard(0, 1)
        assert np.array_equal(np.concatenate(parts), base[1])
    finally:
        c.close()


C5_CASES = [dict(ms=0.0, tol=1e-6), dict(ms=0.0, tol=1e-9), dict(ms=1e-8, tol=1e-6),
            dict(ms=1e-8, tol=1e-9)]


@pytest.mark.parametrize("kw", C5_CASES, ids=lambda k: f"ms{k['ms']:g}-tol{k['tol']:g}")
def test_config5_million_queries_capped_and_bounded(sccd, orc, torch_cuda, kw):
    """10^6 vertex-face + 10^6 edge-edge adversarial queries, max_iter = 10^4, item lists bounded
    to 2^22 entries.  sccd_narrow_phase_checks gives every query's box counter (the reference's
    nbr_checks): a counter <= max_iter + 1 means the cap never touched the query."""
    torch = torch_cuda
    N, MAX_ITER, SAMPLE = 1_000_000, 10_000, 100_000
    K = sccd.capi
    ee, vf = sccd.scenes.queries_c5(N, seed=4)
    c = sccd.Context(0)
    c.set_queue_capacity(1 << 22)
    rng = np.random.default_rng(11)
    try:
        for kind, q in ((0, vf), (1, ee)):
            dq = torch.from_numpy(q).cuda()
            res = {}
            for cull in (1, 0):
                c.set_option(K.OPT_NARROW_CULL, cull)
                tq = torch.empty(N, dtype=torch.float64, device="cuda")
                c.reset_stats()
                toi = c.narrow_phase_queries(kind, dq.data_ptr(), n=N, max_iter=MAX_ITER,
                                             allow_zero_toi=True, d_toi_per_query=tq.data_ptr(),
                                             **kw)
                ptr, n = c.narrow_phase_checks()
                assert n == N
                checks = torch.as_tensor(sccd.multigpu._DevArray(ptr, (N,), "<u4"),
                                         device="cuda").cpu().numpy()
                res[cull] = (toi, tq.cpu().numpy(), checks, c.stats())
            c.set_option(K.OPT_NARROW_CULL, 1)
            toi, tpq, checks, st = res[1]
            toi0, tpq0, checks0, _ = res[0]
            under = checks <= MAX_ITER + 1
            under0 = checks0 <= MAX_ITER + 1
            # a culled query is a 0-check "no collision"; the cull never changes an uncapped answer
            both = under & under0
            assert both.mean() > 0.5
            assert np.array_equal(tpq[both], tpq0[both])
            assert st["n_queries"][kind] == N
            assert st["n_capped"][kind] <= int((~under).sum())   # (a pruned box may pass the cap)
            assert toi == min(1.0, tpq.min()) and toi0 == min(1.0, tpq0.min())
            # the oracle on a 10 % sample (cap_mode 1 = the same conservative rule)
            idx = np.sort(rng.choice(N, SAMPLE, replace=False))
            _, otpq, ost = orc.narrow_phase(q[idx], kind == 0, kw["ms"], MAX_ITER, kw["tol"], True,
                                            cap_mode=1)
            o_under = ost["checks"] <= MAX_ITER + 1
            g, g_under = tpq[idx], under[idx]
            exact = g_under & o_under
            assert exact.sum() > 0.5 * SAMPLE
            assert np.array_equal(g[exact], otpq[exact])                 # tolerance 0
            assert np.array_equal(g[exact] < 1, otpq[exact] < 1)         # hit / miss
            # capped on one side only: the capped value is never later than the exact one
            assert np.all(g[~g_under & o_under] <= otpq[~g_under & o_under])
            assert np.all(otpq[g_under & ~o_under] <= g[g_under & ~o_under])
            if kw["ms"] > 0:
                assert (~under).sum() > 0        # these corners do reach the cap
    finally:
        c.close()


def test_item_list_overflow_is_reported_and_loses_nothing(sccd, orc, torch_cuda):
    """A tiny bounded item list under many concurrent mixed-size hand-ons (the reservation is a
    compare-and-swap that never over-commits): overflow is reported, every query keeps its exact
    answer."""
    torch = torch_cuda
    ee, vf = sccd.scenes.queries_c5(60_000, seed=31)
    c = sccd.Context(0)
    try:
        for cap in (64, 4096):
            c.set_queue_capacity(cap)
            for kind, q in ((0, vf), (1, ee)):
                m = orc.tractable(q, kind == 0, 0.0, 1e-6, limit=4000)
                qq = np.ascontiguousarray(q[m])
                tq = torch.empty(len(qq), dtype=torch.float64, device="cuda")
                c.reset_stats()
                toi = c.narrow_phase_queries(kind, qq, d_toi_per_query=tq.data_ptr())
                otoi, otpq, _ = orc.narrow_phase(qq, kind == 0)
                assert np.array_equal(tq.cpu().numpy(), otpq) and toi == otoi
                if cap == 64:
                    assert c.stats()["queue_overflow"] == 1
    finally:
        c.close()
