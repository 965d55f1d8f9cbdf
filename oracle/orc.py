"""ctypes bindings for the CPU oracle (oracle/liborc.so) and, when present, the
unmodified reference built into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

AABB_DTYPE = np.dtype(
    [("min", np.float64, 3), ("max", np.float64, 3), ("vids", np.int32, 3), ("elem", np.int32)])
assert AABB_DTYPE.itemsize == 64

_p = lambda a: a.ctypes.data_as(C.c_void_p)


def build(ref: bool = False) -> None:
    """Compile liborc.so (and oracle/_ref when the reference sources are present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref_cpu", "ref_cuda", "ref_f32"])


_libs = {}


def lib(f32: bool = False):
    """liborc.so (the reference's default double build) or liborc_f32.so (its float build,
    SCALABLE_CCD_USE_DOUBLE off).  Same C interface: double arrays in and out."""
    f32 = bool(f32)
    if f32 not in _libs:
        path = os.path.join(HERE, "liborc_f32.so" if f32 else "liborc.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_sort_and_sweep.restype = C.c_int64
        L.orc_sort_and_sweep_two_lists.restype = C.c_int64
        L.orc_brute_force.restype = C.c_int64
        L.orc_narrow_phase_bfs.restype = C.c_int64
        assert L.orc_sizeof_aabb() == 64 and L.orc_sizeof_query() == 192
        assert bool(L.orc_is_f32()) == f32
        _libs[f32] = L
    return _libs[f32]


class NpStats(C.Structure):
    _fields_ = [("box_checks", C.c_int64), ("max_stack", C.c_int64), ("capped_queries", C.c_int64)]


def _mesh_args(s):
    V0, V1, E, F = s["V0"], s["V1"], s["E"], s["F"]
    for a in (V0, V1, E, F):
        assert a.flags.f_contiguous
    return V0, V1, E, F


def build_boxes(scene, r: float = 0.0, f32: bool = False):
    """(vertex, edge, face) boxes as AABB_DTYPE arrays -- aabb.cu:115-229."""
    V0, V1, E, F = _mesh_args(scene)
    nV, nE, nF = V0.shape[0], E.shape[0], F.shape[0]
    vb = np.zeros(nV, AABB_DTYPE)
    eb = np.zeros(nE, AABB_DTYPE)
    fb = np.zeros(nF, AABB_DTYPE)
    L = lib(f32)
    L.orc_build_vertex_boxes(_p(V0), _p(V1), C.c_int64(nV), C.c_double(r), _p(vb))
    L.orc_build_edge_boxes(_p(vb), _p(E), C.c_int64(nE), _p(eb))
    L.orc_build_face_boxes(_p(vb), _p(F), C.c_int64(nF), _p(fb))
    return vb, eb, fb


def _grow(fn, guess):
    cap = max(int(guess), 1024)
    while True:
        out = np.empty((cap, 2), np.int32)
        n = fn(out, cap)
        if n <= cap:
            return out[:n].copy()
        cap = int(n)


def sort_and_sweep(boxes, axis: int = 0, f32: bool = False):
    """Single-list SAP -- sort_and_sweep.cpp:198-211.  Returns (pairs, next_axis).
    (f32 only matters for the next axis: the variance is accumulated in Scalar.)"""
    ax = C.c_int(axis)
    L = lib(f32)

    def fn(out, cap):
        ax.value = axis
        return L.orc_sort_and_sweep(
            _p(boxes), C.c_int64(len(boxes)), C.byref(ax), _p(out), C.c_int64(cap))

    pairs = _grow(fn, 32 * len(boxes))
    return pairs, ax.value


def sort_and_sweep_two_lists(A, B, axis: int = 0, f32: bool = False):
    """Two-list SAP (A = vertices, B = faces) -- sort_and_sweep.cpp:213-239."""
    ax = C.c_int(axis)
    L = lib(f32)

    def fn(out, cap):
        ax.value = axis
        return L.orc_sort_and_sweep_two_lists(
            _p(A), C.c_int64(len(A)), _p(B), C.c_int64(len(B)), C.byref(ax), _p(out),
            C.c_int64(cap))

    pairs = _grow(fn, 16 * (len(A) + len(B)))
    return pairs, ax.value


def brute_force(A, B=None):
    L = lib()
    return _grow(lambda out, cap: L.orc_brute_force(
        _p(A), C.c_int64(len(A)), _p(B) if B is not None else None,
        C.c_int64(len(B) if B is not None else 0), _p(out), C.c_int64(cap)), 64 * len(A))


def canonical(pairs) -> np.ndarray:
    """Sorted-unique view of an overlap list: the set the reference is judged on
    (tests/ground_truth.cpp:55-63 compares as sets)."""
    p = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
    if len(p) == 0:
        return p
    key = (p[:, 0].astype(np.int64) << 32) | p[:, 1].astype(np.int64)
    key = np.unique(key)
    return np.stack([(key >> 32).astype(np.int32), (key & 0xFFFFFFFF).astype(np.int32)], 1)


def gather_queries(scene, pairs, is_vf: bool) -> np.ndarray:
    """narrow_phase.cu:24-74 add_data -> (n, 24) float64."""
    V0, V1, E, F = _mesh_args(scene)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32)
    out = np.empty((len(pairs), 24), np.float64)
    lib().orc_gather_queries(
        _p(V0), _p(V1), C.c_int64(V0.shape[0]), _p(E), C.c_int64(E.shape[0]), _p(F),
        C.c_int64(F.shape[0]), _p(pairs), C.c_int64(len(pairs)), C.c_int(int(is_vf)), _p(out))
    return out


def narrow_phase(queries, is_vf: bool, ms: float = 0.0, max_iter: int = -1, tol: float = 1e-6,
                 allow_zero_toi: bool = True, toi: float = 1.0, per_query: bool = True,
                 cap_mode: int = 1, f32: bool = False):
    """Tight-Inclusion over (n, 24) query arrays.  Returns (toi, toi_per_query | None,
    stats dict); stats["checks"] holds the per-query box counts."""
    q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 24)
    t = C.c_double(toi)
    tpq = np.empty(len(q), np.float64) if per_query else None
    st = NpStats()
    checks = np.zeros(len(q), np.int64)
    lib(f32).orc_narrow_phase(
        _p(q), C.c_int64(len(q)), C.c_int(int(is_vf)), C.c_double(ms), C.c_int(max_iter),
        C.c_double(tol), C.c_int(int(allow_zero_toi)), C.c_int(cap_mode), C.byref(t),
        _p(tpq) if per_query else None, C.byref(st), _p(checks))
    return t.value, tpq, {"box_checks": st.box_checks, "max_stack": st.max_stack,
                          "capped_queries": st.capped_queries, "checks": checks}


def tractable(queries, is_vf: bool, ms: float, tol: float, allow_zero_toi: bool = True,
              limit: int = 20000, f32: bool = False) -> np.ndarray:
    """Mask of the queries the solver finishes within `limit` box checks.  Grazing
    queries with ms > 0 can need >1e8 boxes (for the reference as well); uncapped parity
    runs use this mask, capped runs use everything."""
    _, _, st = narrow_phase(queries, is_vf, ms, limit, tol, allow_zero_toi, cap_mode=1, f32=f32)
    return st["checks"] <= limit


def narrow_phase_bfs(queries, is_vf: bool, ms: float = 0.0, tol: float = 1e-6,
                     allow_zero_toi: bool = True, cap_items: int = 0, f32: bool = False):
    """The reference's level-synchronous traversal (root_finder.cu:431-447).  Returns
    (toi_per_query, levels, total_checks, widest_level); levels == -1 if the front outgrew
    4 * cap_items and the run was abandoned."""
    q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 24)
    L = lib(f32)
    tpq = np.empty(len(q), np.float64)
    tc, ml = C.c_int64(0), C.c_int64(0)
    lv = L.orc_narrow_phase_bfs(
        _p(q), C.c_int64(len(q)), C.c_int(int(is_vf)), C.c_double(ms), C.c_double(tol),
        C.c_int(int(allow_zero_toi)), _p(tpq), C.byref(tc), C.byref(ml), C.c_int64(cap_items))
    return tpq, lv, tc.value, ml.value


def tractable_bfs(queries, is_vf: bool, ms: float, tol: float, allow_zero_toi: bool = True,
                  front_limit: int = 4096, f32: bool = False) -> np.ndarray:
    """Mask of the queries whose BREADTH-first front stays below front_limit entries.  The
    reference's ring queue silently wraps over unprocessed boxes when a front outgrows it
    (the full-check is not atomic with the push, ccd_buffer.cuh:25-34), so goldens frozen
    from it are only meaningful on such queries."""
    q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 24)
    mask = np.zeros(len(q), bool)
    for i in range(len(q)):
        _, lv, _, ml = narrow_phase_bfs(q[i:i + 1], is_vf, ms, tol, allow_zero_toi,
                                        cap_items=front_limit, f32=f32)
        mask[i] = lv > 0 and ml <= front_limit
    return mask


def ccd(scene, ms: float = 0.0, max_iter: int = -1, tol: float = 1e-6,
        allow_zero_toi: bool = True, per_query: bool = True, f32: bool = False):
    """Whole pipeline on the CPU, as cuda/ccd.cu:80-146 composes it."""
    vb, eb, fb = build_boxes(scene, ms, f32)
    vf, _ = sort_and_sweep_two_lists(vb, fb, 0, f32)
    ee, _ = sort_and_sweep(eb, 0, f32)
    vf, ee = canonical(vf), canonical(ee)
    toi = 1.0
    toi, tvf, s1 = narrow_phase(gather_queries(scene, vf, True), True, ms, max_iter, tol,
                                allow_zero_toi, toi, per_query, f32=f32)
    toi, tee, s2 = narrow_phase(gather_queries(scene, ee, False), False, ms, max_iter, tol,
                                allow_zero_toi, toi, per_query, f32=f32)
    return {"toi": toi, "vf": vf, "ee": ee, "toi_vf": tvf, "toi_ee": tee,
            "box_checks": s1["box_checks"] + s2["box_checks"]}


# ----------------------------------------------------------------------------- _ref
def ref_cpu(f32: bool = False):
    """The unmodified reference CPU broad phase (f32: built with SCALABLE_CCD_USE_DOUBLE off)."""
    path = os.path.join(REF_DIR, "libref_sccd_cpu_f32.so" if f32 else "libref_sccd_cpu.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_cpu_broad_phase.restype = C.c_double
    return L


def ref_cpu_build_boxes(scene, r: float = 0.0, f32: bool = False):
    L = ref_cpu(f32)
    V0, V1, E, F = _mesh_args(scene)
    nV, nE, nF = V0.shape[0], E.shape[0], F.shape[0]
    vb = np.zeros(nV, AABB_DTYPE)
    eb = np.zeros(nE, AABB_DTYPE)
    fb = np.zeros(nF, AABB_DTYPE)
    L.ref_cpu_build_boxes(_p(V0), _p(V1), C.c_int64(nV), _p(E), C.c_int64(nE), _p(F),
                          C.c_int64(nF), C.c_double(r), _p(vb), _p(eb), _p(fb))
    return vb, eb, fb


def ref_cpu_broad_phase(scene, r: float = 0.0, axis: int = 0, want_pairs: bool = True,
                        f32: bool = False):
    """Unmodified reference CPU path: boxes + sort_and_sweep VF + EE.
    Returns dict(vf, ee, axes, seconds, threads)."""
    L = ref_cpu(f32)
    V0, V1, E, F = _mesh_args(scene)
    nV, nE, nF = V0.shape[0], E.shape[0], F.shape[0]
    counts = (C.c_int64 * 2)()
    axes = (C.c_int * 2)()
    cap_vf = cap_ee = 0
    vf = np.empty((1, 2), np.int32)
    ee = np.empty((1, 2), np.int32)
    sec = 0.0
    for _ in range(2):
        sec = L.ref_cpu_broad_phase(
            _p(V0), _p(V1), C.c_int64(nV), _p(E), C.c_int64(nE), _p(F), C.c_int64(nF),
            C.c_double(r), C.c_int(axis), _p(vf), C.c_int64(cap_vf), _p(ee), C.c_int64(cap_ee),
            counts, axes)
        if not want_pairs or (counts[0] <= cap_vf and counts[1] <= cap_ee):
            break
        cap_vf, cap_ee = counts[0], counts[1]
        vf = np.empty((max(cap_vf, 1), 2), np.int32)
        ee = np.empty((max(cap_ee, 1), 2), np.int32)
    return {"vf": vf[:counts[0]] if want_pairs else None,
            "ee": ee[:counts[1]] if want_pairs else None,
            "n_vf": counts[0], "n_ee": counts[1], "axes": (axes[0], axes[1]),
            "seconds": sec, "threads": L.ref_cpu_num_threads()}


def ref_cuda(per_query: bool = False, f32: bool = False):
    name = "libref_sccd_cuda" + ("_pq" if per_query else "") + ("_f32" if f32 else "") + ".so"
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        return None
    if path in _ref_cuda_libs:
        return _ref_cuda_libs[path]
    # RTLD_DEEPBIND: each variant binds its own copies of the inline spdlog / thrust symbols
    L = C.CDLL(path, mode=os.RTLD_NOW | os.RTLD_LOCAL | os.RTLD_DEEPBIND)
    L.ref_cuda_ccd.restype = C.c_double
    L.ref_cuda_ipc_ccd_strategy.restype = C.c_double
    L.ref_cuda_quiet()
    _ref_cuda_libs[path] = L
    return L


_ref_cuda_libs = {}


def ref_cuda_ccd(scene, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True, per_query=False,
                 coll_cap=0, f32=False):
    L = ref_cuda(per_query, f32)
    V0, V1, E, F = _mesh_args(scene)
    ids = np.empty((max(coll_cap, 1), 2), np.int32)
    tois = np.empty(max(coll_cap, 1), np.float64)
    n_coll = C.c_int64(0)
    ms_el = C.c_double(0)
    toi = L.ref_cuda_ccd(
        _p(V0), _p(V1), C.c_int64(V0.shape[0]), _p(E), C.c_int64(E.shape[0]), _p(F),
        C.c_int64(F.shape[0]), C.c_double(ms), C.c_int(max_iter), C.c_double(tol),
        C.c_int(int(allow_zero_toi)), _p(ids), _p(tois), C.c_int64(coll_cap), C.byref(n_coll),
        C.byref(ms_el))
    n = max(min(n_coll.value, coll_cap), 0)
    return {"toi": toi, "ms": ms_el.value, "n_coll": n_coll.value, "coll_ids": ids[:n],
            "coll_toi": tois[:n]}


def ref_cuda_broad_phase(scene, r=0.0, want_pairs=True, f32=False):
    L = ref_cuda(False, f32)
    V0, V1, E, F = _mesh_args(scene)
    counts = (C.c_int64 * 2)()
    ms_el = C.c_double(0)
    cap_vf = cap_ee = 0
    vf = np.empty((1, 2), np.int32)
    ee = np.empty((1, 2), np.int32)
    for _ in range(2):
        L.ref_cuda_broad_phase(
            _p(V0), _p(V1), C.c_int64(V0.shape[0]), _p(E), C.c_int64(E.shape[0]), _p(F),
            C.c_int64(F.shape[0]), C.c_double(r), _p(vf), C.c_int64(cap_vf), _p(ee),
            C.c_int64(cap_ee), counts, C.byref(ms_el))
        if not want_pairs or (counts[0] <= cap_vf and counts[1] <= cap_ee):
            break
        cap_vf, cap_ee = counts[0], counts[1]
        vf = np.empty((max(cap_vf, 1), 2), np.int32)
        ee = np.empty((max(cap_ee, 1), 2), np.int32)
    return {"vf": vf[:counts[0]] if want_pairs else None,
            "ee": ee[:counts[1]] if want_pairs else None,
            "n_vf": counts[0], "n_ee": counts[1], "ms": ms_el.value}


def ref_cuda_narrow_queries(queries, is_vf, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True,
                            toi=1.0, per_query=True, min_queue_units=1 << 26, f32=False):
    L = ref_cuda(per_query, f32)
    q = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 24)
    t = C.c_double(toi)
    tpq = np.empty(len(q), np.float64) if per_query else None
    ms_el = C.c_double(0)
    reruns = L.ref_cuda_narrow_queries(
        _p(q), C.c_int64(len(q)), C.c_int(int(is_vf)), C.c_double(ms), C.c_int(max_iter),
        C.c_double(tol), C.c_int(int(allow_zero_toi)), C.byref(t),
        _p(tpq) if per_query else None, C.byref(ms_el), C.c_int64(min_queue_units))
    return {"toi": t.value, "toi_per_query": tpq, "ms": ms_el.value, "reruns": reruns}
