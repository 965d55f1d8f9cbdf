"""ctypes binding of the C ABI in include/sccd.h (libsccd_b200.so).

Thin by design: every method is one C call on HOST numpy buffers or raw DEVICE pointers
(ints, e.g. torch.Tensor.data_ptr()).  There is no CPU fallback -- if the CUDA library is
missing or no sm_100 device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsccd_b200.so")

VF, EE, BOXES = 0, 1, 2
F64, F32 = 0, 1  # sccd_set_scalar_type: the reference's SCALABLE_CCD_USE_DOUBLE switch
OK, ERR_CUDA, ERR_ARG, ERR_STATE, ERR_MEMORY = 0, -1, -2, -3, -4
(OPT_NARROW_CULL, OPT_NARROW_FLAGS, OPT_NARROW_FLAGS_EE, OPT_NARROW_MAX_DEPTH, OPT_MAX_ITER_MODE,
 OPT_KEY_STEPS, OPT_GRID_SCALE_MILLI, OPT_GRID_REPL_MILLI, OPT_SWEEP_AXIS, OPT_PROFILE,
 OPT_NARROW_SOLVER, OPT_CONCURRENT_PASSES, OPT_SWEEP_STAGED, OPT_REUSE_GRID) = range(1, 15)
UNIQUE_ID_BYTES = 128

AABB_DTYPE = np.dtype(
    [("min", np.float64, 3), ("max", np.float64, 3), ("vids", np.int32, 3), ("elem", np.int32)])

# every symbol include/sccd.h declares (tests check the library exports all of them)
SYMBOLS = [
    "sccd_create", "sccd_destroy", "sccd_last_error", "sccd_set_memory_limit",
    "sccd_set_max_pairs_per_chunk", "sccd_set_queue_capacity", "sccd_set_grid_cells",
    "sccd_set_shard", "sccd_set_scalar_type",
    "sccd_build_vertex_boxes", "sccd_build_element_boxes",
    "sccd_upload_mesh", "sccd_update_vertices", "sccd_build_boxes", "sccd_get_boxes", "sccd_set_boxes",
    "sccd_broad_phase_begin",
    "sccd_broad_phase_partial", "sccd_broad_phase_is_complete", "sccd_broad_phase",
    "sccd_narrow_phase", "sccd_narrow_phase_queries", "sccd_ccd", "sccd_ccd_collisions",
    "sccd_get_collisions", "sccd_set_option", "sccd_get_option", "sccd_narrow_phase_checks",
    "sccd_stats_size", "sccd_comm_get_unique_id", "sccd_comm_create", "sccd_comm_destroy",
    "sccd_ccd_sharded", "sccd_ccd_sharded_host", "sccd_exchange_plan",
    "sccd_ccd_host", "sccd_ipc_ccd_strategy", "sccd_get_stats", "sccd_reset_stats",
    "sccd_synchronize", "sccd_measure_fp64_peak",
    "sccd_version",
]


class Stats(C.Structure):
    _fields_ = [
        ("n_boxes", C.c_int64 * 2), ("n_pairs", C.c_int64 * 2), ("n_candidates", C.c_int64 * 2),
        ("n_queries", C.c_int64 * 2), ("n_box_checks", C.c_int64 * 2),
        ("n_donated", C.c_int64 * 2), ("n_capped", C.c_int64 * 2), ("n_launches", C.c_int64),
        ("queue_overflow", C.c_int64), ("ms_build", C.c_float), ("ms_sort", C.c_float),
        ("ms_sweep", C.c_float * 2), ("ms_narrow", C.c_float * 2), ("ms_total", C.c_float),
        ("ms_k_sweep_count", C.c_float * 2), ("ms_k_sweep_fill", C.c_float * 2),
        ("ms_k_narrow", C.c_float * 2), ("ms_k_boxes", C.c_float), ("ms_k_gather", C.c_float),
        ("pad_", C.c_float), ("n_records", C.c_int64 * 2), ("grid_cells", (C.c_int32 * 2) * 2),
        ("n_culled", C.c_int64 * 2),
        ("n_round_items", (C.c_int64 * 6) * 2), ("n_round_checks", (C.c_int64 * 5) * 2),
        ("sweep_axis", C.c_int32 * 2), ("next_axis", C.c_int32 * 2), ("n_host_syncs", C.c_int64),
        ("n_records_sent", C.c_int64 * 2), ("n_records_received", C.c_int64 * 2),
        ("ms_exchange", C.c_float), ("ms_k_sort", C.c_float * 2), ("ms_k_expand", C.c_float * 2),
        ("ms_k_cull", C.c_float * 2), ("ms_k_round", (C.c_float * 5) * 2), ("pad2_", C.c_float),
        ("key_bits", C.c_int32 * 2), ("n_skipped", C.c_int64 * 2), ("n_relaunched", C.c_int64),
    ]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            if hasattr(v, "__len__"):
                v = [list(x) if hasattr(x, "__len__") else x for x in v]
            out[name] = v
        return out


class SccdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sccd error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load libsccd_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not found: build it with `make -C {HERE}/csrc` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.sccd_last_error.restype = C.c_char_p
        L.sccd_version.restype = C.c_char_p
        L.sccd_destroy.restype = None
        L.sccd_stats_size.restype = C.c_size_t
        if L.sccd_stats_size() != C.sizeof(Stats):
            raise RuntimeError("capi.Stats does not match sccd_stats of the library")
        _lib = L
    return _lib


def _ptr(a):
    """numpy array -> void*; int -> device pointer; None -> NULL."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def exchange_plan(counts, rank: int, list_: int = 0):
    """sccd_exchange_plan: counts[src, dst] -> (send_off, recv_cnt, recv_off, recv_total) of `rank`."""
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    world = counts.shape[0]
    so, rc, ro = (np.zeros(world, np.uint64) for _ in range(3))
    tot = C.c_uint64(0)
    rcode = load().sccd_exchange_plan(_ptr(counts), C.c_int(list_), C.c_int(rank), C.c_int(world),
                                      _ptr(so), _ptr(rc), _ptr(ro), C.byref(tot))
    if rcode != OK:
        raise SccdError(rcode, "sccd_exchange_plan: bad argument")
    return so, rc, ro, int(tot.value)


class Context:
    """One sccd_ctx: a device + stream + persistent device buffers."""

    def __init__(self, device: int = 0, stream: int = 0):
        self.L = load()
        self._h = C.c_void_p(0)
        rc = self.L.sccd_create(C.c_int(device), C.c_void_p(stream), C.byref(self._h))
        if rc != OK:
            raise SccdError(rc, "sccd_create failed (no usable sm_100 CUDA device?)")
        self.device = device
        self._keep = []

    def close(self):
        if self._h:
            self.L.sccd_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise SccdError(rc, self.L.sccd_last_error(self._h).decode())
        return rc

    # ---- configuration
    def set_memory_limit(self, nbytes: int):
        self._chk(self.L.sccd_set_memory_limit(self._h, C.c_size_t(nbytes)))

    def set_max_pairs_per_chunk(self, n: int):
        self._chk(self.L.sccd_set_max_pairs_per_chunk(self._h, C.c_int64(n)))

    def set_queue_capacity(self, n: int):
        self._chk(self.L.sccd_set_queue_capacity(self._h, C.c_int64(n)))

    def set_grid_cells(self, max_cells: int):
        """0 = automatic (y, z) cell grid, 1 = plain one-axis sweep."""
        self._chk(self.L.sccd_set_grid_cells(self._h, C.c_int(max_cells)))

    def set_shard(self, rank: int, world: int):
        self._chk(self.L.sccd_set_shard(self._h, C.c_int(rank), C.c_int(world)))

    def set_option(self, option: int, value: int):
        """sccd_set_option: OPT_* constants of this module."""
        self._chk(self.L.sccd_set_option(self._h, C.c_int(option), C.c_int64(value)))

    def get_option(self, option: int) -> int:
        v = C.c_int64(0)
        self._chk(self.L.sccd_get_option(self._h, C.c_int(option), C.byref(v)))
        return v.value

    def set_scalar_type(self, scalar: int):
        """F64 (default) or F32 = the reference built with SCALABLE_CCD_USE_DOUBLE=OFF."""
        self._chk(self.L.sccd_set_scalar_type(self._h, C.c_int(scalar)))

    # ---- multi-GPU (one Context per rank; the caller distributes rank 0's id)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        rc = load().sccd_comm_get_unique_id(buf)
        if rc != OK:
            raise SccdError(rc, "sccd_comm_get_unique_id failed (NCCL not loadable?)")
        return buf.raw

    def comm_create(self, unique_id, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), UNIQUE_ID_BYTES) if unique_id else None
        self._chk(self.L.sccd_comm_create(self._h, buf, C.c_int(rank), C.c_int(world)))

    def comm_destroy(self):
        self._chk(self.L.sccd_comm_destroy(self._h))

    def ccd_sharded(self, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True) -> float:
        """Collective ccd() over the communicator's ranks (mesh uploaded on every rank)."""
        t = C.c_double(1.0)
        self._chk(self.L.sccd_ccd_sharded(
            self._h, C.c_double(ms), C.c_int(max_iter), C.c_double(tol),
            C.c_int(int(allow_zero_toi)), C.byref(t)))
        return t.value

    def ccd_sharded_host(self, V0, V1, E, F, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True,
                         sizes=None) -> float:
        """The same from HOST buffers: this rank copies 1 / world of the mesh, NCCL all-gather."""
        nV, nE, nF = sizes if sizes is not None else (V0.shape[0], E.shape[0], F.shape[0])
        t = C.c_double(1.0)
        self._chk(self.L.sccd_ccd_sharded_host(
            self._h, _ptr(V0), _ptr(V1), C.c_int64(nV), _ptr(E), C.c_int64(nE), _ptr(F),
            C.c_int64(nF), C.c_double(ms), C.c_int(max_iter), C.c_double(tol),
            C.c_int(int(allow_zero_toi)), C.byref(t)))
        self.nV, self.nE, self.nF = nV, nE, nF
        return t.value

    # ---- the reference's box builders by name, host arrays in and out
    def build_vertex_boxes(self, V0, V1=None, inflation_radius: float = 0.0):
        """build_vertex_boxes(V0, V1, boxes, r) / build_vertex_boxes(V, boxes, r)."""
        V0 = np.asfortranarray(V0, dtype=np.float64)
        if V1 is not None:
            V1 = np.asfortranarray(V1, dtype=np.float64)
            if V1.shape != V0.shape:
                raise ValueError("V0 and V1 must have the same shape")
        out = np.zeros(max(len(V0), 1), AABB_DTYPE)
        self._chk(self.L.sccd_build_vertex_boxes(
            self._h, _ptr(V0), _ptr(V1), C.c_int64(len(V0)), C.c_double(inflation_radius),
            _ptr(out)))
        return out[:len(V0)]

    def build_element_boxes(self, vertex_boxes, idx):
        """build_edge_boxes (idx n x 2) / build_face_boxes (idx n x 3) from vertex boxes."""
        vb = np.ascontiguousarray(vertex_boxes)
        assert vb.dtype == AABB_DTYPE
        idx = np.asfortranarray(idx, dtype=np.int32)
        out = np.zeros(max(len(idx), 1), AABB_DTYPE)
        self._chk(self.L.sccd_build_element_boxes(
            self._h, _ptr(vb), C.c_int64(len(vb)), _ptr(idx), C.c_int64(len(idx)),
            C.c_int(idx.shape[1]), _ptr(out)))
        return out[:len(idx)]

    # ---- mesh + boxes
    def upload_mesh(self, V0, V1, E, F, sizes=None, host=False):
        """Host numpy arrays (column-major), or raw pointers with sizes=(nV, nE, nF): device
        pointers by default, host pointers (e.g. pinned buffers) with host=True."""
        if sizes is None:
            for a, cols in ((V0, 3), (V1, 3), (E, 2), (F, 3)):
                if a.ndim != 2 or a.shape[1] != cols or not a.flags.f_contiguous:
                    raise ValueError("mesh arrays must be (n, cols) column-major")
            if V0.dtype != np.float64 or V1.dtype != np.float64:
                raise ValueError("V0/V1 must be float64")
            if E.dtype != np.int32 or F.dtype != np.int32:
                raise ValueError("E/F must be int32")
            if V0.shape != V1.shape:
                raise ValueError("V0 and V1 must have the same shape")
            nV, nE, nF = V0.shape[0], E.shape[0], F.shape[0]
            self._keep = [V0, V1, E, F]
            on_dev = 0
        else:
            nV, nE, nF = sizes
            on_dev = 0 if host else 1
        self._chk(self.L.sccd_upload_mesh(
            self._h, _ptr(V0), _ptr(V1), C.c_int64(nV), _ptr(E), C.c_int64(nE), _ptr(F),
            C.c_int64(nF), C.c_int(on_dev)))
        self.nV, self.nE, self.nF = nV, nE, nF

    def update_vertices(self, V0, V1, nV=None, host=False):
        """New positions for the uploaded mesh (same topology): host numpy arrays (column-major
        float64), or raw pointers with nV given -- device pointers unless host=True."""
        if nV is None:
            for a in (V0, V1):
                if a.ndim != 2 or a.shape[1] != 3 or not a.flags.f_contiguous \
                        or a.dtype != np.float64:
                    raise ValueError("V0/V1 must be (n, 3) column-major float64")
            if V0.shape != V1.shape:
                raise ValueError("V0 and V1 must have the same shape")
            nV, on_dev = V0.shape[0], 0
            self._keep_v = [V0, V1]
        else:
            on_dev = 0 if host else 1
        self._chk(self.L.sccd_update_vertices(
            self._h, _ptr(V0), _ptr(V1), C.c_int64(nV), C.c_int(on_dev)))

    def build_boxes(self, inflation_radius: float = 0.0):
        self._chk(self.L.sccd_build_boxes(self._h, C.c_double(inflation_radius)))

    def get_boxes(self):
        """(vertex, edge, face) boxes in the reference's host AABB format."""
        out = []
        for which, n in ((0, self.nV), (1, self.nE), (2, self.nF)):
            a = np.zeros(max(n, 1), AABB_DTYPE)
            self._chk(self.L.sccd_get_boxes(self._h, C.c_int(which), _ptr(a)))
            out.append(a[:n])
        return tuple(out)

    def set_boxes(self, a, b=None, sort_axis: int = 0) -> int:
        """Caller-made boxes (AABB_DTYPE arrays): BroadPhase::build(boxes[, boxesB]) /
        sort_and_sweep(boxes..., axis).  Sweep them with kind=BOXES.  Returns the next sort
        axis the reference's sort_and_sweep would report."""
        a = np.ascontiguousarray(a)
        assert a.dtype == AABB_DTYPE
        if b is not None:
            b = np.ascontiguousarray(b)
            assert b.dtype == AABB_DTYPE
        nxt = C.c_int(0)
        self._chk(self.L.sccd_set_boxes(
            self._h, _ptr(a), C.c_int64(len(a)), _ptr(b), C.c_int64(len(b) if b is not None else 0),
            C.c_int(sort_axis), C.byref(nxt)))
        self._keep_boxes = (a, b)
        return nxt.value

    # ---- broad phase
    def broad_phase(self, kind: int, want_pairs: bool = True):
        """detect_overlaps(): all pairs on the host, (n, 2) int32."""
        n = C.c_int64(0)
        self._chk(self.L.sccd_broad_phase(self._h, C.c_int(kind), None, C.c_int64(0), C.byref(n)))
        if not want_pairs:
            return n.value
        out = np.empty((max(n.value, 1), 2), np.int32)
        self._chk(self.L.sccd_broad_phase(
            self._h, C.c_int(kind), _ptr(out), C.c_int64(n.value), C.byref(n)))
        return out[:n.value]

    def broad_phase_begin(self, kind: int):
        self._chk(self.L.sccd_broad_phase_begin(self._h, C.c_int(kind)))

    def broad_phase_partial(self):
        """-> (device pointer, n_pairs) of the next chunk."""
        p = C.c_void_p(0)
        n = C.c_int64(0)
        self._chk(self.L.sccd_broad_phase_partial(self._h, C.byref(p), C.byref(n)))
        return (p.value or 0), n.value

    def broad_phase_is_complete(self) -> bool:
        return bool(self._chk(self.L.sccd_broad_phase_is_complete(self._h)))

    # ---- narrow phase
    def narrow_phase(self, kind, d_pairs: int, n: int, ms=0.0, max_iter=-1, tol=1e-6,
                     allow_zero_toi=True, toi=1.0, d_toi_per_query: int = 0) -> float:
        t = C.c_double(toi)
        self._chk(self.L.sccd_narrow_phase(
            self._h, C.c_int(kind), C.c_void_p(d_pairs), C.c_int64(n), C.c_double(ms),
            C.c_int(max_iter), C.c_double(tol), C.c_int(int(allow_zero_toi)), C.byref(t),
            C.c_void_p(d_toi_per_query)))
        return t.value

    def narrow_phase_queries(self, kind, queries, n=None, ms=0.0, max_iter=-1, tol=1e-6,
                             allow_zero_toi=True, toi=1.0, d_toi_per_query: int = 0) -> float:
        """queries: host (n, 24) float64 array, or a device pointer with n given."""
        on_dev = isinstance(queries, (int, np.integer))
        if not on_dev:
            queries = np.ascontiguousarray(queries, dtype=np.float64).reshape(-1, 24)
            n = len(queries)
        t = C.c_double(toi)
        self._chk(self.L.sccd_narrow_phase_queries(
            self._h, C.c_int(kind), _ptr(queries), C.c_int64(n), C.c_int(int(on_dev)),
            C.c_double(ms), C.c_int(max_iter), C.c_double(tol), C.c_int(int(allow_zero_toi)),
            C.byref(t), C.c_void_p(d_toi_per_query)))
        return t.value

    def narrow_phase_checks(self):
        """-> (device pointer, n): per-query box counters of the last capped narrow phase."""
        p = C.c_void_p(0)
        n = C.c_int64(0)
        self._chk(self.L.sccd_narrow_phase_checks(self._h, C.byref(p), C.byref(n)))
        return (p.value or 0), n.value

    # ---- pipelines
    def ccd(self, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True) -> float:
        t = C.c_double(1.0)
        self._chk(self.L.sccd_ccd(
            self._h, C.c_double(ms), C.c_int(max_iter), C.c_double(tol),
            C.c_int(int(allow_zero_toi)), C.byref(t)))
        return t.value

    def ccd_host(self, V0, V1, E, F, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True,
                 sizes=None) -> float:
        """ccd() with HOST buffers (numpy arrays, or raw host pointers + sizes=(nV,nE,nF))."""
        nV, nE, nF = sizes if sizes is not None else (V0.shape[0], E.shape[0], F.shape[0])
        t = C.c_double(1.0)
        self._chk(self.L.sccd_ccd_host(
            self._h, _ptr(V0), _ptr(V1), C.c_int64(nV), _ptr(E), C.c_int64(nE),
            _ptr(F), C.c_int64(nF), C.c_double(ms), C.c_int(max_iter), C.c_double(tol),
            C.c_int(int(allow_zero_toi)), C.byref(t)))
        self.nV, self.nE, self.nF = nV, nE, nF
        return t.value

    def ccd_collisions(self, ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True):
        """TOI_PER_QUERY build of ccd(): -> (toi, vf (ids, toi), ee (ids, toi))."""
        t = C.c_double(1.0)
        nvf, nee = C.c_int64(0), C.c_int64(0)
        self._chk(self.L.sccd_ccd_collisions(
            self._h, C.c_double(ms), C.c_int(max_iter), C.c_double(tol),
            C.c_int(int(allow_zero_toi)), C.byref(t), None, None, C.c_int64(0), C.byref(nvf),
            C.byref(nee)))
        k = nvf.value + nee.value
        ids = np.empty((max(k, 1), 2), np.int32)
        tois = np.empty(max(k, 1), np.float64)
        self._chk(self.L.sccd_get_collisions(   # one pipeline pass: fetch what it kept
            self._h, _ptr(ids), _ptr(tois), C.c_int64(k), C.byref(nvf), C.byref(nee)))
        a = nvf.value
        return t.value, (ids[:a], tois[:a]), (ids[a:k], tois[a:k])

    def ipc_ccd_strategy(self, min_distance=0.0, max_iter=-1, tol=1e-6) -> float:
        t = C.c_double(1.0)
        self._chk(self.L.sccd_ipc_ccd_strategy(
            self._h, C.c_double(min_distance), C.c_int(max_iter), C.c_double(tol), C.byref(t)))
        return t.value

    def stats(self) -> dict:
        s = Stats()
        self._chk(self.L.sccd_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._chk(self.L.sccd_reset_stats(self._h))

    def measure_fp64_peak(self) -> float:
        """thread-level DFMA per second of this device (register-only micro-benchmark)"""
        v = C.c_double(0.0)
        self._chk(self.L.sccd_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def synchronize(self):
        self._chk(self.L.sccd_synchronize(self._h))
