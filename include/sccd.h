/*
 * sccd.h -- C ABI of the B200-native continuous-collision-detection hot path.
 *
 * This is the drop-in boundary for the Scalable-CCD GPU path (reference paths are
 * relative to /root/reference/src/scalable_ccd).  The reference exposes a C++17
 * API with Eigen matrices; each entry point below names the reference interface it
 * replaces.  The header-only C++ shim include/sccd.hpp re-creates the reference's
 * names (scalable_ccd::cuda::ccd, BroadPhase, narrow_phase, ...) on top of this ABI.
 *
 * Conventions (identical to the reference):
 *   - V0, V1 : nV x 3 float64, COLUMN-major (Eigen default; cuda/utils/device_matrix.cuh:77-81)
 *   - E      : nE x 2 int32,   F : nF x 3 int32, column-major
 *   - pairs  : int2 (a, b); vertex-face = (vertex id, face id), edge-edge =
 *              (min edge id, max edge id)          (cuda/broad_phase/sweep.cu:152-164)
 *   - toi    : earliest time of impact in [0, 1]; 1.0 = no collision (cuda/ccd.cu:125)
 *   - Scalar : double (the reference's default SCALABLE_CCD_USE_DOUBLE build)
 *
 * All functions return SCCD_OK (0) or a negative error code; sccd_last_error()
 * gives the message (the C++ shim rethrows it as std::runtime_error, the
 * reference's convention, cuda/utils/assert.cuh:12-28).  There is no CPU fallback:
 * every call fails with SCCD_ERR_CUDA if no sm_100 device is usable.
 *
 * A context is bound to one device and one stream and is not thread-safe (neither
 * is the reference: global __constant__ CONFIG, cuda/narrow_phase/root_finder.cu:19).
 * Different contexts may be used concurrently from different threads.
 */
#ifndef SCCD_H
#define SCCD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCCD_OK 0
#define SCCD_ERR_CUDA (-1)   /* CUDA runtime error / no usable device        */
#define SCCD_ERR_ARG (-2)    /* bad argument                                 */
#define SCCD_ERR_STATE (-3)  /* call order (e.g. broad phase before build)   */
#define SCCD_ERR_MEMORY (-4) /* memory budget too small for a single box     */

#define SCCD_VF 0 /* vertex-face  (two lists: vertices, faces) */
#define SCCD_EE 1 /* edge-edge    (single list)                */

/* Scalar type of the computation: the reference's SCALABLE_CCD_USE_DOUBLE switch
 * (scalar.hpp:13-19, CMakeLists.txt:69).  The C ABI stays double either way. */
#define SCCD_F64 0 /* default build of the reference                                   */
#define SCCD_F32 1 /* its float build: inputs cast to float, nextafterf, float solver  */

typedef struct sccd_ctx sccd_ctx;

/* int2 of the reference's overlap lists (cuda/broad_phase/broad_phase.cuh:58-60). */
typedef struct {
    int32_t a, b;
} sccd_pair;

/* Same 64-byte layout as the reference's cuda::AABB in the double build
 * (cuda/broad_phase/aabb.cuh:12-96). */
typedef struct {
    double min[3];
    double max[3];
    int32_t vertex_ids[3];
    int32_t element_id;
} sccd_aabb;

/* Counters and device timings (ms, CUDA events on the context's stream) of the
 * most recent pipeline call.  Replaces the reference's SCALABLE_CCD_*_PROFILE_POINT
 * instrumentation (utils/profiler.hpp:15-18). */
typedef struct {
    int64_t n_boxes[2];      /* boxes swept: [VF list (nV+nF), EE list (nE)]            */
    int64_t n_pairs[2];      /* candidate pairs emitted                                 */
    int64_t n_candidates[2]; /* f32-prefilter survivors that reached the exact test     */
    int64_t n_queries[2];    /* narrow-phase queries run                                */
    int64_t n_box_checks[2]; /* inclusion-function evaluations (ccd_kernel bodies)      */
    int64_t n_donated[2];    /* sub-boxes handed on to a later round through the item list */
    int64_t n_capped[2];     /* queries that hit max_iter (conservatively accepted)     */
    int64_t n_launches;      /* kernels this library launched in the call               */
    int64_t queue_overflow;  /* 1 if the bounded item list was ever full (work kept local) */
    float ms_build;          /* AABB build                                              */
    float ms_sort;           /* radix sort + record gather                              */
    float ms_sweep[2];       /* sweep count + scan + fill                               */
    float ms_narrow[2];      /* narrow phase                                            */
    float ms_total;          /* whole call, device time                                 */
    /* single-kernel device times (summed over chunks), for roofline arithmetic */
    float ms_k_sweep_count[2];
    float ms_k_sweep_fill[2];
    float ms_k_narrow[2];
    float ms_k_boxes;        /* vertex + element box kernels                            */
    float ms_k_gather;       /* both record gathers                                     */
    float pad_;
    int64_t n_records[2];    /* sweep records = boxes replicated into the (y,z) cells   */
    int32_t grid_cells[2][2]; /* (sy, sz) cell grid chosen for each list                */
    int64_t n_culled[2];     /* queries answered "no collision" by the separating-axis cull */
    /* narrow-phase load balance: items the solver rounds read (round 0 = queries that survived
     * the cull; [5] = handed on by the last round) and the box checks of each round */
    int64_t n_round_items[2][6];
    int64_t n_round_checks[2][5];
    int32_t sweep_axis[2];   /* axis each list was swept along (SCCD_OPT_SWEEP_AXIS)            */
    int32_t next_axis[2];    /* variance argmax sort_and_sweep would hand back                  */
    int64_t n_host_syncs;    /* host <-> device synchronisations inside the call                */
    /* multi-GPU (sccd_ccd_sharded): records this rank sent to / received from the others   */
    int64_t n_records_sent[2];
    int64_t n_records_received[2];
    float ms_exchange;       /* record exchange (NCCL send / recv), device time                 */
    float ms_k_sort[2];      /* radix sort passes of each list                                  */
    float ms_k_expand[2];    /* record expansion (count + scan + fill) of each list             */
    float ms_k_cull[2];      /* separating-axis cull kernel                                     */
    float ms_k_round[2][5];  /* solver rounds (SCCD_OPT_PROFILE only)                           */
    float pad2_;
    int32_t key_bits[2];     /* sort-key bits in use (cell + quantised major axis) per list       */
    int64_t n_skipped[2];    /* cull survivors never started: their toi lower bound was not below
                                the running earliest toi (shared-bound mode, max_iter < 0)        */
    int64_t n_relaunched;    /* narrow-phase batches whose solver kernels were chosen for the
                                survivor-list length of the previous batch and had to be
                                launched again because the length class changed              */
} sccd_stats;
/* sizeof(sccd_stats) of the library (binding sanity check) */
size_t sccd_stats_size(void);

/* ---- context ------------------------------------------------------------------ */

/* stream: a cudaStream_t (may be NULL = the legacy default stream).  Work is only
 * ever enqueued on this stream. */
int sccd_create(int device, void* stream, sccd_ctx** out);
void sccd_destroy(sccd_ctx* ctx);
const char* sccd_last_error(const sccd_ctx* ctx);

/* MemoryHandler::memory_limit_GB (cuda/memory_handler.hpp:33; ccd.cuh:38).
 * 0 = 95 % of free device memory (memory_handler.cpp:19-29).  Bounds the candidate
 * pair buffer; larger overlap sets are produced in owner-range chunks. */
int sccd_set_memory_limit(sccd_ctx* ctx, size_t bytes);

/* Cap on pairs per sccd_broad_phase_partial() call (0 = derive from the memory
 * limit).  Test hook for the chunking path (MAX_OVERLAP_CUTOFF semantics,
 * cuda/broad_phase/broad_phase.cu:194-207). */
int sccd_set_max_pairs_per_chunk(sccd_ctx* ctx, int64_t max_pairs);

/* Capacity (items of 64 bytes) of each of the two bounded narrow-phase item lists
 * (0 = default: one per query, 1 Mi .. 16 Mi).  Replaces MemoryHandler::MAX_UNIT_SIZE
 * (cuda/memory_handler.cpp:81-122).  A full list never drops work: the lane keeps its tree. */
int sccd_set_queue_capacity(sccd_ctx* ctx, int64_t items);

/* Upper bound on the number of (y, z) sweep cells per list: 0 = automatic (about twice the
 * mean box extent per cell, <= 2^20 cells), 1 = plain one-axis sweep.  The overlap set does
 * not depend on it; it only changes how many candidates the sweep has to test. */
int sccd_set_grid_cells(sccd_ctx* ctx, int max_cells);

/* The reference's compile-time Scalar (scalar.hpp:13-19) as a run-time mode.  SCCD_F32 gives the
 * results of the reference built with SCALABLE_CCD_USE_DOUBLE=OFF: vertices are cast to float
 * (cuda/broad_phase/aabb.cu:124-128, device_matrix.cuh:21-27), boxes are made with nextafterf and
 * float adds, ms / tolerance / toi are rounded to float on the way in, and the narrow phase
 * runs in float with the arithmetic nvcc gives the reference under --use_fast_math
 * (CMakeLists.txt:219-230; flush-to-zero, a / b = a * rcp(b)) and the float error filters
 * (root_finder.cu:102-119).  Arguments and results stay `double` in this ABI: a float value is
 * exact in a double.  Boxes have to be rebuilt after a change. */
int sccd_set_scalar_type(sccd_ctx* ctx, int type);

/* Tuning and test knobs (the reference has none of them; its equivalents are compile-time
 * constants).  Initial values come from the SCCD_* environment variables read once in
 * sccd_create; nothing on the hot path reads the environment.  SCCD_ERR_ARG for an unknown
 * option or a value out of range. */
#define SCCD_OPT_NARROW_CULL 1      /* cull in front of the solver: 1 (default) = separating-axis
                                       test + the solver's own first box check (root box) on the
                                       queries that pass it; 2 = the separating-axis test alone;
                                       3 = as 1 without the float pre-test in front of the double
                                       test (A/B); 0 = none.  Results never depend on it.        */
#define SCCD_OPT_NARROW_FLAGS 2     /* narrow-phase scheduling knobs (budgets, refill, kernel
                                       choice; see csrc/narrow.cu); results never depend on it */
#define SCCD_OPT_NARROW_FLAGS_EE 3  /* the same for the edge-edge pass alone; < 0 = follow (2)  */
#define SCCD_OPT_NARROW_MAX_DEPTH 4 /* bisection levels a walk tracks before handing on, 2..128 */
#define SCCD_OPT_MAX_ITER_MODE 5    /* a query that reaches max_iter >= 0:
                                       0 (default) its remaining boxes are ACCEPTED at their
                                         t_lo -- never later than the exact answer;
                                       1 they are DROPPED, as in the reference
                                         (root_finder.cu:303-305) -- may miss a collision      */
#define SCCD_OPT_KEY_STEPS 6        /* log2 of the major-axis quantisation steps per record     */
#define SCCD_OPT_GRID_SCALE_MILLI 7 /* cell edge in mean box extents, x 1000 (default 3000)     */
#define SCCD_OPT_GRID_REPL_MILLI 8  /* records per box above which the grid is coarsened, x1000 */
#define SCCD_OPT_NARROW_SOLVER 11   /* how the solver schedules the bisection trees:
                                       0 = by the length of the work list -- one lane per tree,
                                           cut into rounds (long lists); one warp per tree in a
                                           persistent work queue (short lists);
                                       1 = rounds for short lists as well;
                                       4 / 8 = that many lanes per tree, one kernel per round.
                                       Results never depend on it.                               */
#define SCCD_OPT_CONCURRENT_PASSES 12 /* plain ccd(): 1 = the edge-edge solver starts as soon as its
                                       own list is ready instead of after the vertex-face solver
                                       (whose earliest toi it would inherit as its pruning bound).
                                       Both publish to the same toi word, so each still profits
                                       from the other's finds; the result is the same, the work
                                       can be more.  0 = in sequence (default).                  */
#define SCCD_OPT_SWEEP_STAGED 13     /* sweep count pass: 1 = the prefilter stream of a tile (keys +
                                       f32 yz of the next 1024 records) is staged in shared memory
                                       by bulk async copies (cp.async.bulk + mbarrier) before the
                                       window loop; 0 (default) = the loop reads it through L1.
                                       Same pairs either way.                                    */
#define SCCD_OPT_REUSE_GRID 14       /* frame-to-frame (default 1): when a build has the list sizes,
                                       sweep axis and scalar type of the previous one, the cell
                                       grid and key quantisation are chosen from the previous
                                       build's box statistics -- one host sync less per step --
                                       while this build's statistics are computed off the
                                       critical path for the next one.  Any grid gives the same
                                       overlap set; 0 = wait for this build's own statistics.    */
#define SCCD_OPT_PROFILE 10         /* 1: time every kernel with its own event pair
                                       (sccd_stats.ms_k_*); 2: the same with BOTH lists on the
                                       caller's stream, so that a pair brackets its kernel alone
                                       (roofline measurements); 0 (default): total only         */
#define SCCD_OPT_SWEEP_AXIS 9       /* axis the mesh pipeline sorts and sweeps along: 0 (default,
                                       the reference's GPU path, aabb.cu:86), 1, 2, or -1 = the
                                       axis sort_and_sweep would hand back for the NEXT call --
                                       argmax of the box-centre variance of the previous build
                                       (sort_and_sweep.cpp:176-195).  The overlap set and every
                                       TOI are independent of it.                               */
int sccd_set_option(sccd_ctx* ctx, int option, int64_t value);
int sccd_get_option(const sccd_ctx* ctx, int option, int64_t* value);

/* Multi-GPU sharding: this context makes, sorts and sweeps only the records of the rank-th of
 * `world` contiguous (y, z) cell ranges of each list (ranges balanced by record count; every
 * rank derives the same ranges from its own copy of the boxes, so there is no exchange) and
 * therefore emits a disjoint part of the global pair list; the parts concatenated in rank
 * order are the single-device list.  Lists with too few cells fall back to owner slices of
 * the whole sorted list, balanced by sweep-window length.  Takes effect at the next
 * sccd_build_boxes / sccd_set_boxes / sccd_broad_phase_begin.  rank 0 / world 1 = everything. */
int sccd_set_shard(sccd_ctx* ctx, int rank, int world);

/* ---- multi-GPU: one context (process or thread) per GPU ------------------------------ */
/* The reference is single-GPU (its _multigpu directory is dead code that replicates every box
 * and merges on the host, cuda/broad_phase/_multigpu/broad_phase.cu:69-116).  Here the ranks of
 * a communicator split one ccd() call:
 *   1. every rank makes the boxes of 1 / world of the elements and one 8-byte (key, index)
 *      record per (box, cell);
 *   2. records are exchanged by OWNING CELL RANGE (contiguous ranges of the (y, z) cell grid,
 *      balanced by record count; NCCL send / recv over NVLink) -- cells are independent sweep
 *      domains, so there is no halo;
 *   3. the receiver sorts its records and rebuilds their exact boxes from the mesh (replicated,
 *      48 B / vertex), sweeps them, and solves the pairs it found;
 *   4. one all-reduce(min) of the earliest TOI.
 * The pair lists of the ranks are disjoint and their union is the single-GPU overlap set; the
 * TOI is the single-GPU TOI.
 *
 * NCCL (libnccl.so.2) is loaded at run time by sccd_comm_create, so single-GPU users need none.
 * The caller distributes the 128-byte id of rank 0 by its own means (MPI_Bcast, a file, a
 * torch.distributed broadcast ...).  world == 1 needs no NCCL and no id (id may be NULL). */
#define SCCD_UNIQUE_ID_BYTES 128
int sccd_comm_get_unique_id(void* id_out /* SCCD_UNIQUE_ID_BYTES */);
int sccd_comm_create(sccd_ctx* ctx, const void* id, int rank, int world);
int sccd_comm_destroy(sccd_ctx* ctx);

/* ccd() over the communicator's GPUs: collective -- every rank calls it with the same
 * arguments after uploading the SAME mesh (sccd_upload_mesh / sccd_update_vertices); every rank
 * receives the same *toi.  sccd_get_stats afterwards reports this rank's share. */
int sccd_ccd_sharded(
    sccd_ctx* ctx, double min_distance, int max_iter, double tol, int allow_zero_toi,
    double* toi);

/* The same from HOST buffers: every rank holds the whole mesh on its host (the arguments are
 * host pointers, identical content on every rank), copies only 1 / world of it over its own
 * PCIe link, and the slices are all-gathered over NVLink. */
int sccd_ccd_sharded_host(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, double min_distance, int max_iter, double tol,
    int allow_zero_toi, double* toi);

/* Host arithmetic of the record exchange, exposed for tests that have no GPU: from the matrix
 * counts[src * world + dst] (records rank src holds for rank dst, one list) it gives rank
 * `rank` the offsets of its per-destination parts in its send buffer, and the count / offset of
 * what it receives from every source (sources land in rank order).  Arrays of `world` words;
 * any output may be NULL. */
int sccd_exchange_plan(
    const uint64_t* counts, int list, int rank, int world, uint64_t* send_off, uint64_t* recv_cnt,
    uint64_t* recv_off, uint64_t* recv_total);

/* ---- mesh + boxes --------------------------------------------------------------- */

/* DeviceMatrix uploads of ccd() (cuda/ccd.cu:103-106).  on_device != 0: the four
 * pointers are device pointers valid on the context's device (no copy is made of V0/V1
 * beyond the packed vertex table). */
int sccd_upload_mesh(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, const int32_t* E,
    int64_t nE, const int32_t* F, int64_t nF, int on_device);

/* Frame-to-frame reuse: new vertex positions for the mesh of the last
 * sccd_upload_mesh -- same vertex count, E and F stay where that call put them (on the device),
 * so a simulation step moves 48 B per vertex instead of the whole mesh.  The reference has no
 * such entry: its ccd() uploads all four matrices on every call (cuda/ccd.cu:103-106).
 * on_device as in sccd_upload_mesh.  Boxes have to be rebuilt (the pipeline calls do that). */
int sccd_update_vertices(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, int on_device);

/* build_vertex_boxes + build_edge_boxes + build_face_boxes + the three DeviceAABBs
 * constructors (cuda/broad_phase/aabb.cu:75-229), on the device: boxes are built
 * from the uploaded mesh, keyed on min.x and radix-sorted (vertices+faces as one
 * tagged list, edges as another). */
int sccd_build_boxes(sccd_ctx* ctx, double inflation_radius);

/* Boxes in the reference's host AABB format and element order, for parity checks
 * (what build_*_boxes return; cuda/broad_phase/aabb.cuh:150-188).
 * which: 0 vertices, 1 edges, 2 faces.  out: host array of nV / nE / nF boxes. */
int sccd_get_boxes(sccd_ctx* ctx, int which, sccd_aabb* out);

/* The reference's box builders by name, for callers that make and keep the boxes themselves
 * (tests/test_broad_phase.cu:88-91); host arrays in and out, computed on the device, no mesh
 * or context state involved.
 *   build_vertex_boxes(V0, V1, boxes, r)  cuda/broad_phase/aabb.cuh:160-170, aabb.cu:146-184
 *   build_vertex_boxes(V, boxes, r)       aabb.cuh:150-153 (V1 == NULL)
 * V0 / V1: nV x 3 column-major.  ids as the reference sets them: (i, -i-1, -i-1), element i. */
int sccd_build_vertex_boxes(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, double inflation_radius,
    sccd_aabb* out);
/*   build_edge_boxes(vertex_boxes, E, edge_boxes)   aabb.cuh:172-179, aabb.cu:186-206
 *   build_face_boxes(vertex_boxes, F, face_boxes)   aabb.cuh:181-188, aabb.cu:208-229
 * idx: n x verts_per_element (2 or 3) column-major int32.  Box = union of the vertex boxes, ids
 * (e0, e1, -e0-1) / (f0, f1, f2), element id = row.  SCCD_ERR_ARG if an index is out of range. */
int sccd_build_element_boxes(
    sccd_ctx* ctx, const sccd_aabb* vertex_boxes, int64_t nV, const int32_t* idx, int64_t n,
    int verts_per_element, sccd_aabb* out);

/* DeviceAABBs(boxes) + BroadPhase::build(boxes) / build(boxesA, boxesB) for caller-made
 * boxes (cuda/broad_phase/aabb.cu:75-111, broad_phase.cu:29-101; this is how
 * tests/test_broad_phase.cu:88-104 drives the broad phase), and the box-list half of the
 * CPU entry points sort_and_sweep(boxes, axis, overlaps) / sort_and_sweep(boxesA, boxesB,
 * axis, overlaps) (broad_phase/sort_and_sweep.hpp:24-42).  Host arrays in the reference's
 * AABB layout; b == NULL / nb == 0 selects the single-list form.  sort_axis (0, 1, 2) is the
 * axis to sweep along -- the overlap set does not depend on it.  *next_axis (may be NULL)
 * receives the axis sort_and_sweep would return: argmax of the variance of the box centres
 * (sort_and_sweep.cpp:176-195).  The list is then swept with kind = SCCD_BOXES. */
#define SCCD_BOXES 2
int sccd_set_boxes(
    sccd_ctx* ctx, const sccd_aabb* a, int64_t na, const sccd_aabb* b, int64_t nb,
    int sort_axis, int* next_axis);

/* ---- broad phase ---------------------------------------------------------------- */

/* BroadPhase::build (cuda/broad_phase/broad_phase.cu:29-101) for kind = SCCD_VF
 * (vertex list, face list) or SCCD_EE (edge list).  Requires sccd_build_boxes. */
int sccd_broad_phase_begin(sccd_ctx* ctx, int kind);

/* BroadPhase::detect_overlaps_partial (broad_phase.cu:121-224): next chunk of the
 * deterministic pair list, ordered by (owner position, candidate position) in the
 * sorted list.  *d_pairs is device memory owned by the context, valid until the
 * next call on it. */
int sccd_broad_phase_partial(sccd_ctx* ctx, const sccd_pair** d_pairs, int64_t* n_pairs);

/* BroadPhase::is_complete (broad_phase.cuh:50).  Returns 1 / 0, negative on error. */
int sccd_broad_phase_is_complete(sccd_ctx* ctx);

/* BroadPhase::detect_overlaps (broad_phase.cu:226-252): run all chunks and copy the
 * pairs to host.  out may be NULL (count only); at most cap pairs are written;
 * *n_total receives the full count. */
int sccd_broad_phase(
    sccd_ctx* ctx, int kind, sccd_pair* out, int64_t cap, int64_t* n_total);

/* ---- narrow phase --------------------------------------------------------------- */

/* narrow_phase<is_vf> (cuda/narrow_phase/narrow_phase.cuh:30-46) over a device pair
 * list against the uploaded mesh.
 *   ms, max_iter, tol, allow_zero_toi: as in the reference (max_iter < 0 = no cap).
 *   toi_inout (host): running earliest toi, lowered in place (ccd.cu:125-143).
 *   d_toi_per_query (device, n doubles, may be NULL): SCALABLE_CCD_TOI_PER_QUERY
 *     semantics -- every query is solved against its OWN bound (init +inf) and its
 *     earliest toi is stored; hit <=> value < 1 (narrow_phase.cu:69-82).  With NULL
 *     the shared running toi prunes all queries (default reference build).
 * Deviation (documented in DESIGN.md): a query that exceeds max_iter >= 0 has its
 * remaining boxes ACCEPTED at their t_lo instead of silently dropped
 * (root_finder.cu:303-305), so the result is never later than the reference's. */
int sccd_narrow_phase(
    sccd_ctx* ctx, int kind, const sccd_pair* d_pairs, int64_t n, double ms, int max_iter,
    double tol, int allow_zero_toi, double* toi_inout, double* d_toi_per_query);

/* Root finder on explicit query arrays: ccd<is_vf>(d_data, ...)
 * (cuda/narrow_phase/root_finder.cuh:41-50).  queries: n x 24 doubles laid out like
 * the first 192 bytes of CCDData (ccd_data.cuh:8-18): v0s v1s v2s v3s v0e v1e v2e v3e.
 * on_device != 0: queries is a device pointer. */
int sccd_narrow_phase_queries(
    sccd_ctx* ctx, int kind, const double* queries, int64_t n, int on_device, double ms,
    int max_iter, double tol, int allow_zero_toi, double* toi_inout, double* d_toi_per_query);

/* Per-query box counters of the most recent sccd_narrow_phase / sccd_narrow_phase_queries call
 * that ran with max_iter >= 0: the reference's CCDData::nbr_checks (ccd_data.cuh:8-18,
 * root_finder.cu:288-289).  *d_checks: device array of *n counters owned by the context, valid
 * until its next narrow-phase call (NULL / 0 if the last call had max_iter < 0).  A query can
 * only have been cut short by the cap if its counter exceeds max_iter + 1; the TOI of every
 * other query is exactly the uncapped answer. */
int sccd_narrow_phase_checks(sccd_ctx* ctx, const uint32_t** d_checks, int64_t* n);

/* ---- pipelines ------------------------------------------------------------------ */

/* ccd() without the upload (cuda/ccd.cu:108-146): boxes -> VF broad+narrow ->
 * EE broad+narrow on the mesh already uploaded.  *toi (host) receives the result. */
int sccd_ccd(
    sccd_ctx* ctx, double min_distance, int max_iter, double tol, int allow_zero_toi,
    double* toi);

/* ccd() of the TOI_PER_QUERY build (ccd.cuh:35-37): additionally returns the
 * collisions (aid, bid, toi < 1).  ids/tois may be NULL; at most cap are written;
 * VF collisions come first, then EE; *n_vf / *n_ee receive the full counts. */
int sccd_ccd_collisions(
    sccd_ctx* ctx, double min_distance, int max_iter, double tol, int allow_zero_toi,
    double* toi, sccd_pair* ids, double* tois, int64_t cap, int64_t* n_vf, int64_t* n_ee);

/* The collisions of the most recent sccd_ccd_collisions on this context, without running
 * anything again: call sccd_ccd_collisions with cap = 0 to learn the counts, then fetch.
 * (copy_out_collisions, cuda/narrow_phase/narrow_phase.cu:84-103, hands the reference's caller
 * one vector after one pass; this keeps it one pass here as well.) */
int sccd_get_collisions(
    sccd_ctx* ctx, sccd_pair* ids, double* tois, int64_t cap, int64_t* n_vf, int64_t* n_ee);

/* Whole ccd() including the host->device upload (cuda/ccd.cuh:26-38). */
int sccd_ccd_host(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, const int32_t* E,
    int64_t nE, const int32_t* F, int64_t nF, double min_distance, int max_iter, double tol,
    int allow_zero_toi, double* toi);

/* ipc_ccd_strategy() without the upload (cuda/ipc_ccd_strategy.cu:108-152). */
int sccd_ipc_ccd_strategy(
    sccd_ctx* ctx, double min_distance, int max_iter, double tol, double* toi);

/* ---- introspection -------------------------------------------------------------- */
int sccd_get_stats(const sccd_ctx* ctx, sccd_stats* out);
/* Pipeline calls (sccd_ccd*, sccd_ipc_ccd_strategy) start from zeroed stats by themselves;
 * callers that compose the phase-level entry points reset them here. */
int sccd_reset_stats(sccd_ctx* ctx);
int sccd_synchronize(sccd_ctx* ctx);
/* Measures the device's FP64 pipe rate (thread-level DFMA per second) with a register-only
 * micro-benchmark: the compute roofline the narrow phase is reported against (the reference
 * publishes none). */
int sccd_measure_fp64_peak(sccd_ctx* ctx, double* dfma_per_second);
/* "major.minor.patch sm_100a" */
const char* sccd_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SCCD_H */
