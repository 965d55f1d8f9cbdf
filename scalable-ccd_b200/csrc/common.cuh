// Shared declarations of the B200 CCD hot path (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/sccd.h"

namespace sccd {

// ---- errors ---------------------------------------------------------------------
struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check(cudaError_t e, const char* what, const char* file, int line)
{
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(
            buf, sizeof(buf), "%s: %s (%s) at %s:%d", what, cudaGetErrorName(e),
            cudaGetErrorString(e), file, line);
        throw CudaError(buf);
    }
}
#define SCCD_CUDA(expr) ::sccd::check((expr), #expr, __FILE__, __LINE__)

// ---- grow-only device buffer (kept across calls: frame-to-frame reuse) -----------
// bumped by every (re)allocation: lets the context reuse its last cudaMemGetInfo() answer
// (the call takes milliseconds once gigabytes are mapped) while nothing was allocated
// (process-wide and atomic: contexts on different threads allocate concurrently)
inline std::atomic<unsigned long long>& alloc_epoch()
{
    static std::atomic<unsigned long long> e { 0 };
    return e;
}

struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void release()
    {
        if (ptr)
            cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    // Contents are NOT preserved when the buffer grows.
    void* reserve(size_t bytes)
    {
        if (bytes > cap) {
            release();
            size_t want = bytes + bytes / 8 + 256;
            if (cudaMalloc(&ptr, want) != cudaSuccess) {
                (void)cudaGetLastError();
                want = bytes;
                SCCD_CUDA(cudaMalloc(&ptr, want));
            }
            cap = want;
            alloc_epoch().fetch_add(1, std::memory_order_relaxed);
        }
        return ptr;
    }
    template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

// ---- HBM layouts -------------------------------------------------------------------
// Per-vertex two-frame record gathered by the narrow phase: (x0 y0 z0 x1 y1 z1).
struct __align__(16) VertexRec {
    double p0[3];
    double p1[3];
};
static_assert(sizeof(VertexRec) == 48, "VertexRec");

// Exact box record, 64 bytes split in three arrays so every access is one aligned
// vector load: X = (xmin, xmax), YZ = (ymin, zmin, ymax, zmax), ID = (v0, v1, v2, elem).
// elem is the element id, flipped (-id-1) for list A of a two-list sweep exactly as the
// reference does (cuda/broad_phase/broad_phase.cu:20-26).
struct BoxArrays {
    double2* x = nullptr;
    double4* yz = nullptr;
    int4* id = nullptr;
};

// Uniform (y, z) cell grid laid over a list.  Every box is replicated into each cell its
// closed [ymin,ymax] x [zmin,zmax] range touches and the list is sorted on (cell, xmin), so
// the sweep window of a box only holds boxes of the same cell.  A pair is reported in ONE
// cell only: the cell of (max(ymin_a, ymin_b), max(zmin_a, zmin_b)), which both boxes touch
// whenever they overlap (cell_index() is monotone).  sy = sz = 1 is the plain 1-axis sweep.
struct GridParams {
    double y0 = 0, z0 = 0, inv_hy = 0, inv_hz = 0;
    int sy = 1, sz = 1;
    // Multi-GPU: only records of linear cells [cell_lo, cell_hi) are made on this rank.  Cells
    // are independent sweep domains (a pair is reported in its home cell only), so ranks that
    // own disjoint cell ranges emit disjoint pair lists whose rank-order concatenation is the
    // single-GPU list -- no halo, no exchange.
    int cell_lo = 0, cell_hi = 0x7fffffff;
    // major-axis quantisation of the 32-bit sort key (see "32-bit sweep key" below)
    double x0 = 0, inv_hx = 0;
    int x_bits = 29;
};

__host__ __device__ inline int cell_index(double v, double v0, double inv_h, int s)
{
    // monotone non-decreasing in v (subtraction, multiplication by a non-negative constant,
    // truncation and clamping all are); NaN-free inputs assumed
    if (s <= 1)
        return 0;
    const double f = (v - v0) * inv_h;
    int i = f <= 0.0 ? 0 : (f >= (double)(s - 1) ? s - 1 : (int)f);
    return i;
}

// 32-bit sweep key of a record:  [ cell | q(x) : x_bits | fy | fz | type ]
//   q(x)  = monotone quantisation of the major-axis coordinate (floor of an affine map,
//           clamped), so  xmin_j <= xmax_i  =>  q(xmin_j) <= q(xmax_i): the window test on
//           keys is a conservative superset of the reference's  min_j <= max_i;
//   fy/fz = the box STARTS in this cell's row / column.  A record of cell (cy, cz) always has
//           cy0 <= cy, and cell_index() is monotone, so the home-cell rule
//           cell(max(ymin_a, ymin_b)) == cy  is exactly  fy_a | fy_b  (same for z): the
//           duplicate suppression costs two bit operations in the prefilter;
//   type  = 1 for list A (vertices) of a two-list sweep (collision.cuh:27-35).
constexpr int kKeyFlagBits = 3;
constexpr uint32_t kKeyFlagType = 1u, kKeyFlagZ = 2u, kKeyFlagY = 4u;

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t quantize_x(double x, const GridParams& g)
{
    const double f = fmax((x - g.x0) * g.inv_hx, 0.0);
    const uint32_t qmax = g.x_bits >= 32 ? 0xffffffffu : ((1u << g.x_bits) - 1u);
    const uint32_t q = __double2uint_rd(f); // saturates
    return q < qmax ? q : qmax;
}
#endif

// Prefilter view of the SORTED records: key (above), reach = the same cell with q(xmax) and all
// flag bits set, and the f32 conservative yz = (ymin dn, ymax up, zmin dn, zmax up).  For j
// after i in sorted order,  key[j] <= reach[i]  <=>  same cell and q(xmin_j) <= q(xmax_i).
struct PrefilterArrays {
    uint32_t* key = nullptr;
    uint32_t* reach = nullptr;
    float4* yz = nullptr;
};

// One sorted list ready for sweeping (n = number of RECORDS, >= number of boxes).
struct SortedList {
    int n = 0;
    bool two_lists = false;
    bool cell_sharded = false; // holds only this rank's cell range (GridParams::cell_lo/hi)
    GridParams grid;
    BoxArrays box;      // sorted, exact
    PrefilterArrays pf; // sorted prefilter view
};

constexpr int kNumStats = 14;
#ifdef __CUDACC__
// Box statistics of a list (grid choice): r[0..7] = {min ymin, max ymax, min zmin, max zmax,
// sum (ymax-ymin), sum (zmax-zmin), min xmin, max xmax, sum c_x, sum c_y, sum c_z, sum c_x^2,
// sum c_y^2, sum c_z^2} with c = box centre (the variance that picks the next sweep axis,
// sort_and_sweep.cpp:176-195; axes in the record's rotated order).  Fixed reduction trees
// everywhere, so the sums -- and with them the chosen grid -- are deterministic.
__device__ __forceinline__ void stats_identity(double r[kNumStats])
{
    r[0] = 1.7976931348623157e308, r[1] = -1.7976931348623157e308;
    r[2] = 1.7976931348623157e308, r[3] = -1.7976931348623157e308;
    r[4] = 0.0, r[5] = 0.0;
    r[6] = 1.7976931348623157e308, r[7] = -1.7976931348623157e308;
#pragma unroll
    for (int k = 8; k < kNumStats; k++)
        r[k] = 0.0;
}
__device__ __forceinline__ void stats_merge(double r[kNumStats], const double o[kNumStats])
{
    r[0] = fmin(r[0], o[0]);
    r[1] = fmax(r[1], o[1]);
    r[2] = fmin(r[2], o[2]);
    r[3] = fmax(r[3], o[3]);
    r[4] += o[4];
    r[5] += o[5];
    r[6] = fmin(r[6], o[6]);
    r[7] = fmax(r[7], o[7]);
#pragma unroll
    for (int k = 8; k < kNumStats; k++)
        r[k] += o[k];
}
__device__ __forceinline__ void stats_of_box(
    double r[kNumStats], const double lo[3], const double hi[3])
{
    r[0] = lo[1], r[1] = hi[1], r[2] = lo[2], r[3] = hi[2];
    r[4] = hi[1] - lo[1], r[5] = hi[2] - lo[2], r[6] = lo[0], r[7] = hi[0];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double c = (lo[k] + hi[k]) / 2; // sort_and_sweep.cpp:180
        r[8 + k] = c;
        r[11 + k] = c * c;
    }
}
// block reduction (blockDim.x = 32 * warps <= 1024); the result is valid on thread 0.
// red: shared scratch of 32 * kNumStats doubles.
__device__ __forceinline__ void stats_block_reduce(double r[kNumStats], double* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double v[kNumStats];
#pragma unroll
        for (int k = 0; k < kNumStats; k++)
            v[k] = __shfl_xor_sync(0xffffffffu, r[k], o);
        stats_merge(r, v);
    }
    __syncthreads(); // red may still be read by a previous reduction
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < kNumStats; k++)
            red[warp * kNumStats + k] = r[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < warps; w++)
            stats_merge(r, red + w * kNumStats);
}

// read-only 32-byte load (there is no __ldg overload for double4)
__device__ __forceinline__ double4 ldg_d4(const double4* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}
#endif

// ---- narrow-phase parameters -----------------------------------------------------------
struct NarrowParams {
    double ms;
    double tol;     // co-domain tolerance
    int max_iter;   // < 0: unlimited
    int allow_zero_toi;
    int use_ms;     // ms > 0 (selects the error filter, root_finder.cu:95-122)
    int flags;      // debug knobs (SCCD_NP_FLAGS env), see narrow.cu
    int max_depth;  // levels a walk may track before handing on (<= 128)
    int cap_drops;  // max_iter reached: 0 = accept the box at t_lo (conservative), 1 = drop it
    int want_tlb;   // the cull computes every survivor's toi lower bound (else: 0 for all)
    int queue_ctas; // CTAs of the work-queue launch (0: as many as fit)
    int tail_lanes; // lane-per-tree rounds: a warp with this few busy lanes and an empty pool
                    // hands its trees on to the next round (0: never)
    int root_check; // the cull also runs the solver's first box check on its survivors
    int cull_float; // double build: float pre-test + per-CTA compaction in front of the double test
    int solver;     // 0: by list length -- lane per tree in rounds (long), persistent work queue
                    // (short); 1: rounds for short lists too; 4, 8: lanes per tree (group solver)
    // Multi-GPU (sccd_ccd_sharded): the earliest-toi words of the OTHER ranks, mapped into this
    // rank's address space over NVLink (CUDA IPC, shard.cu).  A lane that lowers the bound also
    // lowers it on every peer with a system-scope atomicMin, so all ranks prune with the global
    // earliest toi while they work -- the collective is fused into the solver kernels; the
    // all-reduce at the end of the step only closes the race of the last updates.
    int n_peers;
    double* peer_toi[15];
};

// A pending sub-box of a query, handed from one round of the narrow phase to the next
// (56-byte payload, the size of the reference's CCDDomain).
struct __align__(16) WorkItem {
    double lo[3];
    double w[3];
    unsigned long long pad0;
    uint32_t query;
    uint32_t pad1;
};
static_assert(sizeof(WorkItem) == 64, "WorkItem");

// Rounds of one narrow-phase batch (see narrow.cu); the last one has no check budget.
constexpr int kNarrowRounds = 5;

// Every hot word sits in its own 128-byte line.
// (the running earliest toi is a separate device word so that concurrent batches share it)
struct alignas(128) NarrowCounters {
    alignas(128) unsigned long long next[kNarrowRounds];      // next unclaimed work index
    alignas(128) unsigned long long n_items[kNarrowRounds + 1]; // [r] = items round r reads
    // [r] != 0: the list round r reads was closed at item ~closed[r] (narrow.cu reserve_items)
    alignas(128) unsigned long long closed[kNarrowRounds + 1];
    alignas(128) unsigned long long next_scout;       // claim counter of round 0's scout launch
    // persistent work queues of a short list (narrow_coop_kernel<QUEUE>; [0] scout launch, [1]
    // bulk launch): tickets taken / slots pushed of the item list, "closed at" marker, and the
    // items being worked on or queued
    struct alignas(128) Queue {
        alignas(128) unsigned long long head;
        alignas(128) unsigned long long tail;
        unsigned long long closed;
        alignas(128) unsigned long long outstanding;
    } queue[2];
    alignas(128) int overflow;                        // an item list was full (work kept local)
    int bad_input;                                    // a pair id is no element of the mesh
    int round0_ran; // which launch found round 0's work: 1 lane-per-tree (long list), 2 work queue
    unsigned long long box_checks;
    unsigned long long round_checks[kNarrowRounds]; // box checks per round (load-balance report)
    unsigned long long donated;
    unsigned long long capped;
    unsigned long long started;  // trees round 0 started (survivors - started were skipped)
    // survivors per bucket of their toi lower bound (1/256 of the step each), made by the cull:
    // the digit histogram of the survivor sort, and what ordering_on() decides on (narrow.cu)
    alignas(128) uint32_t tlb_hist[256];
};

// ---- kernel launchers (defined in the .cu files) -----------------------------------
struct LaunchCounter {
    int64_t n = 0;
};

// f32: boxes of the reference's float build (float values, stored widened); radius_up is then
// nextafterf((float)r) computed by the caller
void launch_mesh_boxes(
    const double* V0, const double* V1, int nV, double radius_up, bool f32, VertexRec* vtab,
    double* vbox /* 6*nV: min xyz, max xyz */, const int32_t* E, int nE, const int32_t* F,
    int nF, BoxArrays e_unsorted, BoxArrays vf_unsorted, int axis_e, int axis_vf,
    int* bad /* device flag: set when an E / F entry is not a vertex index */, cudaStream_t s,
    LaunchCounter& lc);

// the reference's box builders by name, on caller-made AoS arrays (aabb.cuh:150-188)
void launch_vertex_aabbs(
    const double* V0, const double* V1, int nV, double radius_up, bool f32, sccd_aabb* out,
    cudaStream_t s, LaunchCounter& lc);
void launch_element_aabbs(
    const sccd_aabb* vb, int nV, const int32_t* idx, int n, int k, sccd_aabb* out, int* bad,
    cudaStream_t s, LaunchCounter& lc);

// measured FP64 pipe rate (thread-level DFMA / s): the narrow phase's compute roofline
double measure_dfma_per_second(int num_sms, cudaStream_t s);

// ---- grid (csrc/grid.cu)
// Statistics of a list (stats_identity() layout) over every stride-th box: they only steer
// the grid and the key quantisation, both of which clamp, so a sample is as good as the whole.
constexpr int kStatsBlocks = 296;
void launch_box_stats(
    const BoxArrays& unsorted, int n, int stride, double* partials /* kStatsBlocks*8 */,
    double* stats,
    cudaStream_t s, LaunchCounter& lc,
    double sum_scale = 0.0 /* factor of the two extent sums; <= 0: stride */);
// copies[i] = number of cells box i touches inside [g.cell_lo, g.cell_hi), or inside the
// range d_range[0..1] held on the device when d_range is not null
void launch_expand_count(
    const BoxArrays& unsorted, int n, GridParams g, const unsigned long long* d_range,
    uint32_t* copies, cudaStream_t s, LaunchCounter& lc);
// Multi-GPU: histogram of the records of every stride-th box per cell and `world` contiguous
// cell ranges of ~equal estimated sweep work.  out (device, 2 * world + 2 words): [0..world]
// first cell of each rank, [world+1] sampled records in total, [world+2+r] of rank r.
// hist: g.sy * g.sz words.
void launch_cell_splits(
    const BoxArrays& unsorted, int n, int stride, GridParams g, int world, uint32_t* hist,
    unsigned long long* out, cudaStream_t s, LaunchCounter& lc);
// one 64-bit record (key << 32 | idx_base + box index) per touched cell, at offsets[i] ...
// sort_hist (launch_sort_prepare): the digit histograms of the sort on key_bits that follows
void launch_expand_fill(
    const BoxArrays& unsorted, int n, GridParams g, const unsigned long long* offsets,
    unsigned long long* rec, uint32_t* sort_hist, int key_bits, cudaStream_t s, LaunchCounter& lc);
uint32_t* launch_sort_prepare(void* temp, size_t temp_bytes, long long m, cudaStream_t s);

// ---- multi-GPU build (shard.cu): slice / sample boxes, records, partition, rebuild
struct MeshView;
void launch_list_boxes(
    const MeshView& m, int list, long long first, int stride, int count, BoxArrays out, int axis,
    int* bad, cudaStream_t s, LaunchCounter& lc);
void launch_expand_fill_records(
    const BoxArrays& unsorted, int n, GridParams g, const unsigned long long* offsets,
    uint32_t idx_base, unsigned long long* rec, cudaStream_t s, LaunchCounter& lc);
size_t partition_temp_bytes(long long m);
// stable partition of the records by the rank that owns their cell; counts: `world` device words
void launch_partition_by_dest(
    long long m, const unsigned long long* rec_in, unsigned long long* rec_out, int cell_shift,
    const unsigned long long* h_first_cell /* world + 1, host */, int world,
    unsigned long long* counts, void* temp, size_t temp_bytes, cudaStream_t s, LaunchCounter& lc);
size_t sort_records_temp_bytes(long long m);
void launch_sort_records_and_rebuild(
    int m, int key_bits, unsigned long long* rec, unsigned long long* rec_tmp, void* temp,
    size_t temp_bytes, const MeshView& mesh, int list, int axis, SortedList out, cudaStream_t s,
    LaunchCounter& lc, cudaEvent_t gather_begin = nullptr, cudaEvent_t gather_end = nullptr);

size_t sort_temp_bytes(long long m);
// sorts m 64-bit records (key << 32 | box index) on key bits [kKeyFlagBits, kKeyFlagBits +
// key_bits) and gathers the sorted views; rec / rec_tmp are ping-pong buffers
void launch_sort_and_gather(
    int m, int key_bits, unsigned long long* rec, unsigned long long* rec_tmp, void* temp,
    size_t temp_bytes, BoxArrays unsorted, SortedList out, cudaStream_t s, LaunchCounter& lc,
    cudaEvent_t gather_begin = nullptr, cudaEvent_t gather_end = nullptr,
    bool hist_ready = false /* launch_sort_prepare + launch_expand_fill made the histograms */);

// window[i] = number of candidates after owner i whose f32 xmin <= owner's f32 xmax
void launch_sweep_windows(
    const SortedList& L, uint32_t* window, cudaStream_t s, LaunchCounter& lc);
// Staging area between the count and the place pass (see sweep.cu): sizes for `owners` owners.
size_t sweep_stage_tiles(int owners);
size_t sweep_stage_pair_bytes(int owners);
size_t sweep_stage_tag_bytes(int owners);
void launch_sweep_count(
    const SortedList& L, int owner_lo, int owner_hi, uint32_t* counts,
    unsigned long long* n_candidates, void* stage_pairs, void* stage_tags, uint32_t* stage_count,
    cudaStream_t s, LaunchCounter& lc,
    bool tma_staged = false /* prefilter stream through shared memory (cp.async.bulk) */);
// counts / offsets are indexed relative to the first owner of the count pass (shard_lo);
// the fill of owner range [owner_lo, owner_hi) writes pairs[offsets[i] - offsets[owner_lo]..).
void launch_sweep_fill(
    const SortedList& L, int shard_lo, int owner_lo, int owner_hi,
    const unsigned long long* offsets, sccd_pair* pairs, void* stage_pairs, void* stage_tags,
    uint32_t* stage_count, cudaStream_t s, LaunchCounter& lc);
size_t scan_temp_bytes(int n);
// offsets[0..n] = exclusive prefix sum of counts[0..n) (offsets[n] = total)
void launch_scan_u32_to_u64(
    const uint32_t* counts, unsigned long long* offsets, int n, void* temp,
    size_t temp_bytes, cudaStream_t s, LaunchCounter& lc);
// first index e in (lo, hi] with offsets[e] - offsets[lo] > budget, minus one (>= lo+1 if
// the first owner alone fits); result written to *d_out
void launch_find_chunk_end(
    const unsigned long long* offsets, int lo, int hi, unsigned long long budget,
    int* d_out, cudaStream_t s, LaunchCounter& lc);

// once per context (sccd_create): opts the solver kernels in to their dynamic shared memory on
// the current device (cudaFuncSetAttribute is per device and idempotent)
void narrow_init_device();

struct NarrowInput {
    // mesh mode
    const VertexRec* vtab = nullptr;
    const int32_t* E = nullptr;
    const int32_t* F = nullptr;
    int nV = 0, nE = 0, nF = 0;
    const sccd_pair* pairs = nullptr;
    // direct mode (n x 24 doubles)
    const double* queries = nullptr;
    long long n = 0;
};
// Enqueues every round of one batch.  items[0/1]: two lists of item_cap WorkItems each.
// After the batch, counters->n_items[kNarrowRounds] != 0 means the last round had to hand
// work on (path deeper than the lane state can track): run launch_narrow_extra_round until
// it is zero.
// f32: the reference's float build (SCALABLE_CCD_USE_DOUBLE off) -- float arithmetic on the
// same (float-valued) double buffers; cull with the float filters, both solver kernels in float
void launch_narrow_phase(
    bool is_vf, bool f32, const NarrowInput& in, const NarrowParams& p, NarrowCounters* counters,
    double* g_toi, WorkItem* items0, WorkItem* items1, unsigned long long item_cap, double* toi_per_query,
    unsigned int* checks_per_query,
    unsigned long long* survivors /* 2 * in.n records, or null: no cull */,
    float* tlb /* in.n: lower bound of every surviving query's toi */, void* sort_temp,
    size_t sort_temp_bytes, int num_sms, cudaStream_t s, LaunchCounter& lc,
    const cudaEvent_t* tev = nullptr /* 2 * (1 + kNarrowRounds) optional timing events */,
    cudaEvent_t solver_waits_for = nullptr /* the rounds (not the cull) start after this event */,
    // Which kernels solve a batch depends on the length of its survivor list, known on the device
    // only; launching all of them costs ~6 kernels per batch that return at once.  mode_hint says
    // what the previous batch of this kind needed (1: short list -> work queue only; 0: long list
    // -> scout + rounds; -1: launch everything).  counters->round0_ran tells afterwards whether
    // the guess was right; if not, call again with solver_only = true and mode_hint = -1.
    int mode_hint = -1, bool solver_only = false);
size_t sort_survivors_temp_bytes(long long n_max);
// hist: the 256-bucket histogram of bits 32..39 the producer made (device).  The sort does
// nothing when the buckets do not discriminate (2 * hist[0] >= *d_n; narrow.cu: ordering_on).
void launch_sort_survivors(
    const unsigned long long* rec, unsigned long long* rec_out, const unsigned long long* d_n,
    const uint32_t* hist, long long n_max, void* temp, size_t temp_bytes, cudaStream_t s,
    LaunchCounter& lc);
// moves n_items[kNarrowRounds] to n_items[kNarrowRounds - 1] and reruns the last round
void launch_narrow_extra_round(
    bool is_vf, bool f32, const NarrowInput& in, const NarrowParams& p, NarrowCounters* counters,
    double* g_toi, WorkItem* items0, WorkItem* items1, unsigned long long item_cap, int extra_index,
    double* toi_per_query, unsigned int* checks_per_query, int num_sms, cudaStream_t s,
    LaunchCounter& lc);
void launch_fill_f64(double* p, long long n, double v, cudaStream_t s, LaunchCounter& lc);
// compacts (pair, toi) of queries with toi < 1 ; *d_count receives the number
void launch_compact_collisions(
    const sccd_pair* pairs, const double* toi_q, long long n, sccd_pair* out_ids,
    double* out_toi, unsigned long long* d_count, cudaStream_t s, LaunchCounter& lc);

} // namespace sccd
