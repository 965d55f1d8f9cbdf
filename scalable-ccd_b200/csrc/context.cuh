// The context behind the C ABI (include/sccd.h) and the host-side helpers shared by the
// translation units that drive it: api.cu (single-device pipeline, C entry points) and shard.cu
// (multi-GPU: NCCL communicator, sliced box build, record exchange).
#pragma once

#include "common.cuh"

#include <algorithm>
#include <string>
#include <vector>

using namespace sccd;

struct sccd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    std::string error;

    size_t memory_limit = 0;
    size_t mem_free = 0, mem_total = 0;   // last cudaMemGetInfo() answer ...
    unsigned long long mem_epoch = 0;     // ... and the allocation epoch (+1) it was taken at
    int64_t max_pairs_per_chunk = 0;
    int64_t queue_cap = 0;
    int rank = 0, world = 1;
    // SCCD_F32: the reference's float build (scalar.hpp:16-18) -- inputs rounded to float, boxes
    // by nextafterf, narrow phase in float arithmetic; every buffer stays double (float values
    // are exact in double and compare the same)
    bool f32 = false;

    // mesh
    int nV = 0, nE = 0, nF = 0;
    bool have_mesh = false, have_boxes = false;
    const double *dV0 = nullptr, *dV1 = nullptr;
    const int32_t *dE = nullptr, *dF = nullptr;
    DevBuf bV0, bV1, bE, bF;
    DevBuf b_vtab, b_vbox;
    DevBuf b_io[3]; // staging of the stand-alone box builders (sccd_build_vertex/element_boxes)

    // box lists: [0] = vertex+face (two lists), [1] = edges
    struct ListBufs {
        DevBuf ux, uyz, uid;     // unsorted exact records (one per box)
        DevBuf copies, offs;     // cells touched per box, and their exclusive scan
        DevBuf scan_temp;        // scratch of that scan (per list: the two lists count together)
        DevBuf keys, keys_tmp;   // 64-bit (key, box index) records, one per (box, cell)
        DevBuf sx, syz, sid;     // sorted exact records
        DevBuf pkey, preach, pyz; // sorted prefilter view
        DevBuf sort_temp;         // radix sort scratch
        SortedList sorted;
        BoxArrays unsorted;
        int n_boxes = 0;
        // frame-to-frame: the statistics the grid was chosen from (valid for a list of stats_n
        // boxes swept along stats_axis in scalar mode stats_f32)
        bool stats_valid = false;
        int stats_n = -1, stats_axis = -1;
        bool stats_f32 = false;
        int axis = 0;      // axis the records of this list are rotated to / swept along
        int next_axis = 0; // variance argmax of the last build (sort_and_sweep.cpp:176-195)
        int built_rank = 0, built_world = 1; // sharding the sorted records were made for
        // sort_list_begin -> sort_list_finish hand-over
        GridParams g_try;
        bool try_sharded = false;
        int try_stride = 1, attempt = 0;
    } lists[3]; // [2] = caller-made boxes (sccd_set_boxes)
    bool have_custom = false;
    DevBuf b_scan_temp, b_stats, b_hist, b_splits;
    // The edge list is sorted on a second stream, under the vertex-face sweep and narrow phase:
    // the sort of a 1 M-box list is a dozen latency-bound launches that leave the GPU mostly
    // idle.  ev_counts: both lists counted (main stream); ev_sorted1: edge list sorted.
    cudaStream_t sort_stream = nullptr;
    cudaStream_t sort_stream_own = nullptr; // the stream created for it (SCCD_OPT_PROFILE 2 aliases
                                            // sort_stream to the caller's stream)
    cudaEvent_t ev_counts = nullptr, ev_sorted1 = nullptr, ev_vf_done = nullptr;
    cudaEvent_t ev_boxes = nullptr, ev_stats = nullptr; // frame-to-frame statistics (build_boxes)
    cudaEvent_t ev_cnt1 = nullptr; // the edge list's record count has arrived (sort stream)
    bool stats_in_flight = false;
    bool sort1_pending = false;
    // pinned: per list, box statistics + record count + multi-GPU cell splits
    struct ListHost {
        double stats[kNumStats];
        double stats_next[kNumStats]; // statistics of the CURRENT boxes, on their way for the next build
        unsigned long long m;
        unsigned long long splits[2 * 16 + 2];
    };
    ListHost* h_lists = nullptr; // [3]
    DevBuf b_flags;            // [0]: an E / F entry is not a vertex index (box kernels)
    int* h_flags = nullptr;    // pinned copy
    int grid_max_cells = -1;   // < 0: choose automatically; 1 forces the plain 1-axis sweep
    // tuning knobs (env SCCD_GRID_SCALE / SCCD_GRID_REPL): cell edge in mean box extents, and
    // the replication (records per box) above which the grid is coarsened
    double grid_scale = 3.0, grid_repl = 2.5;
    // sccd_set_option (initial values from the SCCD_* environment variables, read ONCE in
    // sccd_create; nothing on the hot path calls getenv)
    struct Options {
        int np_cull = 1;        // separating-axis cull in front of the solver
        int np_flags = 0;       // narrow-phase scheduling knobs (narrow.cu)
        int np_flags_ee = -1;   // the same for the edge-edge pass alone (< 0: follow np_flags)
        int np_depth = 128;     // levels a walk tracks before handing on
        int cap_drops = 0;      // max_iter reached: 0 accept at t_lo, 1 drop (reference)
        int key_steps = 3;      // log2 of the x quantisation steps per record of a cell
        int sweep_axis = 0;     // 0/1/2, or -1: variance argmax of the previous build
        int profile = 0;        // time every solver round (sccd_stats.ms_k_round)
        int np_solver = 0;      // 0: lane / warp per tree by list length; 4, 8: lanes per tree
        int concurrent_passes = 0; // edge-edge solver does not wait for the vertex-face one
        int sweep_staged = 0;   // sweep count pass reads its window from TMA-staged shared memory
        int reuse_grid = 1;     // frame-to-frame: grid from the previous build's statistics
        int queue_ctas[2] = { 0, 0 }; // CTAs of the work-queue launch per list (0: all that fit)
        int np_tail_lanes = 16;       // see NarrowParams::tail_lanes (config 3: 5.04 -> 4.92-4.97 ms)
    } opt;
    int next_axis = 0;          // argmax of the box-centre variance of the last build

    // Run state (broad-phase cursor, pair / staging buffers, narrow-phase lists and counters)
    // and the stream its work is enqueued on.  Broad and narrow phase are split in an enqueue
    // and a finish half, so several runs on several streams can be in flight; the pipeline
    // uses one (see run_pipeline).
    struct Run {
        cudaStream_t stream = nullptr;
        int bp_kind = -1;
        int shard_lo = 0, shard_hi = 0, bp_cursor = 0;
        unsigned long long bp_total = 0, bp_emitted = 0;
        DevBuf b_counts, b_offsets, b_scan, b_pairs, b_small;
        DevBuf b_stage_pairs, b_stage_tags, b_stage_count; // count pass -> place pass
        int* h_small = nullptr; // pinned scratch for tiny D2H results
        DevBuf b_counters, b_items[2], b_toi_q, b_checks_q, b_queries, b_surv, b_tlb, b_surv_sort;
        NarrowCounters* h_counters = nullptr; // pinned
        unsigned long long item_cap = 0; // capacity of each of the two hand-on lists
        long long checks_n = 0;          // queries of the last batch that counted its checks
        // narrow_enqueue -> narrow_finish hand-over
        struct Pending {
            bool active = false;
            int kind = 0;
            NarrowInput in;
            NarrowParams P;
            double* d_tq = nullptr;
            unsigned int* checks = nullptr;
            bool culling = false;
            int hint = -1; // the guess the launches were chosen by (launch_narrow_phase)
            unsigned long long* survivors = nullptr;
            float* tlb = nullptr;
        } pending;
    } runs[2]; // [1]: the edge list's broad phase + narrow phase on the sort stream (pipeline)
    Run* cur = &runs[0];
    // what the previous batch of each kind needed: 1 work queue (short survivor list), 0 rounds
    // (long list), -1 unknown.  Frame to frame the list length hardly changes.
    int np_hint[2] = { -1, -1 };
    long long np_last_survivors[2] = { 0, 0 }; // cull survivors of the previous batch of each kind
    DevBuf b_gtoi; // earliest toi shared by the two lists of a pipeline call
    double* h_gtoi = nullptr; // pinned

    // ---- multi-GPU (shard.cu): communicator + buffers of the sliced build
    void* nccl_comm = nullptr;          // ncclComm_t (NCCL is loaded at run time)
    int comm_world = 0;                 // 0: no communicator attached (sccd_comm_create)
    struct SliceList {
        DevBuf samp_x, samp_yz, samp_id;      // every stride-th box of the WHOLE list (statistics)
        DevBuf rec, rec_tmp;                  // records of this rank's slice, then grouped by owner
        DevBuf recv, recv_sorted;             // records of this rank's cell range, from every rank
        DevBuf part_temp, sort_temp;
        long long lo = 0, hi = 0;             // slice [lo, hi) of the list's boxes
        int stride = 1, ns = 0;               // sample stride and size
    } slice[2];
    // earliest-toi words of the other ranks (CUDA IPC mappings over NVLink): the solver kernels
    // publish every improvement of the bound to all ranks (NarrowParams::peer_toi)
    int n_peers = 0;
    double* peer_toi[15] = {};
    bool peer_is_ipc[15] = {};
    bool share_toi = false;             // set for the duration of a sharded pipeline call
    // Frame-to-frame: when the toi lower bounds of a pass let it skip less than half of its
    // survivors -- a pile of rigid bodies whose boxes overlap at t = 0 -- the next batches of
    // that pass do not compute them (a third of the cull's arithmetic); every 16th batch probes
    // again.
    int tlb_pause[2] = { 0, 0 };
    int tlb_pause_len[2] = { 7, 7 }; // doubles (+1) every time the bounds turn out useless again
    bool sliced = false;                // the mesh lists hold slices / received records
    DevBuf b_xcnt, b_xsplits;           // all ranks' send counts; both lists' cell splits
    unsigned long long* h_xcnt = nullptr; // pinned
    DevBuf b_mesh_pad[4];               // sccd_ccd_sharded_host: padded V0 / V1 / E / F
    cudaEvent_t ev_xa = nullptr, ev_xb = nullptr; // around the record exchange

    // collisions of the last sccd_ccd_collisions (fetched by sccd_get_collisions)
    std::vector<sccd_pair> coll_ids;
    std::vector<double> coll_toi;
    int64_t coll_n[2] = { 0, 0 };

    sccd_stats stats {};
    LaunchCounter lc;
    cudaEvent_t ev[24] {};
    bool gather_timed = false;
    // pooled event pairs timing single kernels; resolved into stats at the end of a call
    struct KTimer {
        cudaEvent_t a = nullptr, b = nullptr;
        cudaStream_t st = nullptr;
        float* dst = nullptr;
    };
    std::vector<KTimer> ktimers;
    size_t kt_used = 0;

    ~sccd_ctx()
    {
        for (auto& r : runs) {
            if (r.h_small)
                cudaFreeHost(r.h_small);
            if (r.h_counters)
                cudaFreeHost(r.h_counters);
        }
        if (h_gtoi)
            cudaFreeHost(h_gtoi);
        if (sort_stream_own)
            sort_stream = sort_stream_own;
        if (sort_stream)
            cudaStreamDestroy(sort_stream);
        if (ev_counts)
            cudaEventDestroy(ev_counts);
        if (ev_sorted1)
            cudaEventDestroy(ev_sorted1);
        if (ev_vf_done)
            cudaEventDestroy(ev_vf_done);
        if (ev_cnt1)
            cudaEventDestroy(ev_cnt1);
        if (ev_boxes)
            cudaEventDestroy(ev_boxes);
        if (ev_stats)
            cudaEventDestroy(ev_stats);
        if (h_lists)
            cudaFreeHost(h_lists);
        if (h_flags)
            cudaFreeHost(h_flags);
        if (h_xcnt)
            cudaFreeHost(h_xcnt);
        if (ev_xa)
            cudaEventDestroy(ev_xa);
        if (ev_xb)
            cudaEventDestroy(ev_xb);
        for (auto& e : ev)
            if (e)
                cudaEventDestroy(e);
        for (auto& k : ktimers) {
            cudaEventDestroy(k.a);
            cudaEventDestroy(k.b);
        }
    }
};

namespace sccd {
namespace host {

enum { EV_T0, EV_BUILD, EV_SORT, EV_SW0A, EV_SW0B, EV_NP0A, EV_NP0B, EV_SW1A, EV_SW1B,
       EV_NP1A, EV_NP1B, EV_T1, EV_TMPA, EV_TMPB, EV_GA0, EV_GB0, EV_GA1, EV_GB1, EV_SB0, EV_SB1,
       EV_COUNT };

inline void use_device(sccd_ctx* c) { SCCD_CUDA(cudaSetDevice(c->device)); }

template <typename F> int guarded(sccd_ctx* c, F&& f)
{
    if (!c)
        return SCCD_ERR_ARG;
    try {
        use_device(c);
        c->cur = &c->runs[0]; // step-wise entry points always work on the context's own stream
        return f();
    } catch (const CudaError& e) {
        c->error = e.what();
        (void)cudaGetLastError();
        return SCCD_ERR_CUDA;
    } catch (const std::invalid_argument& e) {
        c->error = e.what();
        return SCCD_ERR_ARG;
    } catch (const std::logic_error& e) {
        c->error = e.what();
        return SCCD_ERR_STATE;
    } catch (const std::bad_alloc&) {
        c->error = "out of host memory";
        return SCCD_ERR_MEMORY;
    } catch (const std::exception& e) {
        c->error = e.what();
        return SCCD_ERR_MEMORY;
    }
}

// The box statistics (grid choice, key quantisation, next sweep axis) and the multi-GPU cell
// histogram look at every stride-th box of a list: ~260 K .. 1 M samples steer them as well as
// the whole list does (both clamp), and the pass stays off the critical path of a 50 M-box step.
inline int stats_stride(long long n) { return (int)std::min<long long>(64, std::max<long long>(1, n >> 18)); }

// every host <-> device synchronisation of a pipeline call goes through here (sccd_stats)
inline void host_sync(sccd_ctx* c, cudaStream_t s)
{
    SCCD_CUDA(cudaStreamSynchronize(s));
    c->stats.n_host_syncs++;
}

// ---- api.cu
void record(sccd_ctx* c, int which);
void rrecord(sccd_ctx* c, int which);
float elapsed(sccd_ctx* c, int a, int b);
size_t kt_alloc(sccd_ctx* c, float* dst);
size_t kt_begin(sccd_ctx* c, float* dst, cudaStream_t st = nullptr); // null: the current run's

void kt_end(sccd_ctx* c, size_t id);
sccd_ctx::ListHost& list_host(sccd_ctx* c, int which);
GridParams choose_grid(const double* st, int n, int max_cells, double scale);
void regrid(GridParams& g, const double st[kNumStats]);
void key_layout(
    GridParams& g, unsigned long long m_total, const double* st, int key_steps, int& cell_bits);
void adopt_stats(sccd_ctx* c, int which, long long n_full);
bool stats_reusable(const sccd_ctx* c, int which, long long n_full);
void drain_stats(sccd_ctx* c);
void join_sort_stream(sccd_ctx* c, cudaStream_t st);
void build_boxes(sccd_ctx* c, double inflation_radius);
void small_scratch(sccd_ctx* c);
void upload_mesh(
    sccd_ctx* c, const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, bool on_device);

void prepare_list(sccd_ctx* c, int which, int n, bool two_lists);
void run_pipeline(
    sccd_ctx* c, double min_distance, int max_iter, double tol, bool allow_zero_toi, bool ipc,
    double* toi_out, bool want_collisions, std::vector<sccd_pair>* coll_ids,
    std::vector<double>* coll_toi, int64_t* n_coll, bool sharded = false);

// ---- shard.cu
// build_boxes() of a rank that holds a communicator: boxes of 1 / world of the elements,
// (key, index) records exchanged by owning cell range, exact records rebuilt by the receiver
void build_boxes_sliced(sccd_ctx* c, double inflation_radius);
// in-place all-reduce(min) of the earliest-toi word on `st` (no-op without a communicator)
void allreduce_min_toi(sccd_ctx* c, double* d_toi, cudaStream_t st);
void comm_destroy(sccd_ctx* c);

} // namespace host
} // namespace sccd
