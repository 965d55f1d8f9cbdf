// Multi-GPU: communicator, sliced box build and record exchange (include/sccd.h "multi-GPU").
//
// The reference has no working multi-GPU path: cuda/broad_phase/_multigpu/broad_phase.cu:69-116
// is dead code in which every device holds ALL boxes, sweeps a range of the sorted list and the
// host merges the per-device overlap vectors.  Here one rank = one GPU = one context:
//
//   replicated (every rank)      vertex table + vertex boxes (144 B / vertex), a 1-in-16 sample of
//                                the boxes -> identical statistics, cell grid and cell ranges on
//                                every rank without a collective
//   sliced (1 / world of it)     element boxes, (key, index) records per (box, cell), stable
//                                partition of the records by owning rank
//   exchanged (NVLink, NCCL)     the send-count matrix (all-gather, 17 words per rank) and the
//                                8-byte records (grouped send / recv): 8 B per record instead of the
//                                64 B exact box, which the receiver REBUILDS from the replicated
//                                vertex boxes while it gathers its sorted list
//   local                        radix sort, sweep, narrow phase (unchanged single-GPU kernels)
//   reduced                      the earliest TOI: one all-reduce(min) of 8 bytes
//
// Received records arrive in source-rank order and every source emits its slice in box order,
// so after the stable radix sort the list of a rank is exactly the single-GPU sorted list
// restricted to its cell range: the ranks' pair lists are disjoint and concatenate to the
// single-GPU list.
#include "context.cuh"

#include "boxmake.cuh"

#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h> // types and prototypes only: the library is loaded at run time

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <mutex>

namespace sccd {
namespace host {

namespace {

// ---- NCCL, resolved at run time so that libsccd_b200.so has no link-time dependency on it ----
struct NcclApi {
    void* lib = nullptr;
    std::string error;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { getenv("SCCD_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            if (!n || !*n)
                continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib)
                break;
        }
        if (!api.lib) {
            api.error = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "");
            return;
        }
        bool ok = true;
        auto sym = [&](const char* name) {
            void* p = dlsym(api.lib, name);
            if (!p) {
                ok = false;
                api.error = std::string("NCCL symbol missing: ") + name;
            }
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) {
            dlclose(api.lib);
            api.lib = nullptr;
        }
    });
    return api;
}

void nccl_check(ncclResult_t r, const char* what)
{
    if (r != ncclSuccess)
        throw CudaError(std::string(what) + ": " + nccl().GetErrorString(r));
}
#define SCCD_NCCL(expr) nccl_check((expr), #expr)

static_assert(sizeof(ncclUniqueId) == SCCD_UNIQUE_ID_BYTES, "ncclUniqueId");

constexpr int kMaxWorld = 16;
// all-gathered per rank: send counts of both lists, then the rank's "bad index" flag
constexpr int kXWords = 2 * kMaxWorld + 1;

inline ncclComm_t comm_of(sccd_ctx* c) { return (ncclComm_t)c->nccl_comm; }

} // namespace

// Maps the earliest-toi word of every other rank into this rank's address space (CUDA IPC over
// NVLink; plain peer access for ranks that are threads of this process).  Collective.
static void map_peer_toi(sccd_ctx* c, int rank, int world)
{
    struct Card {
        cudaIpcMemHandle_t handle;
        unsigned long long pid, ptr, device, pad;
    };
    static_assert(sizeof(Card) % 8 == 0, "Card");
    double* mine = (double*)c->b_gtoi.reserve(64);
    Card card {};
    SCCD_CUDA(cudaIpcGetMemHandle(&card.handle, mine));
    card.pid = (unsigned long long)getpid();
    card.ptr = (unsigned long long)(uintptr_t)mine;
    card.device = (unsigned long long)c->device;
    DevBuf stage;
    Card* d_cards = (Card*)stage.reserve(sizeof(Card) * world);
    std::vector<Card> cards(world);
    SCCD_CUDA(cudaMemcpyAsync(d_cards + rank, &card, sizeof(Card), cudaMemcpyHostToDevice, c->stream));
    SCCD_NCCL(nccl().AllGather(
        d_cards + rank, d_cards, sizeof(Card), ncclChar, (ncclComm_t)c->nccl_comm, c->stream));
    SCCD_CUDA(cudaMemcpyAsync(
        cards.data(), d_cards, sizeof(Card) * world, cudaMemcpyDeviceToHost, c->stream));
    SCCD_CUDA(cudaStreamSynchronize(c->stream));
    c->n_peers = 0;
    for (int p = 0; p < world; p++) {
        if (p == rank)
            continue;
        void* ptr = nullptr;
        bool ipc = false;
        if (cards[p].pid == card.pid) { // a thread of this process: peer access, same pointer
            if ((int)cards[p].device != c->device) {
                const cudaError_t e = cudaDeviceEnablePeerAccess((int)cards[p].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    (void)cudaGetLastError();
                    continue; // no peer access: this peer only gets the final all-reduce
                }
                (void)cudaGetLastError();
            }
            ptr = (void*)(uintptr_t)cards[p].ptr;
        } else {
            if (cudaIpcOpenMemHandle(&ptr, cards[p].handle, cudaIpcMemLazyEnablePeerAccess)
                != cudaSuccess) {
                (void)cudaGetLastError();
                continue;
            }
            ipc = true;
        }
        c->peer_is_ipc[c->n_peers] = ipc;
        c->peer_toi[c->n_peers++] = (double*)ptr;
    }
}

void comm_destroy(sccd_ctx* c)
{
    for (int p = 0; p < c->n_peers; p++)
        if (c->peer_is_ipc[p])
            cudaIpcCloseMemHandle(c->peer_toi[p]);
    c->n_peers = 0;
    if (c->nccl_comm) {
        nccl().CommDestroy(comm_of(c));
        c->nccl_comm = nullptr;
    }
    c->comm_world = 0;
}

void allreduce_min_toi(sccd_ctx* c, double* d_toi, cudaStream_t st)
{
    if (c->comm_world > 1)
        SCCD_NCCL(nccl().AllReduce(d_toi, d_toi, 1, ncclDouble, ncclMin, comm_of(c), st));
}

// Exchange plan of one list from the all-gathered send-count matrix: what this rank receives
// from every source, where it lands (sources in rank order), and where the part for every
// destination starts in this rank's dest-grouped send buffer.  Pure host arithmetic.
struct ExchangePlan {
    unsigned long long send_cnt[kMaxWorld], send_off[kMaxWorld];
    unsigned long long recv_cnt[kMaxWorld], recv_off[kMaxWorld];
    unsigned long long recv_total = 0, send_total = 0;
};
ExchangePlan exchange_plan(const unsigned long long* all /* world x kXWords */, int list, int rank, int world)
{
    ExchangePlan p;
    for (int d = 0; d < world; d++) {
        p.send_cnt[d] = all[(size_t)rank * kXWords + list * kMaxWorld + d];
        p.send_off[d] = p.send_total;
        p.send_total += p.send_cnt[d];
    }
    for (int s = 0; s < world; s++) {
        p.recv_cnt[s] = all[(size_t)s * kXWords + list * kMaxWorld + rank];
        p.recv_off[s] = p.recv_total;
        p.recv_total += p.recv_cnt[s];
    }
    return p;
}

void build_boxes_sliced(sccd_ctx* c, double inflation_radius)
{
    if (!c->have_mesh)
        throw std::logic_error("build_boxes: no mesh uploaded");
    if (c->comm_world < 1)
        throw std::logic_error("sccd_ccd_sharded: no communicator (sccd_comm_create)");
    join_sort_stream(c, c->stream);
    drain_stats(c);
    const int W = c->comm_world, rank = c->rank;
    const int nV = c->nV, nE = c->nE, nF = c->nF;
    const long long n_list[2] = { (long long)nV + nF, (long long)nE };
    if (n_list[0] >= (1ll << 27) || n_list[1] >= (1ll << 27))
        throw std::invalid_argument("build_boxes: more than 2^27 boxes in one list");
    cudaStream_t st = c->stream;
    c->b_vtab.reserve(sizeof(VertexRec) * (size_t)std::max(nV, 1));
    c->b_vbox.reserve(sizeof(double) * 6 * (size_t)std::max(nV, 1));
    const double radius_up = c->f32
        ? (double)std::nextafterf((float)inflation_radius, FLT_MAX)
        : std::nextafter(inflation_radius, DBL_MAX);
    for (int k = 0; k < 2; k++) {
        auto& L = c->lists[k];
        L.axis = c->opt.sweep_axis >= 0 ? c->opt.sweep_axis : L.next_axis;
    }
    int* d_bad = (int*)c->b_flags.reserve(64);
    if (!c->h_flags)
        SCCD_CUDA(cudaMallocHost((void**)&c->h_flags, 64));
    if (!c->h_xcnt)
        SCCD_CUDA(cudaMallocHost((void**)&c->h_xcnt, sizeof(unsigned long long) * kMaxWorld * kXWords));
    if (!c->ev_xa) {
        SCCD_CUDA(cudaEventCreate(&c->ev_xa));
        SCCD_CUDA(cudaEventCreate(&c->ev_xb));
    }
    SCCD_CUDA(cudaMemsetAsync(d_bad, 0, 4, st));

    // ---- 1. replicated: vertex table + vertex boxes of ALL vertices
    const size_t kt_boxes = kt_begin(c, &c->stats.ms_k_boxes);
    launch_mesh_boxes(
        c->dV0, c->dV1, nV, radius_up, c->f32, c->b_vtab.as<VertexRec>(), c->b_vbox.as<double>(),
        c->dE, 0, c->dF, 0, BoxArrays(), BoxArrays(), 0, 0, d_bad, st, c->lc);
    MeshView mv;
    mv.vbox = c->b_vbox.as<double>();
    mv.E = c->dE;
    mv.F = c->dF;
    mv.nV = nV, mv.nE = nE, mv.nF = nF;

    // ---- 2. replicated: every stride-th box of both lists -> statistics (same samples, same
    // reduction tree as the single-GPU build: identical grids)
    const size_t per = (size_t)(kStatsBlocks + 1) * kNumStats;
    double* stats_base = (double*)c->b_stats.reserve(3 * per * sizeof(double));
    BoxArrays samp[2];
    for (int k = 0; k < 2; k++) {
        auto& S = c->slice[k];
        auto& H = list_host(c, k);
        S.stride = stats_stride(n_list[k]);
        S.ns = (int)((n_list[k] + S.stride - 1) / S.stride);
        S.lo = n_list[k] * rank / W;
        S.hi = n_list[k] * (rank + 1) / W;
        const size_t ns = (size_t)std::max(S.ns, 1);
        samp[k].x = (double2*)S.samp_x.reserve(ns * sizeof(double2));
        samp[k].yz = (double4*)S.samp_yz.reserve(ns * sizeof(double4));
        samp[k].id = (int4*)S.samp_id.reserve(ns * sizeof(int4));
        if (n_list[k] <= 0)
            continue;
        launch_list_boxes(mv, k, 0, S.stride, S.ns, samp[k], c->lists[k].axis, d_bad, st, c->lc);
        double* base = stats_base + k * per;
        launch_box_stats(
            samp[k], S.ns, 1, base, base + kStatsBlocks * kNumStats, st, c->lc, (double)S.stride);
        SCCD_CUDA(cudaMemcpyAsync(
            H.stats_next, base + kStatsBlocks * kNumStats, kNumStats * sizeof(double),
            cudaMemcpyDeviceToHost, st));
    }
    kt_end(c, kt_boxes);
    record(c, EV_BUILD);
    // frame-to-frame (as build_boxes): the previous build's statistics choose the grid at once;
    // every rank holds the same history, hence takes the same decision
    const bool reuse = stats_reusable(c, 0, n_list[0]) && stats_reusable(c, 1, n_list[1]);
    if (!reuse) {
        host_sync(c, st); // sync 1: statistics
        adopt_stats(c, 0, n_list[0]);
        adopt_stats(c, 1, n_list[1]);
    }

    // next sweep axis from the sample variance (as build_boxes)
    for (int k = 0; k < 2; k++) {
        auto& L = c->lists[k];
        const double* s = list_host(c, k).stats;
        const double ns = (double)c->slice[k].ns;
        double var[3] = { 0, 0, 0 };
        for (int a = 0; a < 3; a++)
            var[(L.axis + a) % 3] = ns > 0 ? s[11 + a] - s[8 + a] * s[8 + a] / ns : 0.0;
        int best = 0;
        if (var[1] > var[0])
            best = 1;
        if (var[2] > var[best])
            best = 2;
        L.next_axis = n_list[k] > 0 ? best : L.axis;
    }
    c->next_axis = c->lists[0].next_axis;

    // ---- 3. grids (host, identical on every rank), cell ranges (device), slice boxes + counts
    GridParams g[2];
    for (int k = 0; k < 2; k++) {
        g[k] = n_list[k] > 0
            ? choose_grid(list_host(c, k).stats, (int)n_list[k], c->grid_max_cells, c->grid_scale)
            : GridParams();
        if (W > 1 && n_list[k] > 0 && (long long)g[k].sy * g[k].sz < 8ll * W) {
            // too few cells to deal out by range: the replicated build with owner slices (what the
            // reference's dead _multigpu code did).  Same decision on every rank.
            c->sliced = false;
            build_boxes(c, inflation_radius);
            return;
        }
    }
    unsigned long long* d_splits = (unsigned long long*)c->b_xsplits.reserve((size_t)2 * (2 * kMaxWorld + 2) * 8);
    unsigned long long* d_xcnt = (unsigned long long*)c->b_xcnt.reserve(sizeof(unsigned long long) * kMaxWorld * kXWords);
    auto count_list = [&](int k) {
        auto& S = c->slice[k];
        auto& L = c->lists[k];
        auto& H = list_host(c, k);
        const int cnt = (int)(S.hi - S.lo);
        unsigned long long* sp = d_splits + (size_t)k * (2 * kMaxWorld + 2);
        const long long cells = (long long)g[k].sy * g[k].sz;
        uint32_t* hist = (uint32_t*)c->b_hist.reserve(std::max<size_t>((size_t)cells * 4, 4u << 20));
        launch_cell_splits(samp[k], S.ns, 1, g[k], W, hist, sp, st, c->lc);
        SCCD_CUDA(cudaMemcpyAsync(H.splits, sp, (size_t)(2 * W + 2) * 8, cudaMemcpyDeviceToHost, st));
        SCCD_CUDA(cudaMemsetAsync(L.copies.as<uint32_t>() + cnt, 0, 4, st));
        launch_expand_count(L.unsorted, cnt, g[k], nullptr, L.copies.as<uint32_t>(), st, c->lc);
        launch_scan_u32_to_u64(
            L.copies.as<uint32_t>(), L.offs.as<unsigned long long>(), cnt, c->b_scan_temp.ptr,
            c->b_scan_temp.cap, st, c->lc);
        SCCD_CUDA(cudaMemcpyAsync(
            &H.m, L.offs.as<unsigned long long>() + cnt, 8, cudaMemcpyDeviceToHost, st));
    };
    size_t kt_expand[2] = { 0, 0 };
    c->b_scan_temp.reserve(scan_temp_bytes((int)(std::max(n_list[0], n_list[1]) / W + 2)));
    for (int k = 0; k < 2; k++) {
        auto& S = c->slice[k];
        auto& L = c->lists[k];
        const int cnt = (int)(S.hi - S.lo);
        prepare_list(c, k, cnt, k == 0); // compact arrays: the slice only
        L.built_rank = rank;
        L.built_world = W;
        L.copies.reserve(((size_t)cnt + 1) * 4);
        L.offs.reserve(((size_t)cnt + 1) * 8);
        launch_list_boxes(mv, k, S.lo, 1, cnt, L.unsorted, L.axis, d_bad, st, c->lc);
        kt_expand[k] = kt_begin(c, &c->stats.ms_k_expand[k]);
        count_list(k);
        kt_end(c, kt_expand[k]);
    }
    SCCD_CUDA(cudaMemcpyAsync(c->h_flags, d_bad, 4, cudaMemcpyDeviceToHost, st));
    host_sync(c, st); // sync 2: cell ranges + this slice's record count
    int cell_bits[2] = { 0, 0 };
    for (int k = 0; k < 2; k++) {
        auto& H = list_host(c, k);
        // replication bound, decided on the SAMPLE estimate of the whole list so that every rank
        // takes the same decision (the slice counts differ from rank to rank)
        const unsigned long long m_cap = (unsigned long long)(c->grid_repl * (double)n_list[k]) + 1024;
        int attempt = 0;
        unsigned long long m_total = H.splits[W + 1] * (unsigned long long)c->slice[k].stride;
        while ((long long)g[k].sy * g[k].sz > 1 && m_total > m_cap) {
            if (++attempt > 12) {
                g[k] = GridParams();
            } else {
                if (g[k].sy >= g[k].sz)
                    g[k].sy = (g[k].sy + 1) / 2;
                else
                    g[k].sz = (g[k].sz + 1) / 2;
                regrid(g[k], H.stats);
            }
            if (W > 1 && (long long)g[k].sy * g[k].sz < 8ll * W) {
                c->sliced = false;
                build_boxes(c, inflation_radius);
                return;
            }
            count_list(k);
            host_sync(c, st);
            m_total = H.splits[W + 1] * (unsigned long long)c->slice[k].stride;
        }
        key_layout(g[k], m_total, H.stats, c->opt.key_steps, cell_bits[k]);
        if (H.m >= (1ull << 31))
            throw std::invalid_argument("more than 2^31 sweep records in one slice");
    }

    // ---- 4. records of the slice, grouped by owning rank; send counts
    SCCD_CUDA(cudaMemsetAsync(d_xcnt + (size_t)rank * kXWords, 0, sizeof(unsigned long long) * kXWords, st));
    const unsigned long long* send_rec[2] = { nullptr, nullptr };
    for (int k = 0; k < 2; k++) {
        auto& S = c->slice[k];
        auto& L = c->lists[k];
        auto& H = list_host(c, k);
        const int cnt = (int)(S.hi - S.lo);
        const size_t m = (size_t)H.m;
        unsigned long long* rec = (unsigned long long*)S.rec.reserve(std::max<size_t>(m, 1) * 8);
        const size_t kt = kt_begin(c, &c->stats.ms_k_expand[k]);
        launch_expand_fill_records(
            L.unsorted, cnt, g[k], L.offs.as<unsigned long long>(), (uint32_t)S.lo, rec, st, c->lc);
        unsigned long long* cnt_out = d_xcnt + (size_t)rank * kXWords + (size_t)k * kMaxWorld;
        if (W > 1) {
            // stable partition by the rank that owns the record's cell: one digit pass
            unsigned long long* rec2 = (unsigned long long*)S.rec_tmp.reserve(std::max<size_t>(m, 1) * 8);
            S.part_temp.reserve(partition_temp_bytes((long long)m));
            launch_partition_by_dest(
                (long long)m, rec, rec2, 32 + kKeyFlagBits + g[k].x_bits, H.splits, W, cnt_out,
                S.part_temp.ptr, S.part_temp.cap, st, c->lc);
            send_rec[k] = rec2;
        } else {
            SCCD_CUDA(cudaMemcpyAsync(cnt_out, &H.m, 8, cudaMemcpyHostToDevice, st));
            send_rec[k] = rec;
        }
        kt_end(c, kt);
    }
    // the "bad index" flag of this rank travels with its counts: every rank learns of it and
    // fails together instead of one rank leaving the others inside a collective
    SCCD_CUDA(cudaMemcpyAsync(
        d_xcnt + (size_t)rank * kXWords + 2 * kMaxWorld, c->h_flags, 4, cudaMemcpyHostToDevice, st));
    if (W > 1)
        SCCD_NCCL(nccl().AllGather(
            d_xcnt + (size_t)rank * kXWords, d_xcnt, kXWords, ncclUint64, comm_of(c), st));
    SCCD_CUDA(cudaMemcpyAsync(
        c->h_xcnt, d_xcnt, sizeof(unsigned long long) * (size_t)W * kXWords, cudaMemcpyDeviceToHost, st));
    host_sync(c, st); // sync 3: the send-count matrix
    for (int s = 0; s < W; s++)
        if ((uint32_t)c->h_xcnt[(size_t)s * kXWords + 2 * kMaxWorld])
            throw std::invalid_argument(
                "build_boxes: an edge / face refers to a vertex that does not exist");

    // ---- 5. exchange both lists in one group
    ExchangePlan plan[2];
    unsigned long long* recv[2] = { nullptr, nullptr };
    for (int k = 0; k < 2; k++) {
        plan[k] = exchange_plan(c->h_xcnt, k, rank, W);
        if (plan[k].recv_total >= (1ull << 27))
            throw std::invalid_argument("more than 2^27 sweep records in one list of one rank");
        recv[k] = (unsigned long long*)c->slice[k].recv.reserve(std::max<size_t>(plan[k].recv_total, 1) * 8);
        c->stats.n_records_sent[k] = (int64_t)(plan[k].send_total - plan[k].send_cnt[rank]);
        c->stats.n_records_received[k] = (int64_t)(plan[k].recv_total - plan[k].recv_cnt[rank]);
    }
    // One group per list: the vertex-face records on the main stream, the edge records on the
    // sort stream, so that the vertex-face sort starts while the edge records are still moving
    // (NCCL orders the two groups on the communicator; every rank issues them in this order).
    SCCD_CUDA(cudaEventRecord(c->ev_counts, st)); // everything the senders made is on `st`
    SCCD_CUDA(cudaStreamWaitEvent(c->sort_stream, c->ev_counts, 0));
    if (c->opt.profile)
        SCCD_CUDA(cudaEventRecord(c->ev_xa, st));
    for (int k = 0; k < 2; k++) {
        cudaStream_t sk = k == 0 ? st : c->sort_stream;
        const ExchangePlan& P = plan[k];
        if (W > 1)
            SCCD_NCCL(nccl().GroupStart());
        for (int p = 0; p < W; p++) {
            if (p == rank) {
                if (P.send_cnt[p])
                    SCCD_CUDA(cudaMemcpyAsync(
                        recv[k] + P.recv_off[p], send_rec[k] + P.send_off[p], P.send_cnt[p] * 8,
                        cudaMemcpyDeviceToDevice, sk));
                continue;
            }
            if (P.send_cnt[p])
                SCCD_NCCL(nccl().Send(
                    send_rec[k] + P.send_off[p], P.send_cnt[p], ncclUint64, p, comm_of(c), sk));
            if (P.recv_cnt[p])
                SCCD_NCCL(nccl().Recv(
                    recv[k] + P.recv_off[p], P.recv_cnt[p], ncclUint64, p, comm_of(c), sk));
        }
        if (W > 1)
            SCCD_NCCL(nccl().GroupEnd());
    }
    if (c->opt.profile)
        SCCD_CUDA(cudaEventRecord(c->ev_xb, c->sort_stream));

    // ---- 6. per list: sort the received records, rebuild their exact boxes.  The edge list on
    // the sort stream, under the vertex-face sweep and narrow phase (as build_boxes).
    for (int k = 0; k < 2; k++) {
        auto& S = c->slice[k];
        auto& L = c->lists[k];
        auto& H = list_host(c, k);
        cudaStream_t sk = k == 0 ? st : c->sort_stream;
        const size_t mm = std::max<size_t>(plan[k].recv_total, 1);
        GridParams gk = g[k];
        gk.cell_lo = (int)H.splits[rank];
        gk.cell_hi = (int)H.splits[rank + 1];
        L.sorted.n = (int)plan[k].recv_total;
        L.sorted.two_lists = k == 0;
        L.sorted.cell_sharded = true;
        L.sorted.grid = gk;
        L.sorted.box.x = (double2*)L.sx.reserve(mm * sizeof(double2));
        L.sorted.box.yz = (double4*)L.syz.reserve(mm * sizeof(double4));
        L.sorted.box.id = (int4*)L.sid.reserve(mm * sizeof(int4));
        L.sorted.pf.key = (uint32_t*)L.pkey.reserve(mm * 4);
        L.sorted.pf.reach = (uint32_t*)L.preach.reserve(mm * 4);
        L.sorted.pf.yz = (float4*)L.pyz.reserve(mm * sizeof(float4));
        unsigned long long* sorted_rec = (unsigned long long*)S.recv_sorted.reserve(mm * 8);
        S.sort_temp.reserve(sort_records_temp_bytes((long long)plan[k].recv_total));
        const bool prof = c->opt.profile != 0;
        if (prof)
            SCCD_CUDA(cudaEventRecord(c->ev[k == 0 ? EV_SB0 : EV_SB1], sk));
        launch_sort_records_and_rebuild(
            (int)plan[k].recv_total, cell_bits[k] + gk.x_bits, recv[k], sorted_rec, S.sort_temp.ptr,
            S.sort_temp.cap, mv, k, L.axis, L.sorted, sk, c->lc,
            prof ? c->ev[k == 0 ? EV_GA0 : EV_GA1] : nullptr,
            prof ? c->ev[k == 0 ? EV_GB0 : EV_GB1] : nullptr);
    }
    SCCD_CUDA(cudaEventRecord(c->ev_sorted1, c->sort_stream));
    c->sort1_pending = true;
    c->gather_timed = c->opt.profile != 0;
    record(c, EV_SORT);
    if (reuse) { // this build's statistics arrived with sync 2: they steer the next build
        adopt_stats(c, 0, n_list[0]);
        adopt_stats(c, 1, n_list[1]);
    }
    c->have_boxes = true;
    c->sliced = true;
    c->runs[0].bp_kind = c->runs[1].bp_kind = -1;
    for (int k = 0; k < 2; k++) {
        c->stats.n_boxes[k] = c->slice[k].hi - c->slice[k].lo;
        c->stats.n_records[k] = c->lists[k].sorted.n;
        c->stats.grid_cells[k][0] = g[k].sy;
        c->stats.grid_cells[k][1] = g[k].sz;
        c->stats.sweep_axis[k] = c->lists[k].axis;
        c->stats.next_axis[k] = c->lists[k].next_axis;
        c->stats.key_bits[k] = cell_bits[k] + g[k].x_bits;
    }
}

// sccd_ccd_sharded_host: this rank copies bytes [rank, rank + 1) / world of every mesh array
// from the host and the slices are all-gathered in place
void upload_mesh_sharded(
    sccd_ctx* c, const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF)
{
    if (nV < 0 || nE < 0 || nF < 0 || nV >= (1ll << 31) || nE >= (1ll << 31) || nF >= (1ll << 31))
        throw std::invalid_argument("upload_mesh: sizes out of range");
    if ((nV && (!V0 || !V1)) || (nE && !E) || (nF && !F))
        throw std::invalid_argument("upload_mesh: null pointer");
    const int W = std::max(c->comm_world, 1), rank = c->rank;
    const void* src[4] = { V0, V1, E, F };
    const size_t bytes[4] = { sizeof(double) * 3 * (size_t)nV, sizeof(double) * 3 * (size_t)nV,
                              sizeof(int32_t) * 2 * (size_t)nE, sizeof(int32_t) * 3 * (size_t)nF };
    char* dst[4];
    size_t chunk[4];
    for (int a = 0; a < 4; a++) {
        chunk[a] = ((bytes[a] + (size_t)W - 1) / W + 15) / 16 * 16;
        dst[a] = (char*)c->b_mesh_pad[a].reserve(chunk[a] * W + 16);
        const size_t lo = std::min(bytes[a], chunk[a] * (size_t)rank);
        const size_t hi = std::min(bytes[a], chunk[a] * (size_t)(rank + 1));
        if (hi > lo)
            SCCD_CUDA(cudaMemcpyAsync(
                dst[a] + lo, (const char*)src[a] + lo, hi - lo, cudaMemcpyHostToDevice, c->stream));
    }
    if (W > 1) {
        SCCD_NCCL(nccl().GroupStart());
        for (int a = 0; a < 4; a++)
            if (chunk[a])
                SCCD_NCCL(nccl().AllGather(
                    dst[a] + chunk[a] * rank, dst[a], chunk[a], ncclChar, comm_of(c), c->stream));
        SCCD_NCCL(nccl().GroupEnd());
    }
    upload_mesh(
        c, (const double*)dst[0], (const double*)dst[1], nV, (const int32_t*)dst[2], nE,
        (const int32_t*)dst[3], nF, /*on_device=*/true);
}

} // namespace host
} // namespace sccd

using namespace sccd::host;

extern "C" {

int sccd_exchange_plan(
    const uint64_t* counts, int list, int rank, int world, uint64_t* send_off, uint64_t* recv_cnt,
    uint64_t* recv_off, uint64_t* recv_total)
{
    if (!counts || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || list < 0 || list > 1)
        return SCCD_ERR_ARG;
    // same (rank x kXWords) layout the ranks all-gather: send counts of list 0, of list 1, flag
    std::vector<unsigned long long> all((size_t)world * kXWords, 0ull);
    for (int r = 0; r < world; r++)
        for (int d = 0; d < world; d++)
            all[(size_t)r * kXWords + list * kMaxWorld + d] = counts[(size_t)r * world + d];
    const ExchangePlan p = exchange_plan(all.data(), list, rank, world);
    for (int i = 0; i < world; i++) {
        if (send_off)
            send_off[i] = p.send_off[i];
        if (recv_cnt)
            recv_cnt[i] = p.recv_cnt[i];
        if (recv_off)
            recv_off[i] = p.recv_off[i];
    }
    if (recv_total)
        *recv_total = p.recv_total;
    return SCCD_OK;
}

int sccd_comm_get_unique_id(void* id_out)
{
    if (!id_out)
        return SCCD_ERR_ARG;
    NcclApi& n = nccl();
    if (!n.lib)
        return SCCD_ERR_STATE;
    ncclUniqueId id;
    if (n.GetUniqueId(&id) != ncclSuccess)
        return SCCD_ERR_CUDA;
    std::memcpy(id_out, &id, sizeof(id));
    return SCCD_OK;
}

int sccd_comm_create(sccd_ctx* ctx, const void* id, int rank, int world)
{
    return guarded(ctx, [&] {
        if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || (world > 1 && !id))
            throw std::invalid_argument("comm_create: bad rank / world / id");
        comm_destroy(ctx);
        if (world > 1) {
            NcclApi& n = nccl();
            if (!n.lib)
                throw std::logic_error(n.error);
            ncclUniqueId uid;
            std::memcpy(&uid, id, sizeof(uid));
            ncclComm_t comm = nullptr;
            SCCD_NCCL(n.CommInitRank(&comm, world, uid, rank));
            ctx->nccl_comm = comm;
            if (!getenv("SCCD_NO_PEER_TOI"))
                map_peer_toi(ctx, rank, world);
        }
        ctx->comm_world = world;
        ctx->rank = rank;
        ctx->world = world;
        ctx->have_boxes = false;
        ctx->runs[0].bp_kind = ctx->runs[1].bp_kind = -1;
        return SCCD_OK;
    });
}

int sccd_comm_destroy(sccd_ctx* ctx)
{
    return guarded(ctx, [&] {
        SCCD_CUDA(cudaStreamSynchronize(ctx->stream));
        SCCD_CUDA(cudaStreamSynchronize(ctx->sort_stream));
        comm_destroy(ctx);
        ctx->rank = 0;
        ctx->world = 1;
        ctx->sliced = false;
        ctx->have_boxes = false;
        ctx->runs[0].bp_kind = ctx->runs[1].bp_kind = -1;
        return SCCD_OK;
    });
}

int sccd_ccd_sharded(
    sccd_ctx* ctx, double min_distance, int max_iter, double tol, int allow_zero_toi, double* toi)
{
    return guarded(ctx, [&] {
        if (!toi)
            throw std::invalid_argument("ccd_sharded: null toi");
        run_pipeline(
            ctx, min_distance, max_iter, tol, allow_zero_toi != 0, false, toi, false, nullptr,
            nullptr, nullptr, /*sharded=*/true);
        return SCCD_OK;
    });
}

int sccd_ccd_sharded_host(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, double min_distance, int max_iter, double tol,
    int allow_zero_toi, double* toi)
{
    return guarded(ctx, [&] {
        if (!toi)
            throw std::invalid_argument("ccd_sharded: null toi");
        if (ctx->comm_world < 1)
            throw std::logic_error("sccd_ccd_sharded: no communicator (sccd_comm_create)");
        upload_mesh_sharded(ctx, V0, V1, nV, E, nE, F, nF);
        run_pipeline(
            ctx, min_distance, max_iter, tol, allow_zero_toi != 0, false, toi, false, nullptr,
            nullptr, nullptr, /*sharded=*/true);
        return SCCD_OK;
    });
}

} // extern "C"
