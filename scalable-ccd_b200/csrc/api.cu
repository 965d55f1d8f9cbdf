// Context, pipeline drivers and the C ABI (include/sccd.h).
//
// Host-side counterpart of the reference's cuda/ccd.cu, cuda/ipc_ccd_strategy.cu,
// BroadPhase (cuda/broad_phase/broad_phase.cu) and MemoryHandler
// (cuda/memory_handler.cpp): one stream, grow-only device buffers that persist across
// calls, and exactly one host synchronisation per broad-phase list (to size the pair
// buffer) and one per narrow-phase batch (to read the toi) -- against two syncs and two
// D2H copies per BFS level plus one per kernel in the reference.
#include "common.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace sccd {
namespace {
__global__ void find_splits_kernel(
    const unsigned long long* __restrict__ offsets, int n, int world, int* out)
{
    // out[r] = first owner of slice r, chosen so every slice has ~equal window work
    const int r = threadIdx.x;
    if (r > world)
        return;
    if (r == 0) {
        out[0] = 0;
        return;
    }
    if (r == world) {
        out[world] = n;
        return;
    }
    const unsigned long long total = offsets[n];
    const unsigned long long target = (total / (unsigned long long)world) * r;
    int a = 0, b = n;
    while (a < b) {
        const int mid = (a + b) >> 1;
        if (offsets[mid] < target)
            a = mid + 1;
        else
            b = mid;
    }
    out[r] = a;
}
} // namespace
} // namespace sccd

using namespace sccd;

#include "context.cuh"

namespace sccd {
namespace host {




// Stage / kernel timers are opt-in (SCCD_OPT_PROFILE): a 1 ms step is made of ~50 launches, and
// ~40 more event records were a sixth of the host's work per step.  Only the total is always timed.
inline bool timed_event(const sccd_ctx* c, int which)
{
    return c->opt.profile || which == EV_T0 || which == EV_T1 || which == EV_TMPA || which == EV_TMPB;
}
void record(sccd_ctx* c, int which)
{
    if (timed_event(c, which))
        SCCD_CUDA(cudaEventRecord(c->ev[which], c->stream));
}
float elapsed(sccd_ctx* c, int a, int b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0.f;
    }
    return ms;
}

// pooled event pair that will be resolved into *dst (+=) at the end of the call
size_t kt_alloc(sccd_ctx* c, float* dst)
{
    if (c->kt_used == c->ktimers.size()) {
        sccd_ctx::KTimer k;
        SCCD_CUDA(cudaEventCreate(&k.a));
        SCCD_CUDA(cudaEventCreate(&k.b));
        c->ktimers.push_back(k);
    }
    c->ktimers[c->kt_used].dst = dst;
    return c->kt_used++;
}
size_t kt_begin(sccd_ctx* c, float* dst, cudaStream_t st)
{
    if (!c->opt.profile)
        return (size_t)-1;
    const size_t id = kt_alloc(c, dst);
    c->ktimers[id].st = st ? st : c->cur->stream;
    SCCD_CUDA(cudaEventRecord(c->ktimers[id].a, c->ktimers[id].st));
    return id;
}
void kt_end(sccd_ctx* c, size_t id)
{
    if (id == (size_t)-1)
        return;
    SCCD_CUDA(cudaEventRecord(c->ktimers[id].b, c->ktimers[id].st));
}
void kt_resolve(sccd_ctx* c)
{
    for (size_t i = 0; i < c->kt_used; i++) {
        sccd_ctx::KTimer& k = c->ktimers[i];
        float ms = 0.f;
        if (cudaEventSynchronize(k.b) == cudaSuccess
            && cudaEventElapsedTime(&ms, k.a, k.b) == cudaSuccess)
            *k.dst += ms;
        else
            (void)cudaGetLastError();
    }
    c->kt_used = 0;
}

size_t budget_bytes(sccd_ctx* c)
{
    size_t free_b = c->mem_free, total_b = c->mem_total;
    if (c->mem_epoch != alloc_epoch() + 1) {
        SCCD_CUDA(cudaMemGetInfo(&free_b, &total_b));
        c->mem_free = free_b;
        c->mem_total = total_b;
        c->mem_epoch = alloc_epoch() + 1;
    }
    // memory_handler.cpp:11-30: 95 % of what is free, or the user's limit if smaller
    size_t avail = (size_t)(0.95 * (double)free_b);
    if (c->memory_limit) {
        const size_t used = total_b - free_b;
        const size_t user = c->memory_limit > used ? c->memory_limit - used : avail;
        avail = std::min(avail, user);
    }
    return avail;
}

void upload_mesh(
    sccd_ctx* c, const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, bool on_device)
{
    if (nV < 0 || nE < 0 || nF < 0 || nV >= (1ll << 31) || nE >= (1ll << 31) || nF >= (1ll << 31))
        throw std::invalid_argument("upload_mesh: sizes out of range");
    if ((nV && (!V0 || !V1)) || (nE && !E) || (nF && !F))
        throw std::invalid_argument("upload_mesh: null pointer");
    c->nV = (int)nV;
    c->nE = (int)nE;
    c->nF = (int)nF;
    if (on_device) {
        c->dV0 = V0;
        c->dV1 = V1;
        c->dE = E;
        c->dF = F;
    } else {
        const size_t vb = sizeof(double) * 3 * (size_t)nV;
        c->bV0.reserve(vb + 16);
        c->bV1.reserve(vb + 16);
        c->bE.reserve(sizeof(int32_t) * 2 * (size_t)nE + 16);
        c->bF.reserve(sizeof(int32_t) * 3 * (size_t)nF + 16);
        SCCD_CUDA(cudaMemcpyAsync(c->bV0.ptr, V0, vb, cudaMemcpyHostToDevice, c->stream));
        SCCD_CUDA(cudaMemcpyAsync(c->bV1.ptr, V1, vb, cudaMemcpyHostToDevice, c->stream));
        SCCD_CUDA(cudaMemcpyAsync(
            c->bE.ptr, E, sizeof(int32_t) * 2 * (size_t)nE, cudaMemcpyHostToDevice, c->stream));
        SCCD_CUDA(cudaMemcpyAsync(
            c->bF.ptr, F, sizeof(int32_t) * 3 * (size_t)nF, cudaMemcpyHostToDevice, c->stream));
        c->dV0 = c->bV0.as<double>();
        c->dV1 = c->bV1.as<double>();
        c->dE = c->bE.as<int32_t>();
        c->dF = c->bF.as<int32_t>();
    }
    c->have_mesh = true;
    c->have_boxes = false;
    c->runs[0].bp_kind = c->runs[1].bp_kind = -1;
}

// Frame-to-frame: same topology (E, F stay where the last upload put them), new positions.
void update_vertices(sccd_ctx* c, const double* V0, const double* V1, int64_t nV, bool on_device)
{
    if (!c->have_mesh)
        throw std::logic_error("update_vertices: no mesh uploaded");
    if (nV != c->nV)
        throw std::invalid_argument("update_vertices: vertex count differs from the uploaded mesh");
    if (nV && (!V0 || !V1))
        throw std::invalid_argument("update_vertices: null pointer");
    if (on_device) {
        c->dV0 = V0;
        c->dV1 = V1;
    } else {
        const size_t vb = sizeof(double) * 3 * (size_t)nV;
        c->bV0.reserve(vb + 16);
        c->bV1.reserve(vb + 16);
        SCCD_CUDA(cudaMemcpyAsync(c->bV0.ptr, V0, vb, cudaMemcpyHostToDevice, c->stream));
        SCCD_CUDA(cudaMemcpyAsync(c->bV1.ptr, V1, vb, cudaMemcpyHostToDevice, c->stream));
        c->dV0 = c->bV0.as<double>();
        c->dV1 = c->bV1.as<double>();
    }
    c->have_boxes = false;
    c->runs[0].bp_kind = c->runs[1].bp_kind = -1;
}

void prepare_list(sccd_ctx* c, int which, int n, bool two_lists)
{
    auto& L = c->lists[which];
    const size_t m = (size_t)std::max(n, 1);
    L.unsorted.x = (double2*)L.ux.reserve(m * sizeof(double2));
    L.unsorted.yz = (double4*)L.uyz.reserve(m * sizeof(double4));
    L.unsorted.id = (int4*)L.uid.reserve(m * sizeof(int4));
    L.n_boxes = n;
    L.sorted.n = 0;
    L.sorted.two_lists = two_lists;
}

// Choose the (y, z) cell grid of a list from its box statistics: cells about twice the mean
// box extent (so a box touches ~1.5 cells per axis), at most 1024 per axis / 2^20 in total.
GridParams choose_grid(const double* st, int n, int max_cells, double scale)
{
    GridParams g;
    if (n <= 0 || max_cells == 1)
        return g;
    const double ext[2] = { st[1] - st[0], st[3] - st[2] };
    const double mean[2] = { st[4] / n, st[5] / n };
    int s[2] = { 1, 1 };
    for (int a = 0; a < 2; a++) {
        if (!(ext[a] > 0) || !(mean[a] >= 0) || !std::isfinite(ext[a]))
            continue;
        const double cell = std::max(scale * mean[a], ext[a] / 1024.0);
        const double k = cell > 0 ? std::floor(ext[a] / cell) : 1.0;
        s[a] = (int)std::min(1024.0, std::max(1.0, k));
    }
    const long long cap = max_cells > 0 ? max_cells : (1ll << 20);
    while ((long long)s[0] * s[1] > cap) {
        if (s[0] >= s[1])
            s[0] = (s[0] + 1) / 2;
        else
            s[1] = (s[1] + 1) / 2;
    }
    g.sy = s[0];
    g.sz = s[1];
    g.y0 = st[0];
    g.z0 = st[2];
    g.inv_hy = g.sy > 1 ? g.sy / ext[0] : 0.0;
    g.inv_hz = g.sz > 1 ? g.sz / ext[1] : 0.0;
    return g;
}

// DeviceAABBs constructor + BroadPhase::build of the reference, for one list whose unsorted
// exact records are on the device: grid choice, replication into cells, radix sort, gather.
//
// Split in phases so that both mesh lists share their host synchronisations:
//   list_stats()       (caller-made boxes only; mesh lists get theirs from the box kernels)
//   sort_list_begin()  host: grid from the statistics; device: records per box (+ multi-GPU
//                      cell ranges), results on their way to pinned host memory
//   -- one stream sync --
//   sort_list_finish() host: accept the grid (or coarsen and retry); device: keys, radix sort,
//                      gather.
sccd_ctx::ListHost& list_host(sccd_ctx* c, int which)
{
    if (!c->h_lists)
        SCCD_CUDA(cudaMallocHost((void**)&c->h_lists, 3 * sizeof(sccd_ctx::ListHost)));
    return c->h_lists[which];
}

void regrid(GridParams& g, const double st[kNumStats])
{
    g.inv_hy = g.sy > 1 ? g.sy / (st[1] - st[0]) : 0.0;
    g.inv_hz = g.sz > 1 ? g.sz / (st[3] - st[2]) : 0.0;
}

// 32-bit key = [cell | q(x) | 3 flag bits] (common.cuh).  x gets as many bits as the digit
// passes needed for ~8 quantisation steps per record of an average cell leave room for
// (measured on config 2: 8 steps save a digit pass per list over 32 and add 0.01 % ties).
// m_total: records of the whole list (all ranks); st: its box statistics.
void key_layout(
    GridParams& g, unsigned long long m_total, const double* st, int key_steps, int& cell_bits)
{
    cell_bits = 0;
    while ((1ll << cell_bits) < (long long)g.sy * g.sz)
        cell_bits++;
    const int x_max = 32 - kKeyFlagBits - cell_bits;
    const double per_cell = (double)m_total / (double)((long long)g.sy * g.sz);
    const int steps = key_steps; // log2 of the quantisation steps per record
    int want = steps;
    while (want < x_max && (double)(1ll << (want - steps)) < per_cell)
        want++;
    want = std::min(std::max(want, 9), x_max);
    const int passes = (cell_bits + want + 7) / 8;
    g.x_bits = std::min(x_max, passes * 8 - cell_bits);
    const double ext = st[7] - st[6];
    g.x0 = st[6];
    g.inv_hx = (ext > 0 && std::isfinite(ext)) ? std::ldexp(1.0, g.x_bits) / ext : 0.0;
    if (!std::isfinite(g.inv_hx))
        g.inv_hx = 0.0;
}

// enqueue the record count of grid L.g_try (and, multi-GPU, this rank's cell range)
void sort_list_count(sccd_ctx* c, int which, cudaStream_t st = nullptr)
{
    if (!st)
        st = st;
    auto& L = c->lists[which];
    auto& H = list_host(c, which);
    const int n = L.n_boxes;
    const GridParams& g = L.g_try;
    const long long cells = (long long)g.sy * g.sz;
    L.try_sharded = false;
    if (cells == 1) { // plain 1-axis sweep: one record per box, nothing to ask the device
        H.m = (unsigned long long)n;
        SCCD_CUDA(cudaMemsetAsync(L.copies.as<uint32_t>() + n, 0, 4, st));
        launch_expand_count(L.unsorted, n, g, nullptr, L.copies.as<uint32_t>(), st, c->lc);
        launch_scan_u32_to_u64(
            L.copies.as<uint32_t>(), L.offs.as<unsigned long long>(), n, L.scan_temp.ptr,
            L.scan_temp.cap, st, c->lc);
        return;
    }
    const unsigned long long* d_range = nullptr;
    if (c->world > 1 && cells >= 8ll * c->world) {
        // Multi-GPU: this rank makes, sorts and sweeps only the records of its own contiguous
        // cell range (cells are independent sweep domains).  Every rank derives the same
        // ranges from its replica of the boxes -- no exchange.
        const int W = c->world;
        L.try_stride = (int)std::min<long long>(16, std::max<long long>(1, n / (cells * 64)));
        // shared by the lists (stream order keeps them apart): sized for the largest grid
        uint32_t* hist = (uint32_t*)c->b_hist.reserve(std::max<size_t>((size_t)cells * 4, 4u << 20));
        // one split record per list: the two lists are in flight together
        unsigned long long* d_out =
            (unsigned long long*)c->b_splits.reserve((size_t)3 * (2 * 16 + 2) * 8)
            + (size_t)which * (2 * 16 + 2);
        launch_cell_splits(L.unsorted, n, L.try_stride, g, W, hist, d_out, st, c->lc);
        SCCD_CUDA(cudaMemcpyAsync(
            H.splits, d_out, (size_t)(2 * W + 2) * 8, cudaMemcpyDeviceToHost, st));
        d_range = d_out + c->rank;
        L.try_sharded = true;
    }
    SCCD_CUDA(cudaMemsetAsync(L.copies.as<uint32_t>() + n, 0, 4, st));
    launch_expand_count(L.unsorted, n, g, d_range, L.copies.as<uint32_t>(), st, c->lc);
    launch_scan_u32_to_u64(
        L.copies.as<uint32_t>(), L.offs.as<unsigned long long>(), n, L.scan_temp.ptr,
        L.scan_temp.cap, st, c->lc);
    SCCD_CUDA(cudaMemcpyAsync(
        &H.m, L.offs.as<unsigned long long>() + n, 8, cudaMemcpyDeviceToHost, st));
}

// enqueue the statistics of a list towards list_host(c, which).stats_next (no sync)
void list_stats(sccd_ctx* c, int which, cudaStream_t st = nullptr)
{
    if (!st)
        st = c->stream;
    auto& L = c->lists[which];
    auto& H = list_host(c, which);
    if (L.n_boxes <= 0)
        return;
    // one partials + result region per list: the lists are in flight together
    const size_t per = (size_t)(kStatsBlocks + 1) * kNumStats;
    double* base = (double*)c->b_stats.reserve(3 * per * sizeof(double)) + which * per;
    double* d_stats = base + kStatsBlocks * kNumStats;
    const int stride = stats_stride(L.n_boxes);
    launch_box_stats(L.unsorted, L.n_boxes, stride, base, d_stats, st, c->lc);
    SCCD_CUDA(cudaMemcpyAsync(
        H.stats_next, d_stats, kNumStats * sizeof(double), cudaMemcpyDeviceToHost, st));
}
// the statistics that arrived become the ones the next grid is chosen from
void adopt_stats(sccd_ctx* c, int which, long long n_full)
{
    auto& L = c->lists[which];
    auto& H = list_host(c, which);
    std::memcpy(H.stats, H.stats_next, sizeof(H.stats));
    L.stats_valid = n_full > 0;
    L.stats_n = (int)n_full;
    L.stats_axis = L.axis;
    L.stats_f32 = c->f32;
}
bool stats_reusable(const sccd_ctx* c, int which, long long n_full)
{
    const auto& L = c->lists[which];
    return c->opt.reuse_grid != 0
        && (n_full == 0 || (L.stats_valid && L.stats_n == n_full && L.stats_axis == L.axis && L.stats_f32 == c->f32));
}
// statistics a previous build left in flight on the sort stream (long since there)
void drain_stats(sccd_ctx* c)
{
    if (!c->stats_in_flight)
        return;
    SCCD_CUDA(cudaEventSynchronize(c->ev_stats));
    adopt_stats(c, 0, c->lists[0].n_boxes);
    adopt_stats(c, 1, c->lists[1].n_boxes);
    c->stats_in_flight = false;
}

// requires the list's statistics in list_host(c, which).stats
void sort_list_begin(sccd_ctx* c, int which, cudaStream_t st = nullptr)
{
    auto& L = c->lists[which];
    auto& H = list_host(c, which);
    const int n = L.n_boxes;
    L.built_rank = c->rank;
    L.built_world = c->world;
    L.attempt = 0;
    if (n <= 0)
        return;
    L.g_try = choose_grid(H.stats, n, c->grid_max_cells, c->grid_scale);
    L.copies.reserve(((size_t)n + 1) * 4);
    L.offs.reserve(((size_t)n + 1) * 8);
    L.scan_temp.reserve(scan_temp_bytes(n));
    sort_list_count(c, which, st);
}

// st: the stream the fill / sort / gather are enqueued on (the retry path stays on c->stream)
void sort_list_finish(
    sccd_ctx* c, int which, cudaEvent_t ga, cudaEvent_t gb, cudaStream_t st = nullptr)
{
    if (!st)
        st = c->stream;
    auto& L = c->lists[which];
    auto& H = list_host(c, which);
    const int n = L.n_boxes;
    if (n <= 0) {
        L.sorted.n = 0;
        L.sorted.grid = GridParams();
        L.sorted.cell_sharded = false;
        if (ga)
            SCCD_CUDA(cudaEventRecord(ga, st));
        if (gb)
            SCCD_CUDA(cudaEventRecord(gb, st));
        return;
    }
    // replication bound: a few huge boxes can touch every cell -- coarsen until it is modest
    const unsigned long long m_cap = (unsigned long long)(c->grid_repl * n) + 1024;
    GridParams g = L.g_try;
    unsigned long long m = 0, m_total = 0;
    for (;;) {
        g = L.g_try;
        m = H.m;
        m_total = m;
        if (L.try_sharded) // estimate (exact for stride 1) of the records of ALL ranks
            m_total = H.splits[c->world + 1] * (unsigned long long)L.try_stride;
        if ((long long)g.sy * g.sz == 1 || m_total <= m_cap)
            break;
        if (++L.attempt > 12) // give up on the grid rather than explode memory
            L.g_try = GridParams();
        else {
            if (L.g_try.sy >= L.g_try.sz)
                L.g_try.sy = (L.g_try.sy + 1) / 2;
            else
                L.g_try.sz = (L.g_try.sz + 1) / 2;
            regrid(L.g_try, H.stats);
        }
        sort_list_count(c, which);
        host_sync(c, c->stream);
    }
    const bool sharded = L.try_sharded && (long long)g.sy * g.sz > 1;
    if (sharded) {
        g.cell_lo = (int)H.splits[c->rank];
        g.cell_hi = (int)H.splits[c->rank + 1];
    }
    L.sorted.cell_sharded = sharded;
    if (m >= (1ull << 27))
        throw std::invalid_argument("more than 2^27 sweep records in one list");
    int cell_bits = 0;
    key_layout(g, m_total, H.stats, c->opt.key_steps, cell_bits);
    const size_t mm = (size_t)m;
    L.keys.reserve(mm * 8);     // 64-bit records (key << 32 | box index) ...
    L.keys_tmp.reserve(mm * 8); // ... and their ping-pong buffer
    L.sorted.n = (int)m;
    L.sorted.grid = g;
    L.sorted.box.x = (double2*)L.sx.reserve(mm * sizeof(double2));
    L.sorted.box.yz = (double4*)L.syz.reserve(mm * sizeof(double4));
    L.sorted.box.id = (int4*)L.sid.reserve(mm * sizeof(int4));
    L.sorted.pf.key = (uint32_t*)L.pkey.reserve(mm * 4);
    L.sorted.pf.reach = (uint32_t*)L.preach.reserve(mm * 4);
    L.sorted.pf.yz = (float4*)L.pyz.reserve(mm * sizeof(float4));
    L.sort_temp.reserve(sort_temp_bytes((int)m)); // per list: the two sorts may overlap
    // (no cross-stream wait: the host has waited for this list's count -- wherever it ran -- and
    // for the boxes before it)
    const int slot = which == 1 ? 1 : 0;
    c->stats.key_bits[slot] = cell_bits + g.x_bits;
    const size_t kt_e = kt_begin(c, &c->stats.ms_k_expand[slot], st);
    uint32_t* sort_hist = launch_sort_prepare(L.sort_temp.ptr, L.sort_temp.cap, (long long)m, st);
    launch_expand_fill(
        L.unsorted, n, g, L.offs.as<unsigned long long>(), L.keys.as<unsigned long long>(), sort_hist,
        cell_bits + g.x_bits, st, c->lc);
    kt_end(c, kt_e);
    // (radix passes = from here to the gather's begin event; resolved in finish_stats)
    if (ga)
        SCCD_CUDA(cudaEventRecord(c->ev[slot == 0 ? EV_SB0 : EV_SB1], st));
    launch_sort_and_gather(
        (int)m, cell_bits + g.x_bits, L.keys.as<unsigned long long>(),
        L.keys_tmp.as<unsigned long long>(), L.sort_temp.ptr, L.sort_temp.cap, L.unsorted, L.sorted,
        st, c->lc, ga, gb, /*hist_ready=*/true);
}

// one list on its own (caller-made boxes; re-sharding an already built list)
void sort_list(sccd_ctx* c, int which, cudaEvent_t ga, cudaEvent_t gb)
{
    list_stats(c, which);
    host_sync(c, c->stream);
    adopt_stats(c, which, c->lists[which].n_boxes);
    sort_list_begin(c, which);
    host_sync(c, c->stream);
    sort_list_finish(c, which, ga, gb);
}

// make `st` wait for the edge-list sort that may still run on the sort stream
void join_sort_stream(sccd_ctx* c, cudaStream_t st)
{
    if (c->sort1_pending) {
        SCCD_CUDA(cudaStreamWaitEvent(st, c->ev_sorted1, 0));
        c->sort1_pending = false;
    }
}

void build_boxes(sccd_ctx* c, double inflation_radius)
{
    if (!c->have_mesh)
        throw std::logic_error("build_boxes: no mesh uploaded");
    join_sort_stream(c, c->stream); // a previous build whose edge list nobody swept
    drain_stats(c);
    c->sliced = false;
    const int nV = c->nV, nE = c->nE, nF = c->nF;
    const long long nVF = (long long)nV + nF;
    if (nVF >= (1ll << 27) || nE >= (1 << 27))
        throw std::invalid_argument("build_boxes: more than 2^27 boxes in one list");
    c->b_vtab.reserve(sizeof(VertexRec) * (size_t)std::max(nV, 1));
    c->b_vbox.reserve(sizeof(double) * 6 * (size_t)std::max(nV, 1));
    prepare_list(c, 0, (int)nVF, true);
    prepare_list(c, 1, nE, false);

    // aabb.cu:31-34: the radius itself is rounded up once (float build: converted to Scalar
    // first, scalar.hpp:43-49)
    const double radius_up = c->f32
        ? (double)std::nextafterf((float)inflation_radius, FLT_MAX)
        : std::nextafter(inflation_radius, DBL_MAX);
    auto& LV = c->lists[0];
    auto& LE = c->lists[1];
    // SCCD_OPT_SWEEP_AXIS: a fixed axis, or (-1) the axis the previous build handed back
    LV.axis = c->opt.sweep_axis >= 0 ? c->opt.sweep_axis : LV.next_axis;
    LE.axis = c->opt.sweep_axis >= 0 ? c->opt.sweep_axis : LE.next_axis;
    int* d_bad = (int*)c->b_flags.reserve(64);
    if (!c->h_flags)
        SCCD_CUDA(cudaMallocHost((void**)&c->h_flags, 64));
    SCCD_CUDA(cudaMemsetAsync(d_bad, 0, 4, c->stream));
    const size_t kt_boxes = kt_begin(c, &c->stats.ms_k_boxes);
    launch_mesh_boxes(
        c->dV0, c->dV1, nV, radius_up, c->f32, c->b_vtab.as<VertexRec>(), c->b_vbox.as<double>(),
        c->dE, nE, c->dF, nF, LE.unsorted, LV.unsorted, LE.axis, LV.axis, d_bad, c->stream, c->lc);
    kt_end(c, kt_boxes);
    record(c, EV_BUILD);
    // Frame-to-frame: the statistics only steer the cell grid and the key
    // quantisation, and any grid gives the same overlap set.  When the lists have the size,
    // axis and scalar type of the previous build, its statistics choose this build's grid at
    // once -- no host sync, and the statistics of the new boxes are computed on the second
    // stream, off the critical path, for the build after this one.
    const bool reuse = stats_reusable(c, 0, c->lists[0].n_boxes) && stats_reusable(c, 1, c->lists[1].n_boxes);
    SCCD_CUDA(cudaMemcpyAsync(c->h_flags, d_bad, 4, cudaMemcpyDeviceToHost, c->stream));
    if (reuse) {
        SCCD_CUDA(cudaEventRecord(c->ev_boxes, c->stream));
        SCCD_CUDA(cudaStreamWaitEvent(c->sort_stream, c->ev_boxes, 0));
        // (the statistics themselves are enqueued below, behind the edge list's record count)
    } else {
        // sync 1: statistics of both lists (sync 2: record counts of both lists)
        list_stats(c, 0);
        list_stats(c, 1);
        host_sync(c, c->stream);
        adopt_stats(c, 0, c->lists[0].n_boxes);
        adopt_stats(c, 1, c->lists[1].n_boxes);
        if (c->h_flags[0]) // the reference would read out of bounds (aabb.cu:199-226)
            throw std::invalid_argument(
                "build_boxes: an edge / face refers to a vertex that does not exist");
    }
    for (int which = 0; which < 2; which++) {
        // next sweep axis = argmax of the variance of the box centres, as sort_and_sweep hands
        // it back (sort_and_sweep.cpp:176-195); here over the sampled boxes of the statistics
        auto& L = c->lists[which];
        const double* st = list_host(c, which).stats;
        const int stride = stats_stride(L.n_boxes);
        const double ns = (double)((L.n_boxes + stride - 1) / stride);
        double var[3] = { 0, 0, 0 }; // in the caller's axes
        for (int k = 0; k < 3; k++)
            var[(L.axis + k) % 3] = ns > 0 ? st[11 + k] - st[8 + k] * st[8 + k] / ns : 0.0;
        int best = 0;
        if (var[1] > var[0])
            best = 1;
        if (var[2] > var[best])
            best = 2;
        L.next_axis = L.n_boxes > 0 ? best : L.axis;
    }
    c->next_axis = LV.next_axis;
    // The two lists count their records side by side: the vertex-face list on the main stream,
    // the edge list on the sort stream (single GPU; the cell-range split of sccd_set_shard shares
    // its histogram scratch between the lists and stays on one stream).  The host picks the
    // vertex-face count up first and enqueues that list's fill / sort / gather at once.
    const bool side_by_side = c->world == 1;
    if (side_by_side) {
        if (!reuse) { // (else: done above, the sort stream already waits for the boxes)
            SCCD_CUDA(cudaEventRecord(c->ev_boxes, c->stream));
            SCCD_CUDA(cudaStreamWaitEvent(c->sort_stream, c->ev_boxes, 0));
        }
        sort_list_begin(c, 1, c->sort_stream);
        SCCD_CUDA(cudaEventRecord(c->ev_cnt1, c->sort_stream));
    }
    if (reuse) { // this build's statistics, for the next one: nobody waits for them now
        list_stats(c, 0, c->sort_stream);
        list_stats(c, 1, c->sort_stream);
        SCCD_CUDA(cudaEventRecord(c->ev_stats, c->sort_stream));
        c->stats_in_flight = true;
    }
    sort_list_begin(c, 0);
    if (!side_by_side)
        sort_list_begin(c, 1);
    host_sync(c, c->stream);
    if (reuse && c->h_flags[0])
        throw std::invalid_argument("build_boxes: an edge / face refers to a vertex that does not exist");
    const bool prof = c->opt.profile != 0;
    sort_list_finish(c, 0, prof ? c->ev[EV_GA0] : nullptr, prof ? c->ev[EV_GB0] : nullptr);
    if (side_by_side) // (no extra round trip: it ran beside the count the host just waited for)
        SCCD_CUDA(cudaEventSynchronize(c->ev_cnt1));
    // Large lists: the edge list's sort waits for the vertex-face list's, so that it runs under
    // the vertex-face SWEEP -- a bandwidth-bound pass beside an issue-bound one -- instead of
    // beside another sort (config 4: 28.5 ms this way, 29.1 with the sorts side by side).  Small
    // lists are latency-bound and gain from starting at once (config 2: 0.676 -> 0.648 ms).
    if ((unsigned long long)LV.sorted.n + list_host(c, 1).m > (4ull << 20)) {
        SCCD_CUDA(cudaEventRecord(c->ev_counts, c->stream));
        SCCD_CUDA(cudaStreamWaitEvent(c->sort_stream, c->ev_counts, 0));
    }
    // (if the grid of list 1 has to be coarsened, its retry runs -- and syncs -- on the main
    // stream before anything more is enqueued on the sort stream)
    sort_list_finish(
        c, 1, prof ? c->ev[EV_GA1] : nullptr, prof ? c->ev[EV_GB1] : nullptr, c->sort_stream);
    SCCD_CUDA(cudaEventRecord(c->ev_sorted1, c->sort_stream));
    c->sort1_pending = true;
    c->gather_timed = prof;
    record(c, EV_SORT);
    c->have_boxes = true;
    c->runs[0].bp_kind = c->runs[1].bp_kind = -1;
    c->stats.n_boxes[0] = nVF;
    c->stats.n_boxes[1] = nE;
    c->stats.n_records[0] = LV.sorted.n;
    c->stats.n_records[1] = LE.sorted.n;
    c->stats.grid_cells[0][0] = LV.sorted.grid.sy;
    c->stats.grid_cells[0][1] = LV.sorted.grid.sz;
    c->stats.grid_cells[1][0] = LE.sorted.grid.sy;
    c->stats.grid_cells[1][1] = LE.sorted.grid.sz;
    for (int k = 0; k < 2; k++) {
        c->stats.sweep_axis[k] = c->lists[k].axis;
        c->stats.next_axis[k] = c->lists[k].next_axis;
    }
}

// stats are kept per reference pass (VF, EE); caller-made box lists report in slot 0
inline int stat_slot(int kind) { return kind == SCCD_EE ? 1 : 0; }

// sccd_set_boxes: caller-made AABBs -> exact records + keys (host, O(n)) -> device sort.
void set_boxes(
    sccd_ctx* c, const sccd_aabb* a, int64_t na, const sccd_aabb* b, int64_t nb, int sort_axis,
    int* next_axis)
{
    if (na < 0 || nb < 0 || (na && !a) || (nb && !b) || sort_axis < 0 || sort_axis > 2)
        throw std::invalid_argument("set_boxes: bad argument");
    const bool two = b != nullptr && nb > 0;
    const int64_t n = na + (two ? nb : 0);
    if (n >= (1ll << 27))
        throw std::invalid_argument("set_boxes: more than 2^27 boxes in one list");
    // the sweep always runs along the first coordinate of the record: rotate the axes
    const int ax = sort_axis, ay = (sort_axis + 1) % 3, az = (sort_axis + 2) % 3;
    std::vector<double2> hx((size_t)std::max<int64_t>(n, 1));
    std::vector<double4> hyz(hx.size());
    std::vector<int4> hid(hx.size());
    double s1[3] = { 0, 0, 0 }, s2[3] = { 0, 0, 0 };
    // variance accumulation in the reference's order: the boxes as swept, i.e. sorted on
    // min[axis] (sort_and_sweep.cpp:176-186); summation order only matters in the last ulp,
    // so the argmax is taken over the same quantities computed in input order.
    for (int64_t i = 0; i < n; i++) {
        const sccd_aabb& bx = i < na ? a[i] : b[i - na];
        hx[i] = make_double2(bx.min[ax], bx.max[ax]);
        hyz[i] = make_double4(bx.min[ay], bx.min[az], bx.max[ay], bx.max[az]);
        const int elem = (two && i < na) ? -bx.element_id - 1 : bx.element_id;
        hid[i] = make_int4(bx.vertex_ids[0], bx.vertex_ids[1], bx.vertex_ids[2], elem);
        for (int k = 0; k < 3; k++) {
            const double ctr = (bx.min[k] + bx.max[k]) / 2;
            s1[k] += ctr;
            s2[k] += ctr * ctr;
        }
    }
    if (next_axis) {
        double var[3];
        for (int k = 0; k < 3; k++)
            var[k] = n > 0 ? s2[k] - s1[k] * s1[k] / (double)n : 0.0;
        int best = 0;
        if (var[1] > var[0])
            best = 1;
        if (var[2] > var[best])
            best = 2;
        *next_axis = best;
    }
    prepare_list(c, 2, (int)n, two);
    auto& L = c->lists[2];
    if (n > 0) {
        SCCD_CUDA(cudaMemcpyAsync(L.unsorted.x, hx.data(), sizeof(double2) * n, cudaMemcpyHostToDevice, c->stream));
        SCCD_CUDA(cudaMemcpyAsync(L.unsorted.yz, hyz.data(), sizeof(double4) * n, cudaMemcpyHostToDevice, c->stream));
        SCCD_CUDA(cudaMemcpyAsync(L.unsorted.id, hid.data(), sizeof(int4) * n, cudaMemcpyHostToDevice, c->stream));
    }
    sort_list(c, 2, nullptr, nullptr);
    host_sync(c, c->stream); // host staging vectors go out of scope
    c->have_custom = true;
    c->runs[0].bp_kind = c->runs[1].bp_kind = -1;
    c->stats.n_boxes[0] = n;
    c->stats.n_records[0] = L.sorted.n;
}

// ---- everything below runs on c->cur (a Run: its stream, buffers and cursor) ----------------

void small_scratch(sccd_ctx* c)
{
    auto& R = *c->cur;
    R.b_small.reserve(256);
    if (!R.h_small)
        SCCD_CUDA(cudaMallocHost((void**)&R.h_small, 256));
}

void rrecord(sccd_ctx* c, int which)
{
    if (timed_event(c, which))
        SCCD_CUDA(cudaEventRecord(c->ev[which], c->cur->stream));
}

// BroadPhase::build: choose the list, the owner slice of this rank, enqueue the count pass,
// the scan and the D2H of the totals.  broad_phase_begin_finish() waits for them.
void broad_phase_begin_enqueue(sccd_ctx* c, int kind)
{
    auto& R = *c->cur;
    cudaStream_t st = R.stream;
    if (kind != SCCD_VF && kind != SCCD_EE && kind != SCCD_BOXES)
        throw std::invalid_argument("broad_phase: kind must be SCCD_VF, SCCD_EE or SCCD_BOXES");
    if (kind == SCCD_BOXES ? !c->have_custom : !c->have_boxes)
        throw std::logic_error("Must initialize build broad phase before detecting overlaps!");
    if (kind == SCCD_EE)
        join_sort_stream(c, st); // the edge list was sorted on the sort stream
    if (c->lists[kind].built_rank != c->rank || c->lists[kind].built_world != c->world) {
        if (c->sliced && kind != SCCD_BOXES)
            throw std::logic_error("broad_phase: the shard changed since the sliced build");
        sort_list(c, kind, nullptr, nullptr); // sccd_set_shard changed since the list was sorted
    }
    const SortedList& L = c->lists[kind].sorted;
    const int sk = stat_slot(kind);
    small_scratch(c);
    R.bp_kind = kind;
    R.shard_lo = 0;
    R.shard_hi = L.n;
    c->stats.n_pairs[sk] = 0;
    c->stats.n_candidates[sk] = 0;
    rrecord(c, sk == 0 ? EV_SW0A : EV_SW1A);
    if (c->world > 1 && L.n > 0 && !L.cell_sharded) {
        // too few cells to shard by cell range: every rank holds the whole sorted list and
        // sweeps one owner slice of it, slices balanced by sweep-window length
        R.b_counts.reserve(((size_t)L.n + 1) * 4);
        R.b_offsets.reserve(((size_t)L.n + 1) * 8);
        R.b_scan.reserve(scan_temp_bytes(L.n));
        SCCD_CUDA(cudaMemsetAsync(R.b_counts.as<uint32_t>() + L.n, 0, 4, st));
        launch_sweep_windows(L, R.b_counts.as<uint32_t>(), st, c->lc);
        launch_scan_u32_to_u64(
            R.b_counts.as<uint32_t>(), R.b_offsets.as<unsigned long long>(), L.n, R.b_scan.ptr,
            R.b_scan.cap, st, c->lc);
        if (c->world + 1 > 60)
            throw std::invalid_argument("set_shard: world too large");
        find_splits_kernel<<<1, 64, 0, st>>>(
            R.b_offsets.as<unsigned long long>(), L.n, c->world, R.b_small.as<int>());
        SCCD_CUDA(cudaGetLastError());
        c->lc.n++;
        SCCD_CUDA(cudaMemcpyAsync(
            R.h_small, R.b_small.ptr, sizeof(int) * (c->world + 1), cudaMemcpyDeviceToHost, st));
        host_sync(c, st);
        R.shard_lo = R.h_small[c->rank];
        R.shard_hi = R.h_small[c->rank + 1];
    }
    const int m = R.shard_hi - R.shard_lo;
    R.bp_cursor = R.shard_lo;
    R.bp_total = 0;
    R.bp_emitted = 0;
    if (m <= 0)
        return;
    R.b_counts.reserve(((size_t)m + 1) * 4);
    R.b_offsets.reserve(((size_t)m + 1) * 8);
    R.b_scan.reserve(scan_temp_bytes(m));
    SCCD_CUDA(cudaMemsetAsync(R.b_counts.as<uint32_t>() + m, 0, 4, st));
    unsigned long long* d_cand = reinterpret_cast<unsigned long long*>(R.b_small.as<char>() + 128);
    SCCD_CUDA(cudaMemsetAsync(d_cand, 0, 8, st));
    R.b_stage_pairs.reserve(sweep_stage_pair_bytes(m));
    R.b_stage_tags.reserve(sweep_stage_tag_bytes(m));
    R.b_stage_count.reserve(sweep_stage_tiles(m) * 4);
    const size_t kt = kt_begin(c, &c->stats.ms_k_sweep_count[sk]);
    launch_sweep_count(
        L, R.shard_lo, R.shard_hi, R.b_counts.as<uint32_t>(), d_cand, R.b_stage_pairs.ptr,
        R.b_stage_tags.ptr, R.b_stage_count.as<uint32_t>(), st, c->lc, c->opt.sweep_staged != 0);
    kt_end(c, kt);
    launch_scan_u32_to_u64(
        R.b_counts.as<uint32_t>(), R.b_offsets.as<unsigned long long>(), m, R.b_scan.ptr,
        R.b_scan.cap, st, c->lc);
    unsigned long long* h = reinterpret_cast<unsigned long long*>(R.h_small);
    SCCD_CUDA(cudaMemcpyAsync(
        &h[0], R.b_offsets.as<unsigned long long>() + m, 8, cudaMemcpyDeviceToHost, st));
    SCCD_CUDA(cudaMemcpyAsync(&h[1], d_cand, 8, cudaMemcpyDeviceToHost, st));
}

// one host sync (the totals)
void broad_phase_begin_finish(sccd_ctx* c)
{
    auto& R = *c->cur;
    if (R.shard_hi - R.shard_lo <= 0)
        return;
    host_sync(c, R.stream);
    unsigned long long* h = reinterpret_cast<unsigned long long*>(R.h_small);
    R.bp_total = h[0];
    c->stats.n_candidates[stat_slot(R.bp_kind)] = (int64_t)h[1];
}

void broad_phase_begin(sccd_ctx* c, int kind)
{
    broad_phase_begin_enqueue(c, kind);
    broad_phase_begin_finish(c);
}

bool broad_phase_complete(sccd_ctx* c) { return c->cur->bp_cursor >= c->cur->shard_hi; }

// pairs a chunk may hold: the caller's cap, everything when it is small, else what the memory
// budget allows (pair (8 B) + per-query narrow-phase state; MemoryHandler::per_overlap_memory_size)
unsigned long long chunk_budget(sccd_ctx* c, unsigned long long remaining)
{
    if (c->max_pairs_per_chunk > 0)
        return (unsigned long long)c->max_pairs_per_chunk;
    if (remaining * 32ull <= (256ull << 20) && !c->memory_limit)
        return remaining; // small batch: skip cudaMemGetInfo
    return std::max<size_t>(budget_bytes(c) + c->cur->b_pairs.cap, 1 << 20) / 32;
}

// BroadPhase::detect_overlaps_partial
void broad_phase_partial(sccd_ctx* c, const sccd_pair** d_pairs, int64_t* n_pairs)
{
    auto& R = *c->cur;
    cudaStream_t st = R.stream;
    if (R.bp_kind < 0)
        throw std::logic_error("Must initialize build broad phase before detecting overlaps!");
    *d_pairs = nullptr;
    *n_pairs = 0;
    if (broad_phase_complete(c))
        return;
    const int kind = R.bp_kind;
    const int sk = stat_slot(kind);
    const SortedList& L = c->lists[kind].sorted;
    const unsigned long long remaining = R.bp_total - R.bp_emitted;
    const unsigned long long budget = chunk_budget(c, remaining);
    int end = R.shard_hi;
    unsigned long long n_chunk = remaining;
    if (remaining > budget) {
        const int lo = R.bp_cursor - R.shard_lo, hi = R.shard_hi - R.shard_lo;
        launch_find_chunk_end(
            R.b_offsets.as<unsigned long long>(), lo, hi, budget, R.b_small.as<int>(), st, c->lc);
        SCCD_CUDA(cudaMemcpyAsync(
            R.h_small, R.b_small.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
        host_sync(c, st);
        const int e = R.h_small[0];
        if (e <= lo) // memory_handler.cpp:65-69
            throw std::runtime_error(
                "Insufficient memory to increase overlap size; "
                "cannot allocate even a single box's overlaps.");
        end = R.shard_lo + e;
        unsigned long long* h = reinterpret_cast<unsigned long long*>(R.h_small);
        SCCD_CUDA(cudaMemcpyAsync(
            &h[0], R.b_offsets.as<unsigned long long>() + e, 8, cudaMemcpyDeviceToHost, st));
        SCCD_CUDA(cudaMemcpyAsync(
            &h[1], R.b_offsets.as<unsigned long long>() + lo, 8, cudaMemcpyDeviceToHost, st));
        host_sync(c, st);
        n_chunk = h[0] - h[1];
    }
    if (n_chunk > 0) {
        R.b_pairs.reserve((size_t)n_chunk * sizeof(sccd_pair));
        const size_t kt = kt_begin(c, &c->stats.ms_k_sweep_fill[sk]);
        launch_sweep_fill(
            L, R.shard_lo, R.bp_cursor, end, R.b_offsets.as<unsigned long long>(),
            R.b_pairs.as<sccd_pair>(), R.b_stage_pairs.ptr, R.b_stage_tags.ptr,
            R.b_stage_count.as<uint32_t>(), st, c->lc);
        kt_end(c, kt);
    }
    R.bp_cursor = end;
    R.bp_emitted += n_chunk;
    c->stats.n_pairs[sk] += (int64_t)n_chunk;
    rrecord(c, sk == 0 ? EV_SW0B : EV_SW1B);
    *d_pairs = R.b_pairs.as<sccd_pair>();
    *n_pairs = (int64_t)n_chunk;
}

void narrow_setup(sccd_ctx* c, long long n_queries)
{
    auto& R = *c->cur;
    if (!R.h_counters)
        SCCD_CUDA(cudaMallocHost((void**)&R.h_counters, sizeof(NarrowCounters)));
    R.b_counters.reserve(sizeof(NarrowCounters));
    // two bounded lists of sub-boxes handed from round to round; sccd_set_queue_capacity gives
    // the number of items of each (MemoryHandler::MAX_UNIT_SIZE analogue).  Default: one item
    // per query, at least 1 Mi (64 MiB per list), at most 16 Mi (1 GiB per list).
    long long cap = c->queue_cap > 0
        ? c->queue_cap
        : std::min<long long>(std::max<long long>(n_queries, 1 << 20), 1 << 24);
    cap = std::max<long long>(cap, 64);
    if ((unsigned long long)cap != R.item_cap || !R.b_items[0].ptr) {
        for (int i = 0; i < 2; i++) {
            const void* before = R.b_items[i].ptr;
            R.b_items[i].reserve((size_t)cap * sizeof(WorkItem));
            // fresh memory: no slot may carry the ready mark of a work-queue launch by accident
            // (the marks are launch numbers, never zero: narrow.cu Round0::epoch)
            if (R.b_items[i].ptr != before)
                SCCD_CUDA(cudaMemsetAsync(R.b_items[i].ptr, 0, R.b_items[i].cap, R.stream));
        }
        R.item_cap = (unsigned long long)cap;
    }
}

// One narrow-phase batch, the body of narrow_phase<is_vf>() (narrow_phase.cu:108-206), in two
// halves: narrow_enqueue() puts every round on the run's stream, narrow_finish() waits, adds
// rounds for paths deeper than the walk state, and folds the counters into the statistics.
// d_gtoi: device word holding the running earliest toi (shared by concurrent batches).
void narrow_enqueue(
    sccd_ctx* c, int kind, const NarrowInput& in, double ms, int max_iter, double tol,
    bool allow_zero_toi, double* d_gtoi, double* d_toi_per_query, cudaEvent_t solver_waits_for)
{
    auto& R = *c->cur;
    cudaStream_t st = R.stream;
    R.pending.active = false;
    if (in.n <= 0 && solver_waits_for) // nothing to launch, but keep the streams ordered
        SCCD_CUDA(cudaStreamWaitEvent(st, solver_waits_for, 0));
    c->stats.n_queries[kind] += in.n;
    if (in.n <= 0)
        return;
    if (in.n >= (1ll << 32))
        throw std::invalid_argument("narrow_phase: more than 2^32 queries in one batch");
    narrow_setup(c, in.n);
    NarrowParams P;
    // float build: ms / tol are Scalar parameters of the reference's entry points, i.e. rounded
    // to float on the way in (narrow_phase.cuh:30-46 with Scalar = float)
    P.ms = c->f32 ? (double)(float)ms : ms;
    P.tol = c->f32 ? (double)(float)tol : tol;
    ms = P.ms;
    P.max_iter = max_iter;
    P.allow_zero_toi = allow_zero_toi ? 1 : 0;
    P.use_ms = ms > 0 ? 1 : 0;
    P.n_peers = (c->share_toi && !d_toi_per_query) ? c->n_peers : 0;
    for (int p = 0; p < 15; p++)
        P.peer_toi[p] = p < P.n_peers ? c->peer_toi[p] : nullptr;
    // (the edge-edge pass inherits the earliest toi of the vertex-face pass and may want other
    // budgets: SCCD_OPT_NARROW_FLAGS_EE)
    P.flags = (kind == SCCD_EE && c->opt.np_flags_ee >= 0) ? c->opt.np_flags_ee : c->opt.np_flags;
    P.max_depth = c->opt.np_depth;
    P.cap_drops = c->opt.cap_drops;
    P.solver = c->opt.np_solver;
    // lower bounds only where they can pay: the shared minimum (not the per-query list), and not
    // while the previous frames of this pass showed them useless (narrow_finish)
    P.want_tlb = !d_toi_per_query && c->tlb_pause[kind] == 0;
    // Work-queue launch: warps for ~1/24 of the survivors the previous batch of this kind had (at
    // least 32 CTAs).  Every resident warp starts a root at once, before any bound exists, and
    // what it finds after the bound has dropped was wasted: config 2's vertex-face pass (14,600
    // survivors) does 204 K box checks on 296 CTAs, 105 K on 148, 55 K on 74 (0.74 / 0.71 /
    // 0.69 ms per step), 29 K on 37 (0.72: too few warps to cut the deep trees).
    P.root_check = c->opt.np_cull == 1 || c->opt.np_cull == 3;
    P.cull_float = c->opt.np_cull != 3;
    P.tail_lanes = c->opt.np_tail_lanes;
    P.queue_ctas = c->opt.queue_ctas[kind];
    if (P.queue_ctas == 0 && c->np_last_survivors[kind] > 0 && c->opt.reuse_grid)
        P.queue_ctas = (int)std::max<long long>(32, (c->np_last_survivors[kind] + 191) / 192);
    if (!d_toi_per_query && c->tlb_pause[kind] > 0)
        c->tlb_pause[kind]--;
    SCCD_CUDA(cudaMemsetAsync(R.b_counters.ptr, 0, sizeof(NarrowCounters), st));
    if (d_toi_per_query)
        launch_fill_f64(d_toi_per_query, in.n, INFINITY, st, c->lc);
    unsigned int* checks = nullptr;
    R.checks_n = 0;
    if (max_iter >= 0) {
        checks = (unsigned int*)R.b_checks_q.reserve((size_t)in.n * 4);
        SCCD_CUDA(cudaMemsetAsync(checks, 0, (size_t)in.n * 4, st));
        R.checks_n = in.n;
    }
    // separating-axis cull in front of the solver (SCCD_NP_CULL=0 switches it off: A/B, tests)
    // survivor records of the cull (two buffers: as written, and sorted by lower-bound bucket),
    // every survivor's toi lower bound, and the scratch of that one-pass sort
    unsigned long long* survivors = nullptr;
    float* tlb = nullptr;
    if (c->opt.np_cull) {
        survivors = (unsigned long long*)R.b_surv.reserve((size_t)in.n * 16);
        tlb = (float*)R.b_tlb.reserve((size_t)in.n * 4);
        R.b_surv_sort.reserve(sort_survivors_temp_bytes(in.n));
    }
    // kernel-level timers: the cull always, every solver round with SCCD_OPT_PROFILE
    cudaEvent_t tev[2 * (1 + kNarrowRounds)] = {};
    if (survivors && c->opt.profile) {
        const size_t id = kt_alloc(c, &c->stats.ms_k_cull[kind]);
        tev[0] = c->ktimers[id].a, tev[1] = c->ktimers[id].b;
    }
    if (c->opt.profile)
        for (int r = 0; r < kNarrowRounds; r++) {
            const size_t id = kt_alloc(c, &c->stats.ms_k_round[kind][r]);
            tev[2 + 2 * r] = c->ktimers[id].a, tev[3 + 2 * r] = c->ktimers[id].b;
        }
    const size_t kt = kt_begin(c, &c->stats.ms_k_narrow[kind]);
    // (every round has its own timer in profile mode: all of them are launched then)
    const int hint = (c->opt.profile || !c->opt.reuse_grid) ? -1 : c->np_hint[kind];
    launch_narrow_phase(
        kind == SCCD_VF, c->f32, in, P, R.b_counters.as<NarrowCounters>(), d_gtoi,
        R.b_items[0].as<WorkItem>(), R.b_items[1].as<WorkItem>(), R.item_cap, d_toi_per_query,
        checks, survivors, tlb, R.b_surv_sort.ptr, R.b_surv_sort.cap, c->num_sms, st, c->lc, tev,
        solver_waits_for, hint, false);
    kt_end(c, kt);
    SCCD_CUDA(cudaMemcpyAsync(
        R.h_counters, R.b_counters.ptr, sizeof(NarrowCounters), cudaMemcpyDeviceToHost, st));
    SCCD_CUDA(cudaMemcpyAsync(&c->h_gtoi[0], d_gtoi, 8, cudaMemcpyDeviceToHost, st));
    R.pending.active = true;
    R.pending.kind = kind;
    R.pending.in = in;
    R.pending.P = P;
    R.pending.d_tq = d_toi_per_query;
    R.pending.checks = checks;
    R.pending.culling = survivors != nullptr;
    R.pending.hint = hint;
    R.pending.survivors = survivors;
    R.pending.tlb = tlb;
}

// returns whether it waited for the run's stream (which is then idle)
bool narrow_finish(sccd_ctx* c, double* d_gtoi)
{
    auto& R = *c->cur;
    cudaStream_t st = R.stream;
    if (!R.pending.active)
        return false;
    R.pending.active = false;
    const int kind = R.pending.kind;
    host_sync(c, st);
    if (R.pending.hint >= 0 && R.h_counters->n_items[0] != 0 && R.h_counters->round0_ran == 0) {
        // the survivor list is not of the length the launches were chosen for (a short list
        // where rounds were launched, or the other way round): none of them found work.  Launch
        // the solver again, every kernel this time.
        launch_narrow_phase(
            kind == SCCD_VF, c->f32, R.pending.in, R.pending.P, R.b_counters.as<NarrowCounters>(),
            d_gtoi, R.b_items[0].as<WorkItem>(), R.b_items[1].as<WorkItem>(), R.item_cap,
            R.pending.d_tq, R.pending.checks, R.pending.survivors, R.pending.tlb, R.b_surv_sort.ptr,
            R.b_surv_sort.cap, c->num_sms, st, c->lc, nullptr, nullptr, -1, true);
        SCCD_CUDA(cudaMemcpyAsync(
            R.h_counters, R.b_counters.ptr, sizeof(NarrowCounters), cudaMemcpyDeviceToHost, st));
        SCCD_CUDA(cudaMemcpyAsync(&c->h_gtoi[0], d_gtoi, 8, cudaMemcpyDeviceToHost, st));
        host_sync(c, st);
        c->stats.n_relaunched++;
    }
    if (R.h_counters->round0_ran)
        c->np_hint[kind] = R.h_counters->round0_ran == 2 ? 1 : 0;
    if (R.pending.culling)
        c->np_last_survivors[kind] = (long long)R.h_counters->n_items[0];
    // the last round only hands work on when a path outgrows the lane state: rerun it
    for (int extra = 0; R.h_counters->n_items[kNarrowRounds] != 0 && R.h_counters->overflow != 2;
         extra++) {
        if (extra > 64)
            throw std::runtime_error("narrow phase: bisection deeper than 8192 levels");
        launch_narrow_extra_round(
            kind == SCCD_VF, c->f32, R.pending.in, R.pending.P, R.b_counters.as<NarrowCounters>(), d_gtoi,
            R.b_items[0].as<WorkItem>(), R.b_items[1].as<WorkItem>(), R.item_cap, extra,
            R.pending.d_tq, R.pending.checks, c->num_sms, st, c->lc);
        SCCD_CUDA(cudaMemcpyAsync(
            R.h_counters, R.b_counters.ptr, sizeof(NarrowCounters), cudaMemcpyDeviceToHost, st));
        SCCD_CUDA(cudaMemcpyAsync(&c->h_gtoi[0], d_gtoi, 8, cudaMemcpyDeviceToHost, st));
        host_sync(c, st);
    }
    const NarrowCounters& r = *R.h_counters;
    c->stats.n_box_checks[kind] += (int64_t)r.box_checks;
    if (R.pending.culling) {
        c->stats.n_culled[kind] += R.pending.in.n - (int64_t)r.n_items[0];
        c->stats.n_skipped[kind] += (int64_t)r.n_items[0] - (int64_t)r.started;
        // did the lower bounds earn their keep?  (only judged where skipping was allowed)
        if (R.pending.P.want_tlb && R.pending.P.max_iter < 0 && !R.pending.d_tq
            && !(R.pending.P.flags & (1 << 7)) && r.n_items[0] >= 4096
            && (r.n_items[0] - r.started) * 2ull < r.n_items[0])
            // (measured: config 4's edge pass skips 25 % of its survivors and saves 7 % of its
            // box checks, for a cull that is 2 ms = 66 % longer; config 2 skips 66 % / 99 %)
        {
            // (again useless: ask less and less often -- 15, 31 .. 255 batches)
            c->tlb_pause_len[kind] = std::min(255, c->tlb_pause_len[kind] * 2 + 1);
            c->tlb_pause[kind] = c->tlb_pause_len[kind];
        } else if (R.pending.P.want_tlb) {
            c->tlb_pause_len[kind] = 7;
        }
    }
    c->stats.n_donated[kind] += (int64_t)r.donated;
    for (int i = 0; i <= kNarrowRounds; i++) {
        unsigned long long n = r.n_items[i];
        if (i > 0 && r.closed[i])
            n = std::min(n, ~r.closed[i]);
        c->stats.n_round_items[kind][i] += (int64_t)(i > 0 ? std::min(n, R.item_cap) : n);
    }
    for (int i = 0; i < kNarrowRounds; i++)
        c->stats.n_round_checks[kind][i] += (int64_t)r.round_checks[i];
    c->stats.n_capped[kind] += (int64_t)r.capped;
    if (r.overflow)
        c->stats.queue_overflow = 1;
    if (r.bad_input)
        throw std::invalid_argument("narrow_phase: a pair refers to an element that is not in the mesh");
    if (r.overflow == 2)
        throw std::runtime_error(
            "narrow phase: item list too small to hand on a sub-tree deeper than 128 levels; "
            "raise it with sccd_set_queue_capacity");
    return true;
}

// the shared earliest-toi word of the context
double* gtoi_set(sccd_ctx* c, double v, cudaStream_t st)
{
    if (!c->h_gtoi)
        SCCD_CUDA(cudaMallocHost((void**)&c->h_gtoi, 64));
    double* d = (double*)c->b_gtoi.reserve(64);
    c->h_gtoi[1] = v; // staging slot (slot 0 receives results)
    SCCD_CUDA(cudaMemcpyAsync(d, &c->h_gtoi[1], 8, cudaMemcpyHostToDevice, st));
    return d;
}
// synchronous batch on the current run: lowers *toi_inout in place
void narrow_run(
    sccd_ctx* c, int kind, const NarrowInput& in, double ms, int max_iter, double tol,
    bool allow_zero_toi, double* toi_inout, double* d_toi_per_query)
{
    if (!(*toi_inout >= 0))
        throw std::invalid_argument("narrow_phase: toi must be >= 0");
    if (in.n > 0 && !d_toi_per_query && *toi_inout <= 0) { // narrow_phase.cu:136
        c->stats.n_queries[kind] += in.n;
        return;
    }
    if (c->f32) // Scalar& toi of the float build
        *toi_inout = (double)(float)*toi_inout;
    double* d_gtoi = gtoi_set(c, *toi_inout, c->cur->stream);
    narrow_enqueue(c, kind, in, ms, max_iter, tol, allow_zero_toi, d_gtoi, d_toi_per_query, nullptr);
    if (!c->cur->pending.active)
        return;
    narrow_finish(c, d_gtoi);
    const double t = c->h_gtoi[0]; // copied back with the counters
    if (t < *toi_inout)
        *toi_inout = t;
}

NarrowInput mesh_input(sccd_ctx* c, const sccd_pair* d_pairs, int64_t n)
{
    if (!c->have_boxes)
        throw std::logic_error("narrow_phase: call sccd_build_boxes first");
    NarrowInput in;
    in.vtab = c->b_vtab.as<VertexRec>();
    in.E = c->dE;
    in.F = c->dF;
    in.nV = c->nV;
    in.nE = c->nE;
    in.nF = c->nF;
    in.pairs = d_pairs;
    in.n = n;
    return in;
}

void reset_stats(sccd_ctx* c)
{
    c->stats = sccd_stats {};
    c->lc.n = 0;
}

void finish_stats(sccd_ctx* c, bool pipeline)
{
    c->stats.n_launches = c->lc.n;
    kt_resolve(c);
    // (not while the edge list is still being sorted on the sort stream: waiting for its gather
    // here would serialise what build_boxes() just overlapped -- a later call picks it up)
    if (c->gather_timed && !c->sort1_pending) {
        SCCD_CUDA(cudaEventSynchronize(c->ev[EV_GB1]));
        c->stats.ms_k_gather = elapsed(c, EV_GA0, EV_GB0) + elapsed(c, EV_GA1, EV_GB1);
        c->stats.ms_k_sort[0] = elapsed(c, EV_SB0, EV_GA0);
        c->stats.ms_k_sort[1] = elapsed(c, EV_SB1, EV_GA1);
        c->gather_timed = false;
    }
    if (!pipeline)
        return;
    SCCD_CUDA(cudaEventSynchronize(c->ev[EV_T1]));
    c->stats.ms_total = elapsed(c, EV_T0, EV_T1);
    if (!c->opt.profile)
        return;
    c->stats.ms_build = elapsed(c, EV_T0, EV_BUILD);
    c->stats.ms_sort = elapsed(c, EV_BUILD, EV_SORT);
    c->stats.ms_sweep[0] = elapsed(c, EV_SW0A, EV_SW0B);
    c->stats.ms_sweep[1] = elapsed(c, EV_SW1A, EV_SW1B);
    c->stats.ms_narrow[0] = elapsed(c, EV_NP0A, EV_NP0B);
    c->stats.ms_narrow[1] = elapsed(c, EV_NP1A, EV_NP1B);
    c->stats.ms_total = elapsed(c, EV_T0, EV_T1);
    if (c->sliced && c->ev_xa) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_xa, c->ev_xb) == cudaSuccess)
            c->stats.ms_exchange = ms;
        else
            (void)cudaGetLastError();
    }
}

// ccd() body (ccd.cu:108-146) and ipc_ccd_strategy() body (ipc_ccd_strategy.cu:108-152).
void run_pipeline(
    sccd_ctx* c, double min_distance, int max_iter, double tol, bool allow_zero_toi, bool ipc,
    double* toi_out, bool want_collisions, std::vector<sccd_pair>* coll_ids,
    std::vector<double>* coll_toi, int64_t* n_coll, bool sharded)
{
    reset_stats(c);
    c->cur = &c->runs[0];
    record(c, EV_T0);
    if (sharded && (ipc || want_collisions))
        throw std::logic_error("sharded pipeline: only the plain ccd() is supported");
    // (sharded: the solver kernels publish the bound to the other ranks' toi words)
    struct ShareGuard {
        sccd_ctx* c;
        ~ShareGuard() { c->share_toi = false; }
    } share_guard { c };
    c->share_toi = sharded;
    if (sharded) // collective over the context's communicator (shard.cu)
        build_boxes_sliced(c, min_distance);
    else
        build_boxes(c, min_distance);
    // (Measured and dropped: running the two lists on two streams so that the tail rounds of
    // one overlap the bulk of the other.  The earliest toi is established late -- in the tail
    // rounds of the vertex-face pass -- so an edge-edge pass that starts before it is final
    // prunes less: +44 % box checks on config 2, +11 % on config 4, no net gain.)
    if (!ipc && !want_collisions) {
        // Plain ccd().  Two streams, ONE order of the narrow phases:
        //   main stream : VF sweep -> VF narrow phase
        //   sort stream : EE sort -> EE sweep .......... (waits for VF narrow) -> EE narrow phase
        // The edge list's sort and sweep run under the vertex-face work (they depend on
        // nothing but the boxes); its narrow phase starts after the vertex-face one, so it
        // prunes with the same earliest toi as in a sequential run.  The narrow phase of a
        // list is only enqueued; the host picks its counters up later and the toi travels
        // back with them.
        double* d_gtoi = gtoi_set(c, 1.0, c->stream); // ccd.cu:125
        bool any_batch = false;
        c->cur = &c->runs[1];
        broad_phase_begin_enqueue(c, SCCD_EE); // stream order: after the edge sort
        for (int kind = 0; kind < 2; kind++) {
            c->cur = &c->runs[kind];
            if (kind == SCCD_VF)
                broad_phase_begin(c, kind);
            else
                broad_phase_begin_finish(c);
            if (broad_phase_complete(c)) {
                rrecord(c, kind == SCCD_VF ? EV_SW0B : EV_SW1B);
                rrecord(c, kind == SCCD_VF ? EV_NP0A : EV_NP1A);
            }
            bool first = true;
            while (!broad_phase_complete(c)) {
                narrow_finish(c, d_gtoi); // chunked lists: one batch in flight at a time
                const sccd_pair* d_pairs = nullptr;
                int64_t n = 0;
                broad_phase_partial(c, &d_pairs, &n);
                if (first)
                    rrecord(c, kind == SCCD_VF ? EV_NP0A : EV_NP1A);
                // The edge-edge SOLVER starts after the vertex-face narrow phase, so that it
                // prunes with its earliest toi; the edge-edge cull and the sort of its survivors
                // need no bound and run before that wait, under the vertex-face work.
                narrow_enqueue(
                    c, kind, mesh_input(c, d_pairs, n), min_distance, max_iter, tol,
                    allow_zero_toi, d_gtoi, nullptr,
                    (first && kind == SCCD_EE && !c->opt.concurrent_passes) ? c->ev_vf_done : nullptr);
                first = false;
                any_batch = any_batch || c->cur->pending.active;
            }
            rrecord(c, kind == SCCD_VF ? EV_NP0B : EV_NP1B);
            if (kind == SCCD_VF)
                SCCD_CUDA(cudaEventRecord(c->ev_vf_done, c->stream));
        }
        for (int kind = 0; kind < 2; kind++) {
            c->cur = &c->runs[kind];
            if (!narrow_finish(c, d_gtoi))
                host_sync(c, c->cur->stream);
        }
        c->cur = &c->runs[0];
        if (sharded) {
            // every batch of both lists is complete (extra rounds included): ONE all-reduce(min)
            // of the earliest toi.  The edge-edge pass pruned with this rank's own vertex-face
            // bound, which changes no result (the minimum is order-independent).
            allreduce_min_toi(c, d_gtoi, c->stream);
            SCCD_CUDA(cudaMemcpyAsync(&c->h_gtoi[2], d_gtoi, 8, cudaMemcpyDeviceToHost, c->stream));
            record(c, EV_T1);
            host_sync(c, c->stream);
            c->stats.n_host_syncs++;
            finish_stats(c, true);
            *toi_out = std::min(1.0, c->h_gtoi[2]);
            return;
        }
        record(c, EV_T1);
        finish_stats(c, true);
        // the toi travelled back with the counters of the last batch (1.0 if there was none)
        *toi_out = any_batch ? std::min(1.0, c->h_gtoi[0]) : 1.0;
        return;
    }
    double toi = 1.0; // ccd.cu:125
    for (int kind = 0; kind < 2; kind++) {
        broad_phase_begin(c, kind);
        bool first = true;
        if (broad_phase_complete(c)) {
            record(c, kind == SCCD_VF ? EV_SW0B : EV_SW1B);
            record(c, kind == SCCD_VF ? EV_NP0A : EV_NP1A);
        }
        while (!broad_phase_complete(c)) {
            const sccd_pair* d_pairs = nullptr;
            int64_t n = 0;
            broad_phase_partial(c, &d_pairs, &n);
            if (first)
                record(c, kind == SCCD_VF ? EV_NP0A : EV_NP1A);
            first = false;
            const NarrowInput in = mesh_input(c, d_pairs, n);
            double* d_tq = nullptr;
            if (want_collisions && n > 0)
                d_tq = (double*)c->cur->b_toi_q.reserve((size_t)n * 8);
            if (!ipc) {
                narrow_run(c, kind, in, min_distance, max_iter, tol, allow_zero_toi, &toi, d_tq);
            } else {
                // ipc_ccd_strategy.cu:54-92
                const double before = toi;
                narrow_run(c, kind, in, min_distance, max_iter, tol, true, &toi, nullptr);
                if (toi < 1e-6) {
                    toi = before;
                    narrow_run(c, kind, in, 0.0, -1, tol, false, &toi, nullptr);
                    toi *= 0.8;
                    if (c->f32) // earliest_toi is a float there (ipc_ccd_strategy.cu:88)
                        toi = (double)(float)toi;
                }
            }
            if (d_tq) {
                // copy_out_collisions (narrow_phase.cu:84-103)
                auto& R = *c->cur;
                small_scratch(c);
                unsigned long long* d_cnt =
                    reinterpret_cast<unsigned long long*>(R.b_small.as<char>() + 192);
                SCCD_CUDA(cudaMemsetAsync(d_cnt, 0, 8, c->stream));
                sccd_pair* d_ids = (sccd_pair*)R.b_queries.reserve((size_t)n * 16);
                double* d_t = reinterpret_cast<double*>(d_ids + n);
                launch_compact_collisions(d_pairs, d_tq, n, d_ids, d_t, d_cnt, c->stream, c->lc);
                unsigned long long* h = reinterpret_cast<unsigned long long*>(R.h_small);
                SCCD_CUDA(cudaMemcpyAsync(&h[2], d_cnt, 8, cudaMemcpyDeviceToHost, c->stream));
                host_sync(c, c->stream);
                const size_t k = (size_t)h[2], old = coll_ids->size();
                coll_ids->resize(old + k);
                coll_toi->resize(old + k);
                if (k) {
                    SCCD_CUDA(cudaMemcpyAsync(
                        coll_ids->data() + old, d_ids, k * sizeof(sccd_pair),
                        cudaMemcpyDeviceToHost, c->stream));
                    SCCD_CUDA(cudaMemcpyAsync(
                        coll_toi->data() + old, d_t, k * 8, cudaMemcpyDeviceToHost, c->stream));
                    host_sync(c, c->stream);
                }
                n_coll[kind] += (int64_t)k;
            }
        }
        record(c, kind == SCCD_VF ? EV_NP0B : EV_NP1B);
    }
    record(c, EV_T1);
    finish_stats(c, true);
    *toi_out = toi;
}

} // namespace host
} // namespace sccd
using namespace sccd::host;

// =================================================================================== C ABI
extern "C" {

const char* sccd_version(void) { return "0.2.0 sm_100a"; }
size_t sccd_stats_size(void) { return sizeof(sccd_stats); }

int sccd_create(int device, void* stream, sccd_ctx** out)
{
    if (!out)
        return SCCD_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        (void)cudaGetLastError();
        return SCCD_ERR_CUDA; // no CPU fallback
    }
    sccd_ctx* c = new (std::nothrow) sccd_ctx();
    if (!c)
        return SCCD_ERR_MEMORY;
    c->device = device;
    c->stream = (cudaStream_t)stream;
    const int rc = guarded(c, [&] {
        cudaDeviceProp prop;
        SCCD_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw CudaError("sccd: this library is built for sm_100a (B200) only");
        c->num_sms = prop.multiProcessorCount;
        // initial option values; afterwards only sccd_set_option changes them
        if (const char* e = getenv("SCCD_GRID_SCALE"))
            c->grid_scale = std::min(64.0, std::max(0.25, atof(e)));
        if (const char* e = getenv("SCCD_GRID_REPL"))
            c->grid_repl = std::min(16.0, std::max(1.0, atof(e)));
        if (const char* e = getenv("SCCD_NP_FLAGS"))
            c->opt.np_flags = (int)strtoll(e, nullptr, 0);
        if (const char* e = getenv("SCCD_NP_FLAGS_EE"))
            c->opt.np_flags_ee = (int)strtoll(e, nullptr, 0);
        if (const char* e = getenv("SCCD_NP_DEPTH"))
            c->opt.np_depth = std::min(128, std::max(2, atoi(e)));
        if (const char* e = getenv("SCCD_NP_CULL"))
            c->opt.np_cull = std::max(0, std::min(3, atoi(e)));
        if (const char* e = getenv("SCCD_KEY_STEPS"))
            c->opt.key_steps = std::min(16, std::max(0, atoi(e)));
        if (const char* e = getenv("SCCD_SWEEP_STAGED"))
            c->opt.sweep_staged = atoi(e) != 0;
        if (const char* e = getenv("SCCD_CONCURRENT_PASSES"))
            c->opt.concurrent_passes = atoi(e) != 0;
        if (const char* e = getenv("SCCD_NP_SOLVER"))
            c->opt.np_solver = (atoi(e) == 1 || atoi(e) == 4 || atoi(e) == 8) ? atoi(e) : 0;
        if (const char* e = getenv("SCCD_SWEEP_AXIS"))
            c->opt.sweep_axis = std::min(2, std::max(-1, atoi(e)));
        narrow_init_device();
        for (auto& e : c->ev)
            SCCD_CUDA(cudaEventCreate(&e));
        c->runs[0].stream = c->stream;
        SCCD_CUDA(cudaStreamCreateWithFlags(&c->sort_stream, cudaStreamNonBlocking));
        SCCD_CUDA(cudaEventCreateWithFlags(&c->ev_counts, cudaEventDisableTiming));
        SCCD_CUDA(cudaEventCreateWithFlags(&c->ev_sorted1, cudaEventDisableTiming));
        SCCD_CUDA(cudaEventCreateWithFlags(&c->ev_vf_done, cudaEventDisableTiming));
        SCCD_CUDA(cudaEventCreateWithFlags(&c->ev_boxes, cudaEventDisableTiming));
        SCCD_CUDA(cudaEventCreateWithFlags(&c->ev_cnt1, cudaEventDisableTiming));
        SCCD_CUDA(cudaEventCreateWithFlags(&c->ev_stats, cudaEventDisableTiming));
        if (const char* e = getenv("SCCD_NP_TAIL")) // (experiments)
            c->opt.np_tail_lanes = std::max(0, std::min(32, atoi(e)));
        if (const char* e = getenv("SCCD_QUEUE_CTAS")) { // "vf,ee" (experiments)
            int a = 0, b = 0;
            if (sscanf(e, "%d,%d", &a, &b) >= 1)
                c->opt.queue_ctas[0] = a, c->opt.queue_ctas[1] = b;
        }
        if (const char* e = getenv("SCCD_REUSE_GRID"))
            c->opt.reuse_grid = atoi(e) != 0;
        c->runs[1].stream = c->sort_stream;
        return SCCD_OK;
    });
    if (rc != SCCD_OK) {
        delete c;
        return rc;
    }
    *out = c;
    return SCCD_OK;
}

void sccd_destroy(sccd_ctx* ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->sort_stream)
        cudaStreamSynchronize(ctx->sort_stream);
    comm_destroy(ctx);
    delete ctx;
}

const char* sccd_last_error(const sccd_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int sccd_set_memory_limit(sccd_ctx* ctx, size_t bytes)
{
    if (!ctx)
        return SCCD_ERR_ARG;
    ctx->memory_limit = bytes;
    return SCCD_OK;
}

int sccd_set_max_pairs_per_chunk(sccd_ctx* ctx, int64_t max_pairs)
{
    if (!ctx || max_pairs < 0)
        return SCCD_ERR_ARG;
    ctx->max_pairs_per_chunk = max_pairs;
    return SCCD_OK;
}

int sccd_set_queue_capacity(sccd_ctx* ctx, int64_t items)
{
    if (!ctx || items < 0)
        return SCCD_ERR_ARG;
    ctx->queue_cap = items;
    return SCCD_OK;
}

int sccd_set_grid_cells(sccd_ctx* ctx, int max_cells)
{
    if (!ctx || max_cells < 0)
        return SCCD_ERR_ARG;
    ctx->grid_max_cells = max_cells == 0 ? -1 : max_cells;
    return SCCD_OK;
}

int sccd_set_scalar_type(sccd_ctx* ctx, int type)
{
    if (!ctx || (type != SCCD_F64 && type != SCCD_F32))
        return SCCD_ERR_ARG;
    if (ctx->f32 != (type == SCCD_F32)) {
        ctx->f32 = type == SCCD_F32;
        ctx->have_boxes = false; // boxes and the vertex table depend on the scalar type
        ctx->runs[0].bp_kind = ctx->runs[1].bp_kind = -1;
    }
    return SCCD_OK;
}

int sccd_set_option(sccd_ctx* ctx, int option, int64_t value)
{
    if (!ctx)
        return SCCD_ERR_ARG;
    auto& o = ctx->opt;
    switch (option) {
    case SCCD_OPT_NARROW_CULL:
        if (value < 0 || value > 3)
            return SCCD_ERR_ARG;
        o.np_cull = (int)value;
        break;
    case SCCD_OPT_NARROW_FLAGS: o.np_flags = (int)value; break;
    case SCCD_OPT_NARROW_FLAGS_EE: o.np_flags_ee = (int)value; break;
    case SCCD_OPT_NARROW_MAX_DEPTH:
        if (value < 2 || value > 128)
            return SCCD_ERR_ARG;
        o.np_depth = (int)value;
        break;
    case SCCD_OPT_MAX_ITER_MODE:
        if (value != 0 && value != 1)
            return SCCD_ERR_ARG;
        o.cap_drops = (int)value;
        break;
    case SCCD_OPT_KEY_STEPS:
        if (value < 0 || value > 16)
            return SCCD_ERR_ARG;
        o.key_steps = (int)value;
        break;
    case SCCD_OPT_GRID_SCALE_MILLI:
        if (value < 250 || value > 64000)
            return SCCD_ERR_ARG;
        ctx->grid_scale = (double)value / 1000.0;
        break;
    case SCCD_OPT_GRID_REPL_MILLI:
        if (value < 1000 || value > 16000)
            return SCCD_ERR_ARG;
        ctx->grid_repl = (double)value / 1000.0;
        break;
    case SCCD_OPT_PROFILE: {
        if (value < 0 || value > 2)
            return SCCD_ERR_ARG;
        // 2: everything on the caller's stream, so that a kernel's event pair brackets that
        // kernel alone (with two streams the pairs of one list include the time its kernels
        // wait behind the other list's)
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess
            || cudaStreamSynchronize(ctx->sort_stream) != cudaSuccess)
            return SCCD_ERR_CUDA;
        if (!ctx->sort_stream_own)
            ctx->sort_stream_own = ctx->sort_stream;
        ctx->sort_stream = value == 2 ? ctx->stream : ctx->sort_stream_own;
        ctx->runs[1].stream = ctx->sort_stream;
        o.profile = (int)value;
        break;
    }
    case SCCD_OPT_CONCURRENT_PASSES: o.concurrent_passes = value != 0; break;
    case SCCD_OPT_SWEEP_STAGED: o.sweep_staged = value != 0; break;
    case SCCD_OPT_REUSE_GRID: o.reuse_grid = value != 0; break;
    case SCCD_OPT_NARROW_SOLVER:
        if (value != 0 && value != 1 && value != 4 && value != 8)
            return SCCD_ERR_ARG;
        o.np_solver = (int)value;
        break;
    case SCCD_OPT_SWEEP_AXIS:
        if (value < -1 || value > 2)
            return SCCD_ERR_ARG;
        if (o.sweep_axis != (int)value) {
            o.sweep_axis = (int)value;
            ctx->have_boxes = false;
            ctx->runs[0].bp_kind = ctx->runs[1].bp_kind = -1;
        }
        break;
    default:
        ctx->error = "set_option: unknown option";
        return SCCD_ERR_ARG;
    }
    return SCCD_OK;
}

int sccd_get_option(const sccd_ctx* ctx, int option, int64_t* value)
{
    if (!ctx || !value)
        return SCCD_ERR_ARG;
    const auto& o = ctx->opt;
    switch (option) {
    case SCCD_OPT_NARROW_CULL: *value = o.np_cull; break;
    case SCCD_OPT_NARROW_FLAGS: *value = o.np_flags; break;
    case SCCD_OPT_NARROW_FLAGS_EE: *value = o.np_flags_ee; break;
    case SCCD_OPT_NARROW_MAX_DEPTH: *value = o.np_depth; break;
    case SCCD_OPT_MAX_ITER_MODE: *value = o.cap_drops; break;
    case SCCD_OPT_KEY_STEPS: *value = o.key_steps; break;
    case SCCD_OPT_GRID_SCALE_MILLI: *value = (int64_t)std::llround(ctx->grid_scale * 1000.0); break;
    case SCCD_OPT_GRID_REPL_MILLI: *value = (int64_t)std::llround(ctx->grid_repl * 1000.0); break;
    case SCCD_OPT_SWEEP_AXIS: *value = o.sweep_axis; break;
    case SCCD_OPT_PROFILE: *value = o.profile; break;
    case SCCD_OPT_NARROW_SOLVER: *value = o.np_solver; break;
    case SCCD_OPT_CONCURRENT_PASSES: *value = o.concurrent_passes; break;
    case SCCD_OPT_SWEEP_STAGED: *value = o.sweep_staged; break;
    case SCCD_OPT_REUSE_GRID: *value = o.reuse_grid; break;
    default: return SCCD_ERR_ARG;
    }
    return SCCD_OK;
}

int sccd_set_shard(sccd_ctx* ctx, int rank, int world)
{
    if (!ctx || world < 1 || rank < 0 || rank >= world || world > 16)
        return SCCD_ERR_ARG;
    ctx->rank = rank;
    ctx->world = world;
    ctx->runs[0].bp_kind = ctx->runs[1].bp_kind = -1;
    return SCCD_OK;
}

int sccd_upload_mesh(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, int on_device)
{
    return guarded(ctx, [&] {
        upload_mesh(ctx, V0, V1, nV, E, nE, F, nF, on_device != 0);
        return SCCD_OK;
    });
}

int sccd_update_vertices(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, int on_device)
{
    return guarded(ctx, [&] {
        update_vertices(ctx, V0, V1, nV, on_device != 0);
        return SCCD_OK;
    });
}

int sccd_build_boxes(sccd_ctx* ctx, double inflation_radius)
{
    return guarded(ctx, [&] {
        record(ctx, EV_T0);
        build_boxes(ctx, inflation_radius);
        finish_stats(ctx, false);
        return SCCD_OK;
    });
}

int sccd_build_vertex_boxes(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, double inflation_radius,
    sccd_aabb* out)
{
    return guarded(ctx, [&] {
        if (nV < 0 || nV >= (1ll << 31) || (nV && (!V0 || !out)))
            throw std::invalid_argument("build_vertex_boxes: bad argument");
        if (nV == 0)
            return SCCD_OK;
        const size_t vb = sizeof(double) * 3 * (size_t)nV;
        double* d0 = (double*)ctx->b_io[0].reserve(vb);
        double* d1 = d0; // single frame: the box of a point (aabb.cuh:150-153)
        SCCD_CUDA(cudaMemcpyAsync(d0, V0, vb, cudaMemcpyHostToDevice, ctx->stream));
        if (V1) {
            d1 = (double*)ctx->b_io[1].reserve(vb);
            SCCD_CUDA(cudaMemcpyAsync(d1, V1, vb, cudaMemcpyHostToDevice, ctx->stream));
        }
        sccd_aabb* d_out = (sccd_aabb*)ctx->b_io[2].reserve(sizeof(sccd_aabb) * (size_t)nV);
        const double radius_up = ctx->f32
            ? (double)std::nextafterf((float)inflation_radius, FLT_MAX)
            : std::nextafter(inflation_radius, DBL_MAX);
        launch_vertex_aabbs(d0, d1, (int)nV, radius_up, ctx->f32, d_out, ctx->stream, ctx->lc);
        SCCD_CUDA(cudaMemcpyAsync(
            out, d_out, sizeof(sccd_aabb) * (size_t)nV, cudaMemcpyDeviceToHost, ctx->stream));
        SCCD_CUDA(cudaStreamSynchronize(ctx->stream));
        return SCCD_OK;
    });
}

int sccd_build_element_boxes(
    sccd_ctx* ctx, const sccd_aabb* vertex_boxes, int64_t nV, const int32_t* idx, int64_t n,
    int verts_per_element, sccd_aabb* out)
{
    return guarded(ctx, [&] {
        if (nV < 0 || n < 0 || nV >= (1ll << 31) || n >= (1ll << 31)
            || (verts_per_element != 2 && verts_per_element != 3)
            || (n && (!idx || !out || !vertex_boxes)))
            throw std::invalid_argument("build_element_boxes: bad argument");
        if (n == 0)
            return SCCD_OK;
        const size_t ib = sizeof(int32_t) * (size_t)verts_per_element * (size_t)n;
        sccd_aabb* d_vb = (sccd_aabb*)ctx->b_io[0].reserve(sizeof(sccd_aabb) * (size_t)std::max<int64_t>(nV, 1));
        int32_t* d_idx = (int32_t*)ctx->b_io[1].reserve(ib + 16);
        sccd_aabb* d_out = (sccd_aabb*)ctx->b_io[2].reserve(sizeof(sccd_aabb) * (size_t)n);
        int* d_bad = reinterpret_cast<int*>(reinterpret_cast<char*>(d_idx) + ((ib + 3) & ~(size_t)3));
        SCCD_CUDA(cudaMemcpyAsync(
            d_vb, vertex_boxes, sizeof(sccd_aabb) * (size_t)nV, cudaMemcpyHostToDevice, ctx->stream));
        SCCD_CUDA(cudaMemcpyAsync(d_idx, idx, ib, cudaMemcpyHostToDevice, ctx->stream));
        SCCD_CUDA(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
        launch_element_aabbs(
            d_vb, (int)nV, d_idx, (int)n, verts_per_element, d_out, d_bad, ctx->stream, ctx->lc);
        int bad = 0;
        SCCD_CUDA(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
        SCCD_CUDA(cudaMemcpyAsync(
            out, d_out, sizeof(sccd_aabb) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        SCCD_CUDA(cudaStreamSynchronize(ctx->stream));
        if (bad)
            throw std::invalid_argument("build_element_boxes: vertex index out of range");
        return SCCD_OK;
    });
}

int sccd_get_boxes(sccd_ctx* ctx, int which, sccd_aabb* out)
{
    return guarded(ctx, [&] {
        if (!ctx->have_boxes)
            throw std::logic_error("get_boxes: call sccd_build_boxes first");
        if (ctx->sliced)
            throw std::logic_error("get_boxes: after sccd_ccd_sharded every rank holds a slice only");
        if (which < 0 || which > 2 || !out)
            throw std::invalid_argument("get_boxes: bad argument");
        const int n = which == 0 ? ctx->nV : (which == 1 ? ctx->nE : ctx->nF);
        const auto& U = ctx->lists[which == 1 ? 1 : 0].unsorted;
        const size_t off = which == 2 ? (size_t)ctx->nV : 0;
        std::vector<double2> x(n);
        std::vector<double4> yz(n);
        std::vector<int4> id(n);
        SCCD_CUDA(cudaStreamSynchronize(ctx->stream));
        SCCD_CUDA(cudaMemcpy(x.data(), U.x + off, sizeof(double2) * n, cudaMemcpyDeviceToHost));
        SCCD_CUDA(cudaMemcpy(yz.data(), U.yz + off, sizeof(double4) * n, cudaMemcpyDeviceToHost));
        SCCD_CUDA(cudaMemcpy(id.data(), U.id + off, sizeof(int4) * n, cudaMemcpyDeviceToHost));
        // records are rotated to (axis, axis + 1, axis + 2) mod 3
        const int ax = ctx->lists[which == 1 ? 1 : 0].axis, ay = (ax + 1) % 3, az = (ax + 2) % 3;
        for (int i = 0; i < n; i++) {
            out[i].min[ax] = x[i].x;
            out[i].max[ax] = x[i].y;
            out[i].min[ay] = yz[i].x;
            out[i].min[az] = yz[i].y;
            out[i].max[ay] = yz[i].z;
            out[i].max[az] = yz[i].w;
            out[i].vertex_ids[0] = id[i].x;
            out[i].vertex_ids[1] = id[i].y;
            out[i].vertex_ids[2] = id[i].z;
            // vertices carry the flipped id on the device; the reference's host boxes do not
            out[i].element_id = which == 0 ? -id[i].w - 1 : id[i].w;
        }
        return SCCD_OK;
    });
}

int sccd_set_boxes(
    sccd_ctx* ctx, const sccd_aabb* a, int64_t na, const sccd_aabb* b, int64_t nb,
    int sort_axis, int* next_axis)
{
    return guarded(ctx, [&] {
        set_boxes(ctx, a, na, b, nb, sort_axis, next_axis);
        return SCCD_OK;
    });
}

int sccd_broad_phase_begin(sccd_ctx* ctx, int kind)
{
    return guarded(ctx, [&] {
        broad_phase_begin(ctx, kind);
        return SCCD_OK;
    });
}

int sccd_broad_phase_partial(sccd_ctx* ctx, const sccd_pair** d_pairs, int64_t* n_pairs)
{
    return guarded(ctx, [&] {
        if (!d_pairs || !n_pairs)
            throw std::invalid_argument("broad_phase_partial: null output");
        broad_phase_partial(ctx, d_pairs, n_pairs);
        return SCCD_OK;
    });
}

int sccd_broad_phase_is_complete(sccd_ctx* ctx)
{
    if (!ctx)
        return SCCD_ERR_ARG;
    if (ctx->cur->bp_kind < 0) {
        ctx->error = "Must initialize build broad phase before detecting overlaps!";
        return SCCD_ERR_STATE;
    }
    return broad_phase_complete(ctx) ? 1 : 0;
}

int sccd_broad_phase(sccd_ctx* ctx, int kind, sccd_pair* out, int64_t cap, int64_t* n_total)
{
    return guarded(ctx, [&] {
        broad_phase_begin(ctx, kind);
        int64_t written = 0, total = 0;
        while (!broad_phase_complete(ctx)) {
            const sccd_pair* d = nullptr;
            int64_t n = 0;
            broad_phase_partial(ctx, &d, &n);
            if (out && written < cap && n > 0) {
                const int64_t k = std::min(n, cap - written);
                SCCD_CUDA(cudaMemcpyAsync(
                    out + written, d, (size_t)k * sizeof(sccd_pair), cudaMemcpyDeviceToHost,
                    ctx->stream));
                SCCD_CUDA(cudaStreamSynchronize(ctx->stream));
                written += k;
            }
            total += n;
        }
        if (n_total)
            *n_total = total;
        finish_stats(ctx, false);
        return SCCD_OK;
    });
}

int sccd_narrow_phase(
    sccd_ctx* ctx, int kind, const sccd_pair* d_pairs, int64_t n, double ms, int max_iter,
    double tol, int allow_zero_toi, double* toi_inout, double* d_toi_per_query)
{
    return guarded(ctx, [&] {
        if (!toi_inout || (n > 0 && !d_pairs) || n < 0 || (kind != SCCD_VF && kind != SCCD_EE))
            throw std::invalid_argument("narrow_phase: bad argument");
        record(ctx, EV_TMPA);
        narrow_run(
            ctx, kind, mesh_input(ctx, d_pairs, n), ms, max_iter, tol, allow_zero_toi != 0,
            toi_inout, d_toi_per_query);
        finish_stats(ctx, false);
        return SCCD_OK;
    });
}

int sccd_narrow_phase_queries(
    sccd_ctx* ctx, int kind, const double* queries, int64_t n, int on_device, double ms,
    int max_iter, double tol, int allow_zero_toi, double* toi_inout, double* d_toi_per_query)
{
    return guarded(ctx, [&] {
        if (!toi_inout || (n > 0 && !queries) || n < 0 || (kind != SCCD_VF && kind != SCCD_EE))
            throw std::invalid_argument("narrow_phase_queries: bad argument");
        NarrowInput in;
        in.n = n;
        if (on_device || n == 0) {
            in.queries = queries;
        } else {
            double* d = (double*)ctx->cur->b_queries.reserve((size_t)n * 24 * 8);
            SCCD_CUDA(cudaMemcpyAsync(
                d, queries, (size_t)n * 24 * 8, cudaMemcpyHostToDevice, ctx->stream));
            in.queries = d;
        }
        record(ctx, EV_TMPA);
        narrow_run(ctx, kind, in, ms, max_iter, tol, allow_zero_toi != 0, toi_inout, d_toi_per_query);
        record(ctx, EV_TMPB);
        SCCD_CUDA(cudaEventSynchronize(ctx->ev[EV_TMPB]));
        ctx->stats.ms_narrow[kind] = elapsed(ctx, EV_TMPA, EV_TMPB);
        finish_stats(ctx, false);
        return SCCD_OK;
    });
}

int sccd_narrow_phase_checks(sccd_ctx* ctx, const uint32_t** d_checks, int64_t* n)
{
    if (!ctx || !d_checks || !n)
        return SCCD_ERR_ARG;
    const auto& R = ctx->runs[0];
    *d_checks = R.checks_n > 0 ? R.b_checks_q.as<uint32_t>() : nullptr;
    *n = R.checks_n;
    return SCCD_OK;
}

int sccd_ccd(
    sccd_ctx* ctx, double min_distance, int max_iter, double tol, int allow_zero_toi, double* toi)
{
    return guarded(ctx, [&] {
        if (!toi)
            throw std::invalid_argument("ccd: null toi");
        run_pipeline(
            ctx, min_distance, max_iter, tol, allow_zero_toi != 0, false, toi, false, nullptr,
            nullptr, nullptr);
        return SCCD_OK;
    });
}

int sccd_ccd_collisions(
    sccd_ctx* ctx, double min_distance, int max_iter, double tol, int allow_zero_toi,
    double* toi, sccd_pair* ids, double* tois, int64_t cap, int64_t* n_vf, int64_t* n_ee)
{
    return guarded(ctx, [&] {
        if (!toi)
            throw std::invalid_argument("ccd_collisions: null toi");
        std::vector<sccd_pair>& cid = ctx->coll_ids;
        std::vector<double>& ct = ctx->coll_toi;
        int64_t* nc = ctx->coll_n;
        cid.clear();
        ct.clear();
        nc[0] = nc[1] = 0;
        run_pipeline(
            ctx, min_distance, max_iter, tol, allow_zero_toi != 0, false, toi, true, &cid, &ct, nc);
        const int64_t k = std::min<int64_t>(cap, (int64_t)cid.size());
        if (ids && k > 0)
            std::memcpy(ids, cid.data(), (size_t)k * sizeof(sccd_pair));
        if (tois && k > 0)
            std::memcpy(tois, ct.data(), (size_t)k * 8);
        if (n_vf)
            *n_vf = nc[0];
        if (n_ee)
            *n_ee = nc[1];
        return SCCD_OK;
    });
}

int sccd_get_collisions(
    sccd_ctx* ctx, sccd_pair* ids, double* tois, int64_t cap, int64_t* n_vf, int64_t* n_ee)
{
    if (!ctx || cap < 0)
        return SCCD_ERR_ARG;
    const int64_t k = std::min<int64_t>(cap, (int64_t)ctx->coll_ids.size());
    if (ids && k > 0)
        std::memcpy(ids, ctx->coll_ids.data(), (size_t)k * sizeof(sccd_pair));
    if (tois && k > 0)
        std::memcpy(tois, ctx->coll_toi.data(), (size_t)k * 8);
    if (n_vf)
        *n_vf = ctx->coll_n[0];
    if (n_ee)
        *n_ee = ctx->coll_n[1];
    return SCCD_OK;
}

int sccd_ccd_host(
    sccd_ctx* ctx, const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, double min_distance, int max_iter, double tol,
    int allow_zero_toi, double* toi)
{
    return guarded(ctx, [&] {
        if (!toi)
            throw std::invalid_argument("ccd: null toi");
        upload_mesh(ctx, V0, V1, nV, E, nE, F, nF, false);
        run_pipeline(
            ctx, min_distance, max_iter, tol, allow_zero_toi != 0, false, toi, false, nullptr,
            nullptr, nullptr);
        return SCCD_OK;
    });
}

int sccd_ipc_ccd_strategy(
    sccd_ctx* ctx, double min_distance, int max_iter, double tol, double* toi)
{
    return guarded(ctx, [&] {
        if (!toi)
            throw std::invalid_argument("ipc_ccd_strategy: null toi");
        run_pipeline(
            ctx, min_distance, max_iter, tol, true, true, toi, false, nullptr, nullptr, nullptr);
        return SCCD_OK;
    });
}

int sccd_get_stats(const sccd_ctx* ctx, sccd_stats* out)
{
    if (!ctx || !out)
        return SCCD_ERR_ARG;
    *out = ctx->stats;
    return SCCD_OK;
}

int sccd_reset_stats(sccd_ctx* ctx)
{
    return guarded(ctx, [&] {
        kt_resolve(ctx);
        reset_stats(ctx);
        return SCCD_OK;
    });
}

int sccd_measure_fp64_peak(sccd_ctx* ctx, double* dfma_per_second)
{
    return guarded(ctx, [&] {
        if (!dfma_per_second)
            throw std::invalid_argument("measure_fp64_peak: null output");
        SCCD_CUDA(cudaStreamSynchronize(ctx->stream));
        *dfma_per_second = measure_dfma_per_second(ctx->num_sms, ctx->stream);
        return SCCD_OK;
    });
}

int sccd_synchronize(sccd_ctx* ctx)
{
    return guarded(ctx, [&] {
        SCCD_CUDA(cudaStreamSynchronize(ctx->stream));
        SCCD_CUDA(cudaStreamSynchronize(ctx->sort_stream));
        return SCCD_OK;
    });
}

} // extern "C"
