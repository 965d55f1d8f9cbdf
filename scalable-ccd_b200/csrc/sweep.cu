// Sweep over the (cell, major-axis)-sorted records (north-star item 2b).
//
// Replaces sweep_and_tiniest_queue (cuda/broad_phase/sweep.cu:101-182: one warp per CTA,
// a 64-slot shared ring rebalanced with shared atomics every round, uncoalesced 48 B
// MiniBox loads per candidate, one global atomicAdd per emitted pair into a buffer that is
// value-initialised on every call) with a tiled sweep:
//
//   * a CTA owns a tile of 256 consecutive sorted records ("owners", one per lane);
//   * the candidate window to the right of the tile is streamed through shared memory in
//     256-record chunks of the 24-byte prefilter view (register double-buffered), every
//     lane testing its owner against each staged candidate by shared-memory broadcast:
//     one 64-bit compare key_j <= reach_i (same cell AND f32 xmin_j <= f32 xmax_i, see
//     common.cuh) plus four f32 compares on y / z;
//   * prefilter survivors (a conservative superset: min rounded down / max rounded up to
//     f32) are ballot/scan-compacted into a per-warp shared queue -- the "tiniest queue" --
//     and drained 32 at a time with ALL lanes running the exact double test + id tests,
//     so the rare expensive path never diverges;
//   * output is count -> exclusive scan -> fill: the pair list is exactly sized, ordered by
//     (owner position, candidate position) and therefore deterministic; no global atomics
//     and no giant memset; 64-bit offsets.
//
// The emitted SET equals the reference's: every pair with closed overlap on x, y, z
// (cuda/broad_phase/aabb.cuh:100-104, sweep.cu:131,173), valid list membership
// (collision.cuh:27-35) and no shared vertex (collision.cuh:17-21).  Why nothing is lost or
// duplicated: (1) sorting on the f32 key instead of the double only changes the ORDER in
// which ties are visited -- the window test on rounded values is a superset of
// min_j <= max_i and the exact test checks both x directions; (2) two overlapping boxes both
// own a record in the cell of (max(ymin_a,ymin_b), max(zmin_a,zmin_b)) because cell_index()
// is monotone, and the pair is accepted in that cell only.
#include "common.cuh"

#include <cfloat>

namespace sccd {

namespace {

constexpr int kTile = 256;  // owners per CTA == threads per CTA
constexpr int kWarps = kTile / 32;
constexpr int kChunk = 256; // candidates staged per step
constexpr int kQueueCap = 32 * 32 + 32;
constexpr int kRelBits = 27; // candidate position relative to the tile start
constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned long long kKeyInf = ~0ull;

struct SweepSmem {
    unsigned long long c_key[kChunk];
    float4 c_yz[kChunk];
    double2 o_x[kTile];
    double4 o_yz[kTile];
    int4 o_id[kTile];
    unsigned long long o_off[kTile];
    uint32_t o_cell[kTile];
    uint32_t o_cnt[kTile];
    uint32_t q[kWarps][kQueueCap];
    unsigned long long red[kWarps];
};

template <bool FILL, bool TWO_LISTS>
__device__ __forceinline__ void drain32(
    SweepSmem& sm, const BoxArrays& box, const GridParams& g, int tile0, int warp, int lane,
    uint32_t entry, bool active, sccd_pair* __restrict__ pairs)
{
    const int ol = active ? (int)(entry >> kRelBits) : 0;
    const int t = warp * 32 + ol;
    const int j = tile0 + (int)(entry & ((1u << kRelBits) - 1));
    bool hit = false;
    int ea = 0, eb = 0;
    if (active) {
        const double2 ax = sm.o_x[t];
        const double4 ayz = sm.o_yz[t];
        const int4 aid = sm.o_id[t];
        const double2 bx = __ldg(&box.x[j]);
        const double4 byz = ldg_d4(&box.yz[j]);
        const int4 bid = __ldg(&box.id[j]);
        // closed-interval overlap on all three axes (aabb.cuh:67-72 / 100-104)
        hit = ax.y >= bx.x && ax.x <= bx.y && ayz.z >= byz.x && ayz.x <= byz.z
            && ayz.w >= byz.y && ayz.y <= byz.w;
        if (TWO_LISTS) // exactly one of the two comes from list A (collision.cuh:27-35)
            hit = hit && ((aid.w ^ bid.w) < 0);
        // collision.cuh:17-21
        const bool share = aid.x == bid.x || aid.x == bid.y || aid.x == bid.z
            || aid.y == bid.x || aid.y == bid.y || aid.y == bid.z || aid.z == bid.x
            || aid.z == bid.y || aid.z == bid.z;
        hit = hit && !share;
        if (g.sy * g.sz > 1) {
            // report the pair only in its home cell (both boxes own a record there)
            const int cy = cell_index(fmax(ayz.x, byz.x), g.y0, g.inv_hy, g.sy);
            const int cz = cell_index(fmax(ayz.y, byz.y), g.z0, g.inv_hz, g.sz);
            hit = hit && (uint32_t)(cy * g.sz + cz) == sm.o_cell[t];
        }
        ea = aid.w;
        eb = bid.w;
    }
    // deterministic slot = rank among the lanes of this batch that hit for the same owner
    const unsigned peers = __match_any_sync(kFull, active ? ol : 32 + lane);
    const unsigned hits = __ballot_sync(kFull, hit);
    const unsigned mine = peers & hits;
    const int tot = __popc(mine);
    const bool leader = active && (lane == __ffs(peers) - 1);
    if (FILL) {
        const uint32_t cur = active ? sm.o_cnt[t] : 0;
        __syncwarp();
        if (hit) {
            const int rank = __popc(mine & ((1u << lane) - 1));
            const int mn = min(ea, eb), mx = max(ea, eb);
            sccd_pair p;
            // sweep.cu:152-164: (vertex, face) for two lists, (min id, max id) otherwise
            p.a = TWO_LISTS ? (-mn - 1) : mn;
            p.b = mx;
            pairs[sm.o_off[t] + cur + rank] = p;
        }
        if (leader && tot)
            sm.o_cnt[t] = cur + tot;
        __syncwarp();
    } else {
        if (leader && tot)
            sm.o_cnt[t] += tot;
        __syncwarp();
    }
}

template <bool FILL, bool TWO_LISTS>
__global__ void __launch_bounds__(kTile) sweep_kernel(
    PrefilterArrays pf, BoxArrays box, GridParams g, int n, int shard_lo, int owner_lo,
    int owner_hi, uint32_t* __restrict__ counts, const unsigned long long* __restrict__ offsets,
    sccd_pair* __restrict__ pairs, unsigned long long* __restrict__ n_candidates)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SweepSmem& sm = *reinterpret_cast<SweepSmem*>(smem_raw);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int tile0 = owner_lo + blockIdx.x * kTile;
    const int i = tile0 + tid;
    const bool valid = i < owner_hi;

    unsigned long long my_reach = 0; // an invalid owner reaches nothing (keys are > 0)
    float4 my = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        my_reach = pf.reach[i];
        my = pf.yz[i];
        sm.o_x[tid] = box.x[i];
        sm.o_yz[tid] = box.yz[i];
        sm.o_id[tid] = box.id[i];
        sm.o_cell[tid] = (uint32_t)(my_reach >> 32);
        if (FILL) // position of this owner's first pair inside the chunk being filled
            sm.o_off[tid] = offsets[i - shard_lo] - offsets[owner_lo - shard_lo];
    }
    sm.o_cnt[tid] = 0;

    // reach of the warp / of the tile in sorted-key order
    unsigned long long wmax = my_reach;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(kFull, wmax, o);
        wmax = v > wmax ? v : wmax;
    }
    if (lane == 0)
        sm.red[warp] = wmax;
    __syncthreads();
    unsigned long long tile_max = sm.red[0];
#pragma unroll
    for (int w = 1; w < kWarps; w++)
        tile_max = sm.red[w] > tile_max ? sm.red[w] : tile_max;

    uint32_t* q = sm.q[warp];
    int qn = 0;                    // entries waiting in this warp's queue (warp-uniform)
    unsigned long long tested = 0; // exact tests run by this warp (lane 0's copy is used)

    // register double buffer for the next chunk
    int cs = tile0 + 1;
    unsigned long long nk = kKeyInf; // +inf pads the list
    float4 nyz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cs + tid < n) {
        nk = pf.key[cs + tid];
        nyz = pf.yz[cs + tid];
    }
    for (; cs < n; cs += kChunk) {
        __syncthreads(); // previous chunk fully consumed
        sm.c_key[tid] = nk;
        sm.c_yz[tid] = nyz;
        __syncthreads();
        if (sm.c_key[0] > tile_max)
            break; // block-uniform: the sorted list has left the tile's reach
        {
            const int jn = cs + kChunk + tid;
            nk = kKeyInf;
            if (jn < n) {
                nk = pf.key[jn];
                nyz = pf.yz[jn];
            }
        }
        if (sm.c_key[0] > wmax)
            continue; // warp-uniform
        for (int k0 = 0; k0 < kChunk; k0 += 32) {
            if (sm.c_key[k0] > wmax)
                break; // warp-uniform
            uint32_t mask = 0;
#pragma unroll
            for (int kk = 0; kk < 32; kk++) {
                const unsigned long long kj = sm.c_key[k0 + kk];
                const float4 b = sm.c_yz[k0 + kk];
                const bool p = (kj <= my_reach) && (b.x <= my.y) && (my.x <= b.y)
                    && (b.z <= my.w) && (my.z <= b.w);
                mask |= (p ? 1u : 0u) << kk;
            }
            // only candidates strictly after the owner in sorted order
            const int jbase = cs + k0;
            const int d = i - jbase;
            if (d >= 31)
                mask = 0;
            else if (d >= 0)
                mask &= ~((2u << d) - 1u);
            if (!__any_sync(kFull, mask != 0))
                continue;

            // ballot/scan compaction of the survivors into the warp queue
            const int cnt = __popc(mask);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o)
                    incl += v;
            }
            const int total = __shfl_sync(kFull, incl, 31);
            int pos = qn + incl - cnt;
            const uint32_t rel0 = (uint32_t)(jbase - tile0);
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                q[pos++] = ((uint32_t)lane << kRelBits) | (rel0 + b);
            }
            qn += total;
            __syncwarp();
            if (qn >= 32) {
                const int nfull = qn & ~31;
                for (int h = 0; h < nfull; h += 32)
                    drain32<FILL, TWO_LISTS>(
                        sm, box, g, tile0, warp, lane, q[h + lane], true, pairs);
                tested += nfull;
                const int r = qn - nfull;
                const uint32_t v = (lane < r) ? q[nfull + lane] : 0u;
                __syncwarp();
                if (lane < r)
                    q[lane] = v;
                qn = r;
                __syncwarp();
            }
        }
    }
    if (qn > 0) {
        drain32<FILL, TWO_LISTS>(
            sm, box, g, tile0, warp, lane, lane < qn ? q[lane] : 0u, lane < qn, pairs);
        tested += qn;
    }
    if (!FILL) {
        __syncwarp();
        if (valid)
            counts[i - shard_lo] = sm.o_cnt[tid];
        if (lane == 0 && tested && n_candidates)
            atomicAdd(n_candidates, tested);
    }
}

// window[i] = #records j > i with key_j <= reach_i (sweep work estimate used to balance
// owner ranges across GPUs).
__global__ void __launch_bounds__(256)
    sweep_window_kernel(PrefilterArrays pf, int n, uint32_t* __restrict__ window)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n)
        return;
    const unsigned long long reach = pf.reach[i];
    int lo = i + 1, hi = n; // first j in (i, n) with key[j] > reach
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pf.key[mid] <= reach)
            lo = mid + 1;
        else
            hi = mid;
    }
    window[i] = (uint32_t)(lo - i - 1);
}

__global__ void find_chunk_end_kernel(
    const unsigned long long* __restrict__ offsets, int lo, int hi,
    unsigned long long budget, int* out)
{
    // largest e in [lo, hi] with offsets[e] - offsets[lo] <= budget
    const unsigned long long base = offsets[lo];
    int a = lo, b = hi;
    while (a < b) {
        const int mid = (a + b + 1) >> 1;
        if (offsets[mid] - base <= budget)
            a = mid;
        else
            b = mid - 1;
    }
    out[0] = a;
}

template <bool FILL, bool TWO>
void launch_sweep(
    const SortedList& L, int shard_lo, int owner_lo, int owner_hi, uint32_t* counts,
    const unsigned long long* offsets, sccd_pair* pairs, unsigned long long* n_candidates,
    cudaStream_t s, LaunchCounter& lc)
{
    const int owners = owner_hi - owner_lo;
    if (owners <= 0)
        return;
    static bool configured = false;
    auto kern = sweep_kernel<FILL, TWO>;
    if (!configured) {
        SCCD_CUDA(cudaFuncSetAttribute(
            kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SweepSmem)));
        configured = true;
    }
    const int grid = (owners + kTile - 1) / kTile;
    kern<<<grid, kTile, sizeof(SweepSmem), s>>>(
        L.pf, L.box, L.grid, L.n, shard_lo, owner_lo, owner_hi, counts, offsets, pairs,
        n_candidates);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

} // namespace

void launch_sweep_windows(
    const SortedList& L, uint32_t* window, cudaStream_t s, LaunchCounter& lc)
{
    if (L.n <= 0)
        return;
    sweep_window_kernel<<<(L.n + 255) / 256, 256, 0, s>>>(L.pf, L.n, window);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_sweep_count(
    const SortedList& L, int owner_lo, int owner_hi, uint32_t* counts,
    unsigned long long* n_candidates, cudaStream_t s, LaunchCounter& lc)
{
    if (L.n >= (1 << kRelBits))
        throw std::runtime_error("sweep: more than 2^27 records in one list is not supported");
    if (L.two_lists)
        launch_sweep<false, true>(
            L, owner_lo, owner_lo, owner_hi, counts, nullptr, nullptr, n_candidates, s, lc);
    else
        launch_sweep<false, false>(
            L, owner_lo, owner_lo, owner_hi, counts, nullptr, nullptr, n_candidates, s, lc);
}

void launch_sweep_fill(
    const SortedList& L, int shard_lo, int owner_lo, int owner_hi,
    const unsigned long long* offsets, sccd_pair* pairs, cudaStream_t s, LaunchCounter& lc)
{
    if (L.two_lists)
        launch_sweep<true, true>(
            L, shard_lo, owner_lo, owner_hi, nullptr, offsets, pairs, nullptr, s, lc);
    else
        launch_sweep<true, false>(
            L, shard_lo, owner_lo, owner_hi, nullptr, offsets, pairs, nullptr, s, lc);
}

void launch_find_chunk_end(
    const unsigned long long* offsets, int lo, int hi, unsigned long long budget,
    int* d_out, cudaStream_t s, LaunchCounter& lc)
{
    find_chunk_end_kernel<<<1, 1, 0, s>>>(offsets, lo, hi, budget, d_out);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

} // namespace sccd
