// Sweep over the (cell, major-axis)-sorted records (north-star item 2b).
//
// Replaces sweep_and_tiniest_queue (cuda/broad_phase/sweep.cu:101-182: one warp per CTA,
// a 64-slot shared ring rebalanced with shared atomics every round, uncoalesced 48 B
// MiniBox loads per candidate, one global atomicAdd per emitted pair into a buffer that is
// value-initialised on every call) with a sliding-window sweep:
//
//   * a warp owns 32 consecutive sorted records ("owners", one per lane).  In step k every
//     lane looks at the record k places after its own: the 32 loads of a step are
//     consecutive addresses (one coalesced request), and consecutive steps re-read the same
//     lines from L1.  With the (y, z) cell grid a window holds a few dozen records, so this
//     does ~32 tests per ~10 instructions with no staging, no barriers and no wasted
//     owner x candidate combinations;
//   * the prefilter is one unsigned compare key_j <= reach_i (same cell AND quantised
//     xmin_j <= xmax_i), two bit operations on the key's flag bits (other list; pair is at
//     home in this cell -- see common.cuh) and four f32 compares on y / z; the 16-byte yz
//     record is only loaded for candidates that pass the key tests;
//   * prefilter survivors (a conservative superset: min rounded down / max rounded up) are
//     ballot-compacted into a per-warp shared queue -- the "tiniest queue" -- and drained 32
//     at a time with ALL lanes running the exact double test + vertex-sharing test, so the
//     rare expensive path never diverges;
//   * output is count -> exclusive scan -> place: the pair list is exactly sized, ordered by
//     (owner position, candidate position) and therefore deterministic; no global atomics
//     and no giant memset; 64-bit offsets.  The sweep itself runs ONCE: the count pass parks
//     every pair it finds, tagged (owner, rank within the owner), in a fixed-size staging
//     area of its tile, and the place pass only moves the parked pairs to offsets[owner] +
//     rank.  A tile whose pairs outgrow its staging area (kStage) is swept again by the
//     place pass -- the exception, not the rule.
//
// The emitted SET equals the reference's: every pair with closed overlap on x, y, z
// (cuda/broad_phase/aabb.cuh:100-104, sweep.cu:131,173), valid list membership
// (collision.cuh:27-35) and no shared vertex (collision.cuh:17-21).  Why nothing is lost or
// duplicated: (1) sorting on a quantised key instead of the double only changes the ORDER in
// which ties are visited -- the window test on quantised values is a superset of
// min_j <= max_i and the exact test checks both x directions; (2) two overlapping boxes both
// own a record in the cell of (max(ymin_a,ymin_b), max(zmin_a,zmin_b)) because cell_index()
// is monotone, and the pair is accepted in that cell only (flag bits fy / fz).
#include "common.cuh"

#include <cfloat>

namespace sccd {

namespace {

constexpr int kTile = 256; // owners per CTA == threads per CTA
constexpr int kWarps = kTile / 32;
constexpr int kQueueCap = 64;
constexpr int kStage = 1024; // pairs a tile can park between the count and the place pass
constexpr int kRelBits = 27; // candidate position relative to the tile start
constexpr unsigned kFull = 0xffffffffu;
constexpr int kRankBits = 24; // staging tag = owner-in-tile << 24 | rank within the owner

struct SweepSmem {
    unsigned long long o_off[kTile];
    uint32_t o_cnt[kTile];
    uint32_t q[kWarps][kQueueCap];
    uint32_t staged; // pairs this tile has parked (or tried to park) so far
};

// ---- TMA-staged variant of the count pass (SCCD_OPT_SWEEP_STAGED) --------------------------------
// The window loop reads the prefilter stream of the records that FOLLOW the tile's owners: keys
// (4 B) and, for candidates that pass the key tests, the f32 yz record (16 B).  Here one thread
// of the CTA brings the next kStageRecs records of both arrays into shared memory with two bulk
// async copies (cp.async.bulk, completion on an mbarrier) before the sweep starts; the loop then
// reads shared memory and only falls back to global memory for windows that reach beyond the
// staged records.  A/B against the L1-resident loop: DESIGN.md 3.3.
constexpr int kStageRecs = 1024;
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase)
{
    unsigned done = 0;
    while (!done)
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
}
struct __align__(16) StageSmem {
    float4 yz[kStageRecs];
    uint32_t key[kStageRecs];
    unsigned long long bar;
};

struct Staging {
    sccd_pair* pairs = nullptr; // kStage per tile
    uint32_t* tags = nullptr;   // kStage per tile
    uint32_t* count = nullptr;  // one per tile; > kStage: the tile overflowed
};

// Exact test of up to 32 queued (owner lane, candidate) entries, one per lane.
template <bool FILL, bool TWO_LISTS>
__device__ __forceinline__ void drain32(
    SweepSmem& sm, const BoxArrays& box, int tile0, int warp, int lane, uint32_t entry,
    bool active, sccd_pair* __restrict__ pairs, sccd_pair* __restrict__ st_pairs,
    uint32_t* __restrict__ st_tags)
{
    const int ol = active ? (int)(entry >> kRelBits) : 0;
    const int t = warp * 32 + ol;
    const int j = tile0 + (int)(entry & ((1u << kRelBits) - 1));
    bool hit = false;
    int ea = 0, eb = 0;
    if (active) {
        const int i = tile0 + t;
        // ids first: mesh neighbours always overlap and are never admissible
        // (collision.cuh:17-21) -- two thirds of a mesh's survivors stop here, 32 bytes in,
        // instead of after the 128-byte exact test.  (Doing this in the window loop instead
        // costs more than it saves: that loop is latency-bound, this stage is not.)
        const int4 aid = __ldg(&box.id[i]);
        const int4 bid = __ldg(&box.id[j]);
        const bool share = aid.x == bid.x || aid.x == bid.y || aid.x == bid.z
            || aid.y == bid.x || aid.y == bid.y || aid.y == bid.z || aid.z == bid.x
            || aid.z == bid.y || aid.z == bid.z;
        if (!share) {
            const double2 ax = __ldg(&box.x[i]);
            const double4 ayz = ldg_d4(&box.yz[i]);
            const double2 bx = __ldg(&box.x[j]);
            const double4 byz = ldg_d4(&box.yz[j]);
            // closed-interval overlap on all three axes (aabb.cuh:67-72 / 100-104)
            hit = ax.y >= bx.x && ax.x <= bx.y && ayz.z >= byz.x && ayz.x <= byz.z
                && ayz.w >= byz.y && ayz.y <= byz.w;
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
        const uint32_t cur = active ? sm.o_cnt[t] : 0;
        __syncwarp();
        if (hits) {
            // park the pairs of this batch in the tile's staging area
            uint32_t base = 0;
            const int first = __ffs(hits) - 1;
            if (lane == first)
                base = atomicAdd(&sm.staged, (uint32_t)__popc(hits));
            base = __shfl_sync(kFull, base, first);
            const uint32_t slot = base + __popc(hits & ((1u << lane) - 1));
            const uint32_t rank = cur + (uint32_t)__popc(mine & ((1u << lane) - 1));
            if (hit && slot < (uint32_t)kStage && rank < (1u << kRankBits)) {
                const int mn = min(ea, eb), mx = max(ea, eb);
                sccd_pair p;
                p.a = TWO_LISTS ? (-mn - 1) : mn;
                p.b = mx;
                st_pairs[slot] = p;
                st_tags[slot] = ((uint32_t)t << kRankBits) | rank;
            }
            if (hit && rank >= (1u << kRankBits))
                atomicAdd(&sm.staged, (uint32_t)kStage); // cannot be tagged: force a re-sweep
        }
        if (leader && tot)
            sm.o_cnt[t] = cur + tot;
        __syncwarp();
    }
}

// Sweep of the tile of kTile owners starting at tile0; only owners in [owner_lo, owner_hi)
// take part.  FILL = false: count (counts[]) and park the pairs (staging, st_*);
// FILL = true: write the pairs of the chunk that starts at owner chunk_lo to pairs[].
template <bool FILL, bool TWO_LISTS, bool STAGED = false>
__device__ __forceinline__ void sweep_tile(
    SweepSmem& sm, const PrefilterArrays& pf, const BoxArrays& box, int n, int shard_lo,
    int tile0, int owner_lo, int owner_hi, int chunk_lo, uint32_t* __restrict__ counts,
    const unsigned long long* __restrict__ offsets, sccd_pair* __restrict__ pairs,
    unsigned long long* __restrict__ n_candidates, sccd_pair* __restrict__ st_pairs,
    uint32_t* __restrict__ st_tags, const StageSmem* stg = nullptr, int staged = 0)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int i = tile0 + tid;
    const bool valid = i >= owner_lo && i < owner_hi;

    uint32_t my_reach = 0, my_flags = 0;
    float4 my = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        my_reach = __ldg(&pf.reach[i]);
        my_flags = __ldg(&pf.key[i]) & ((1u << kKeyFlagBits) - 1u);
        my = __ldg(&pf.yz[i]);
        if (FILL) // position of this owner's first pair inside the chunk being filled
            sm.o_off[tid] = offsets[i - shard_lo] - offsets[chunk_lo - shard_lo];
    }
    sm.o_cnt[tid] = 0;
    __syncwarp();

    uint32_t* q = sm.q[warp];
    int qn = 0;                    // entries waiting in this warp's queue (warp-uniform)
    unsigned long long tested = 0; // exact tests run by this warp (warp-uniform)
    const uint32_t rel_base = (uint32_t)(i - tile0);
    // the pair must be at home in this cell, and (two lists) join a vertex with a face
    constexpr uint32_t kHome = kKeyFlagY | kKeyFlagZ;

    // kStep candidates per trip: their keys (then the yz records of those that pass the key
    // tests) are loaded together, so a trip waits for two L1 round trips instead of 2 * kStep
    // (in the single-candidate loop 7 of 13 stalled warps per issue were waiting on a load).
    // Tried and dropped: keeping the owners' exact records in shared memory and skipping the
    // exact fetch when inner f32 bounds already prove the overlap -- reading 64 B for EVERY
    // owner costs more than the survivors' fetches it saves (config 4: 8.2 -> 17.8 ms).
    constexpr int kStep = 4;
    for (int k = 1;; k += kStep) {
        uint32_t kj[kStep];
#pragma unroll
        for (int u = 0; u < kStep; u++) {
            const int j = i + k + u;
            if (STAGED)
                kj[u] = (valid && j < n)
                    ? (j - tile0 < staged ? stg->key[j - tile0] : __ldg(&pf.key[j]))
                    : 0xffffffffu;
            else
                kj[u] = (valid && j < n) ? __ldg(&pf.key[j]) : 0xffffffffu;
        }
        // keys are sorted: nothing later can be in a window if the first of the trip is in none
        if (!__any_sync(kFull, valid && kj[0] <= my_reach && i + k < n))
            break;
        bool p[kStep];
        float4 b[kStep];
#pragma unroll
        for (int u = 0; u < kStep; u++) {
            const int j = i + k + u;
            p[u] = valid && j < n && kj[u] <= my_reach && ((kj[u] | my_flags) & kHome) == kHome;
            if (TWO_LISTS)
                p[u] = p[u] && ((kj[u] ^ my_flags) & kKeyFlagType) != 0u;
            b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p[u])
                b[u] = (STAGED && j - tile0 < staged) ? stg->yz[j - tile0] : __ldg(&pf.yz[j]);
        }
#pragma unroll
        for (int u = 0; u < kStep; u++) {
            const bool hit =
                p[u] && (b[u].x <= my.y) && (my.x <= b[u].y) && (b[u].z <= my.w) && (my.z <= b[u].w);
            const unsigned mask = __ballot_sync(kFull, hit);
            if (mask == 0u)
                continue;
            if (hit)
                q[qn + __popc(mask & ((1u << lane) - 1u))] =
                    ((uint32_t)lane << kRelBits) | (rel_base + (uint32_t)(k + u));
            qn += __popc(mask);
            __syncwarp();
            if (qn >= 32) {
                drain32<FILL, TWO_LISTS>(
                    sm, box, tile0, warp, lane, q[lane], true, pairs, st_pairs, st_tags);
                tested += 32;
                const int r = qn - 32;
                const uint32_t v = (lane < r) ? q[32 + lane] : 0u;
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
            sm, box, tile0, warp, lane, lane < qn ? q[lane] : 0u, lane < qn, pairs, st_pairs,
            st_tags);
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


// count pass: one tile per CTA, tiles start at shard_lo
template <bool TWO_LISTS>
__global__ void __launch_bounds__(kTile) sweep_count_kernel(
    PrefilterArrays pf, BoxArrays box, int n, int shard_lo, int shard_hi,
    uint32_t* __restrict__ counts, unsigned long long* __restrict__ n_candidates, Staging st)
{
    __shared__ SweepSmem sm;
    if (threadIdx.x == 0)
        sm.staged = 0;
    __syncthreads();
    const int tile0 = shard_lo + blockIdx.x * kTile;
    sweep_tile<false, TWO_LISTS>(
        sm, pf, box, n, shard_lo, tile0, shard_lo, shard_hi, shard_lo, counts, nullptr, nullptr,
        n_candidates, st.pairs + (size_t)blockIdx.x * kStage,
        st.tags + (size_t)blockIdx.x * kStage);
    __syncthreads();
    if (threadIdx.x == 0)
        st.count[blockIdx.x] = sm.staged;
}

// count pass with the prefilter stream of the tile staged in shared memory by bulk async copies
template <bool TWO_LISTS>
__global__ void __launch_bounds__(kTile) sweep_count_staged_kernel(
    PrefilterArrays pf, BoxArrays box, int n, int shard_lo, int shard_hi,
    uint32_t* __restrict__ counts, unsigned long long* __restrict__ n_candidates, Staging st)
{
    __shared__ SweepSmem sm;
    __shared__ StageSmem stg;
    const int tile0 = shard_lo + blockIdx.x * kTile;
    // whole 16-byte units only, from a 16-byte aligned start (else: nothing staged)
    int staged = min(kStageRecs, n - tile0) & ~3;
    if ((tile0 & 3) != 0 || staged < 0)
        staged = 0;
    if (threadIdx.x == 0) {
        sm.staged = 0;
        mbar_init(&stg.bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0 && staged > 0) {
        mbar_expect_tx(&stg.bar, (unsigned)staged * 20u);
        bulk_g2s(stg.key, pf.key + tile0, (unsigned)staged * 4u, &stg.bar);
        bulk_g2s(stg.yz, pf.yz + tile0, (unsigned)staged * 16u, &stg.bar);
    }
    if (staged > 0)
        mbar_wait(&stg.bar, 0);
    sweep_tile<false, TWO_LISTS, true>(
        sm, pf, box, n, shard_lo, tile0, shard_lo, shard_hi, shard_lo, counts, nullptr, nullptr,
        n_candidates, st.pairs + (size_t)blockIdx.x * kStage,
        st.tags + (size_t)blockIdx.x * kStage, &stg, staged);
    __syncthreads();
    if (threadIdx.x == 0)
        st.count[blockIdx.x] = sm.staged;
}

// place pass over the count tiles that overlap the owner chunk [owner_lo, owner_hi)
template <bool TWO_LISTS>
__global__ void __launch_bounds__(kTile) sweep_place_kernel(
    PrefilterArrays pf, BoxArrays box, int n, int shard_lo, int first_tile, int owner_lo,
    int owner_hi, const unsigned long long* __restrict__ offsets, sccd_pair* __restrict__ pairs,
    Staging st)
{
    __shared__ SweepSmem sm;
    const int tile = first_tile + blockIdx.x;
    const int tile0 = shard_lo + tile * kTile;
    const uint32_t staged = st.count[tile];
    if (staged <= (uint32_t)kStage) {
        // the usual case: move the parked pairs to offsets[owner] + rank
        const sccd_pair* sp = st.pairs + (size_t)tile * kStage;
        const uint32_t* stg = st.tags + (size_t)tile * kStage;
        const unsigned long long chunk_base = offsets[owner_lo - shard_lo];
        for (uint32_t e = threadIdx.x; e < staged; e += kTile) {
            const uint32_t tag = stg[e];
            const int owner = tile0 + (int)(tag >> kRankBits);
            if (owner >= owner_lo && owner < owner_hi)
                pairs[offsets[owner - shard_lo] - chunk_base + (tag & ((1u << kRankBits) - 1u))] =
                    sp[e];
        }
        return;
    }
    sweep_tile<true, TWO_LISTS>(
        sm, pf, box, n, shard_lo, tile0, owner_lo, owner_hi, owner_lo, nullptr, offsets, pairs,
        nullptr, nullptr, nullptr);
}

// window[i] = #records j > i with key_j <= reach_i (sweep work estimate used to balance
// owner ranges across GPUs when a list has too few cells to be split by cell range).
__global__ void __launch_bounds__(256)
    sweep_window_kernel(PrefilterArrays pf, int n, uint32_t* __restrict__ window)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t reach = pf.reach[i];
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

} // namespace

size_t sweep_stage_tiles(int owners) { return (size_t)((owners + kTile - 1) / kTile); }
size_t sweep_stage_pair_bytes(int owners)
{
    return sweep_stage_tiles(owners) * kStage * sizeof(sccd_pair);
}
size_t sweep_stage_tag_bytes(int owners)
{
    return sweep_stage_tiles(owners) * kStage * sizeof(uint32_t);
}

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
    unsigned long long* n_candidates, void* stage_pairs, void* stage_tags, uint32_t* stage_count,
    cudaStream_t s, LaunchCounter& lc, bool tma_staged)
{
    if (L.n >= (1 << kRelBits))
        throw std::runtime_error("sweep: more than 2^27 records in one list is not supported");
    const int owners = owner_hi - owner_lo;
    if (owners <= 0)
        return;
    Staging st;
    st.pairs = (sccd_pair*)stage_pairs;
    st.tags = (uint32_t*)stage_tags;
    st.count = stage_count;
    const int grid = (owners + kTile - 1) / kTile;
    if (tma_staged && L.two_lists)
        sweep_count_staged_kernel<true><<<grid, kTile, 0, s>>>(
            L.pf, L.box, L.n, owner_lo, owner_hi, counts, n_candidates, st);
    else if (tma_staged)
        sweep_count_staged_kernel<false><<<grid, kTile, 0, s>>>(
            L.pf, L.box, L.n, owner_lo, owner_hi, counts, n_candidates, st);
    else if (L.two_lists)
        sweep_count_kernel<true><<<grid, kTile, 0, s>>>(
            L.pf, L.box, L.n, owner_lo, owner_hi, counts, n_candidates, st);
    else
        sweep_count_kernel<false><<<grid, kTile, 0, s>>>(
            L.pf, L.box, L.n, owner_lo, owner_hi, counts, n_candidates, st);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_sweep_fill(
    const SortedList& L, int shard_lo, int owner_lo, int owner_hi,
    const unsigned long long* offsets, sccd_pair* pairs, void* stage_pairs, void* stage_tags,
    uint32_t* stage_count, cudaStream_t s, LaunchCounter& lc)
{
    if (owner_hi <= owner_lo)
        return;
    Staging st;
    st.pairs = (sccd_pair*)stage_pairs;
    st.tags = (uint32_t*)stage_tags;
    st.count = stage_count;
    const int first_tile = (owner_lo - shard_lo) / kTile;
    const int last_tile = (owner_hi - 1 - shard_lo) / kTile;
    const int grid = last_tile - first_tile + 1;
    if (L.two_lists)
        sweep_place_kernel<true><<<grid, kTile, 0, s>>>(
            L.pf, L.box, L.n, shard_lo, first_tile, owner_lo, owner_hi, offsets, pairs, st);
    else
        sweep_place_kernel<false><<<grid, kTile, 0, s>>>(
            L.pf, L.box, L.n, shard_lo, first_tile, owner_lo, owner_hi, offsets, pairs, st);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
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
