// Major-axis sort (north-star item 2a): own one-launch-per-digit radix sort, prefix scan and
// record gather.  No library kernels.
//
// Replaces thrust::sort_by_key with a comparator on 16 B keys / 48 B values, the D2D copy,
// flip_element_ids and thrust::merge_by_key of the reference
// (cuda/broad_phase/aabb.cu:107-109, broad_phase.cu:57-101) by:
//   1. an LSD radix sort of 64-bit records  [ u32 key = [cell | quantised min.x | flags]
//      (common.cuh) | u32 box index ]  over only the key bits in use (3-4 digit passes) -- 8 B
//      per record per pass instead of 64 B.  Every pass is ONE launch: a tile of 4096 records
//      ranks its digits with warp match operations, learns where its runs start from the tiles
//      before it by decoupled look-back (a 32-bit status word per (tile, digit): flag | count),
//      and scatters through shared memory so that the global writes are runs of consecutive
//      addresses.  One more launch builds the histograms of all passes in a single read.
//   2. ONE gather that moves each 64 B exact record to its sorted position and emits the
//      24 B prefilter view (key, reach, f32 yz) the sweep streams.  Multi-GPU: the exact record
//      is rebuilt from the replicated vertex boxes instead of being read (gather_rebuild_kernel).
// The sort is stable, so ties keep element order and the result is deterministic.  The
// vertex-face list is sorted as one tagged list (vertex element ids are already flipped at
// build time), so no merge is needed.
//
// The same pass kernel with another digit functor -- "the rank that owns this record's cell" --
// is the stable partition of a rank's records by destination in the multi-GPU build (shard.cu).
// The exclusive scans (records per box -> offsets, pairs per owner -> offsets) are a single-pass
// chained scan with the same look-back scheme on 64-bit status words.
#include "boxmake.cuh"

#include <algorithm>

namespace sccd {

namespace {
constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// radix sort of 64-bit records
// ------------------------------------------------------------------------------------------
constexpr int kSortItems = 16;                   // records per thread
constexpr int kSortTile = kThreads * kSortItems; // 4096 records per CTA
constexpr int kRadix = 256;
constexpr int kMaxPasses = 4;
constexpr uint32_t kFlagPrefix = 2u << 30, kFlagAgg = 1u << 30, kValMask = (1u << 30) - 1u;

// digit = 8 bits of the record starting at `shift` (fewer in the top pass)
struct BitsDigit {
    int shift;
    uint32_t mask;
    __device__ __forceinline__ uint32_t operator()(unsigned long long r) const
    {
        return (uint32_t)(r >> shift) & mask;
    }
};
// digit = rank that owns the record's cell (cells are dealt out in contiguous ranges)
struct DestDigit {
    int cell_shift; // position of the cell field in the 64-bit record
    int world;
    uint32_t first_cell[17];
    __device__ __forceinline__ uint32_t operator()(unsigned long long r) const
    {
        const uint32_t cell = cell_shift >= 64 ? 0u : (uint32_t)(r >> cell_shift);
        // (fully unrolled, compile-time indices: the table stays in the kernel's constant bank;
        // entries past `world` are 0xffffffff and never count)
        uint32_t d = 0;
#pragma unroll
        for (int i = 1; i < 16; i++)
            d += cell >= first_cell[i] ? 1u : 0u;
        return d;
    }
};

// histograms of up to kMaxPasses digit positions in one read of the records
__global__ void __launch_bounds__(kThreads) radix_hist_kernel(
    const unsigned long long* __restrict__ rec, long long n, int first_shift, int passes,
    int top_bits, uint32_t* __restrict__ hist /* passes x 256 */)
{
    __shared__ uint32_t h[kMaxPasses * kRadix];
    for (int i = threadIdx.x; i < kMaxPasses * kRadix; i += kThreads)
        h[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        const unsigned long long r = rec[i];
        for (int p = 0; p < passes; p++) {
            const uint32_t mask = (p == passes - 1) ? ((1u << top_bits) - 1u) : 255u;
            atomicAdd(&h[p * kRadix + ((uint32_t)(r >> (first_shift + 8 * p)) & mask)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += kThreads)
        if (h[i])
            atomicAdd(&hist[i], h[i]);
}
// (d_n, where given: the record count lives on the device -- n is then only its upper bound)
template <typename Digit>
__global__ void __launch_bounds__(kThreads) digit_hist_kernel(
    const unsigned long long* __restrict__ rec, long long n, Digit digit, uint32_t* __restrict__ hist,
    const unsigned long long* __restrict__ d_n)
{
    __shared__ uint32_t h[kRadix];
    h[threadIdx.x] = 0;
    if (d_n)
        n = min(n, (long long)*d_n);
    __syncthreads();
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
        atomicAdd(&h[digit(rec[i])], 1u);
    __syncthreads();
    if (h[threadIdx.x])
        atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

struct SortSmem {
    unsigned long long rec[kSortTile];
    uint32_t warp_hist[kThreads / 32][kRadix]; // counts, then exclusive offsets over the warps
    uint32_t digit_off[kRadix];                // where each digit's run starts inside the tile
    uint32_t global_base[kRadix];              // ... and in the output
    uint32_t wsum_g[kThreads / 32], wsum_l[kThreads / 32];
    uint32_t tile;
};

// One digit pass.  status: n_tiles x 256 words, zero before the launch; ctr: tile ticket.
template <typename Digit>
__global__ void __launch_bounds__(kThreads, 4) radix_pass_kernel(
    const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, long long n,
    Digit digit, const uint32_t* __restrict__ hist /* 256: this pass */,
    uint32_t* __restrict__ status, uint32_t* __restrict__ ctr,
    const unsigned long long* __restrict__ d_n, bool gated)
{
    // gated: a sort that is only worth doing if its digit discriminates (see launch_sort_survivors)
    if (gated && 2ull * (unsigned long long)hist[0] >= *d_n)
        return;
    __shared__ SortSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        sm.tile = atomicAdd(ctr, 1u); // tiles start in ticket order: look-back cannot deadlock
    for (int i = tid; i < (kThreads / 32) * kRadix; i += kThreads)
        (&sm.warp_hist[0][0])[i] = 0;
    if (d_n)
        n = min(n, (long long)*d_n);
    __syncthreads();
    const uint32_t tile = sm.tile;
    const long long tile0 = (long long)tile * kSortTile;
    if (tile0 >= n)
        return; // (launched for the upper bound of a device-side count)
    const int tile_n = (int)min((long long)kSortTile, n - tile0);

    // ---- load (warp-striped: item i of lane l is record warp * 512 + i * 32 + l) and rank
    unsigned long long r[kSortItems];
    uint16_t rank[kSortItems];
    const int wbase = warp * (kSortItems * 32);
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        const int p = wbase + i * 32 + lane;
        r[i] = p < tile_n ? in[tile0 + p] : ~0ull;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        const int p = wbase + i * 32 + lane;
        const bool valid = p < tile_n;
        const uint32_t d = valid ? digit(r[i]) : (uint32_t)kRadix; // invalid: a digit of its own
        const unsigned peers = __match_any_sync(kFull, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = sm.warp_hist[warp][d];
            sm.warp_hist[warp][d] = old + (uint32_t)__popc(peers);
        }
        old = __shfl_sync(kFull, old, leader);
        rank[i] = (uint16_t)(old + (uint32_t)__popc(peers & ((1u << lane) - 1u)));
        __syncwarp();
    }
    __syncthreads();

    // ---- thread d owns digit d: offsets over the warps, look-back over the tiles
    {
        const int d = tid;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) {
            const uint32_t t = sm.warp_hist[w][d];
            sm.warp_hist[w][d] = run;
            run += t;
        }
        const uint32_t count = run;
        volatile uint32_t* st = status + (size_t)tile * kRadix + d;
        volatile const uint32_t* status_v = status;
        uint32_t excl = 0;
        if (tile == 0) {
            *st = kFlagPrefix | count;
        } else {
            *st = kFlagAgg | count;
            // look back over the tiles before this one, kWin status words per round trip (the
            // loads of a window are independent; a serial walk pays one L2 latency per tile,
            // which IS the run time of a small sort whose tiles all start together)
            constexpr int kWin = 8;
            bool done = false;
            for (long long t = (long long)tile - 1; t >= 0 && !done; t -= kWin) {
                uint32_t v[kWin];
#pragma unroll
                for (int j = 0; j < kWin; j++) {
                    v[j] = 2u << 30; // (before tile 0: a prefix of 0)
                    if (t - j >= 0)
                        v[j] = status_v[(size_t)(t - j) * kRadix + d];
                }
#pragma unroll
                for (int j = 0; j < kWin; j++) {
                    if (done)
                        break;
                    while ((v[j] & ~kValMask) == 0u) // not published yet
                        v[j] = status_v[(size_t)(t - j) * kRadix + d];
                    excl += v[j] & kValMask;
                    if (v[j] & kFlagPrefix)
                        done = true;
                }
            }
            *st = kFlagPrefix | (excl + count);
        }
        // exclusive scan over the digits: start of digit d in the whole array (from the global
        // histogram) and inside the tile
        const uint32_t g = hist[d], l = count;
        uint32_t gi = g, li = l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(kFull, gi, o), b = __shfl_up_sync(kFull, li, o);
            if (lane >= o)
                gi += a, li += b;
        }
        if (lane == 31)
            sm.wsum_g[warp] = gi, sm.wsum_l[warp] = li;
        __syncthreads();
        uint32_t og = 0, ol = 0;
        for (int w = 0; w < warp; w++)
            og += sm.wsum_g[w], ol += sm.wsum_l[w];
        sm.global_base[d] = og + gi - g + excl;
        sm.digit_off[d] = ol + li - l;
    }
    __syncthreads();

    // ---- scatter through shared memory: runs of equal digits become consecutive addresses
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        const int p = wbase + i * 32 + lane;
        if (p < tile_n) {
            const uint32_t d = digit(r[i]);
            sm.rec[sm.digit_off[d] + sm.warp_hist[warp][d] + rank[i]] = r[i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const int p = k * kThreads + tid;
        if (p < tile_n) {
            const unsigned long long v = sm.rec[p];
            const uint32_t d = digit(v);
            out[(size_t)sm.global_base[d] + (uint32_t)p - sm.digit_off[d]] = v;
        }
    }
}

inline int sort_tiles(long long m) { return (int)((m + kSortTile - 1) / kSortTile); }
// layout of the sort scratch: [hist: kMaxPasses x 256 | ctr: 64 | status: passes x tiles x 256]
inline size_t sort_scratch_words(long long m, int passes)
{
    return (size_t)kMaxPasses * kRadix + 64 + (size_t)passes * sort_tiles(m) * kRadix;
}

// ------------------------------------------------------------------------------------------
// exclusive scan u32 -> u64, single pass (chained look-back on 64-bit status words)
// ------------------------------------------------------------------------------------------
constexpr int kScanItems = 16;
constexpr int kScanTile = kThreads * kScanItems; // 4096
constexpr unsigned long long kSFlagAgg = 1ull << 62, kSFlagPrefix = 2ull << 62,
                             kSValMask = (1ull << 62) - 1ull;

__global__ void __launch_bounds__(kThreads) scan_u32_to_u64_kernel(
    const uint32_t* __restrict__ counts, unsigned long long* __restrict__ offsets, long long n_out,
    unsigned long long* __restrict__ status, uint32_t* __restrict__ ctr)
{
    __shared__ unsigned long long wsum[kThreads / 32];
    __shared__ unsigned long long tile_excl;
    __shared__ uint32_t tile_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
        tile_s = atomicAdd(ctr, 1u);
    __syncthreads();
    const uint32_t tile = tile_s;
    const long long base = (long long)tile * kScanTile + (long long)tid * kScanItems;
    uint32_t v[kScanItems];
    unsigned long long sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = base + i < n_out ? counts[base + i] : 0u;
        sum += v[i];
    }
    unsigned long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(kFull, incl, o);
        if (lane >= o)
            incl += a;
    }
    if (lane == 31)
        wsum[warp] = incl;
    __syncthreads();
    unsigned long long before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
        if (w < warp)
            before += wsum[w];
        total += wsum[w];
    }
    if (warp == 0) {
        // warp-wide look-back: 32 tiles before this one per round trip
        volatile unsigned long long* st = status;
        unsigned long long excl = 0;
        if (tile > 0) {
            if (lane == 0)
                st[tile] = kSFlagAgg | total;
            for (long long t0 = (long long)tile - 1;; t0 -= 32) {
                const long long t = t0 - lane;
                unsigned long long s = 2ull << 62; // (before tile 0: a prefix of 0)
                if (t >= 0)
                    s = st[t];
                while (__any_sync(kFull, (s & ~kSValMask) == 0ull))
                    if ((s & ~kSValMask) == 0ull)
                        s = st[t];
                const unsigned pm = __ballot_sync(kFull, (s & kSFlagPrefix) != 0ull);
                const int first = pm ? __ffs(pm) - 1 : 32; // nearest tile with a full prefix
                unsigned long long v = lane <= first ? (s & kSValMask) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    v += __shfl_xor_sync(kFull, v, o);
                excl += v;
                if (pm)
                    break;
            }
        }
        if (lane == 0) {
            st[tile] = kSFlagPrefix | (excl + total);
            tile_excl = excl;
        }
    }
    __syncthreads();
    unsigned long long run = tile_excl + before + incl - sum;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n_out)
            offsets[base + i] = run;
        run += v[i];
    }
}

// ------------------------------------------------------------------------------------------
// gathers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_sorted(
    int j, uint32_t key, const double2 x, const double4 yz, const int4 id, const BoxArrays& out,
    const PrefilterArrays& pf, const GridParams& g)
{
    out.x[j] = x;
    out.yz[j] = yz;
    out.id[j] = id;
    pf.key[j] = key;
    // same cell, q(xmax), every flag bit set (common.cuh)
    const int cell_shift = g.x_bits + kKeyFlagBits;
    const uint32_t cell_part = cell_shift >= 32 ? 0u : (key >> cell_shift) << cell_shift;
    pf.reach[j] = cell_part | (quantize_x(x.y, g) << kKeyFlagBits) | ((1u << kKeyFlagBits) - 1u);
    pf.yz[j] = make_float4(
        __double2float_rd(yz.x), __double2float_ru(yz.z), __double2float_rd(yz.y),
        __double2float_ru(yz.w));
}

__global__ void __launch_bounds__(kThreads) gather_sorted_kernel(
    int m, const unsigned long long* __restrict__ sorted_rec, BoxArrays in, BoxArrays out,
    PrefilterArrays pf, GridParams g)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= m)
        return;
    const unsigned long long r = sorted_rec[j];
    const uint32_t src = (uint32_t)r;
    store_sorted(
        j, (uint32_t)(r >> 32), __ldg(&in.x[src]), ldg_d4(&in.yz[src]), __ldg(&in.id[src]), out,
        pf, g);
}

// Multi-GPU receiver side: the exact record of a received (key, box index) record is REBUILT
// from the replicated per-vertex boxes and topology instead of being shipped (64 B) or gathered
// from a replica of every box of the mesh.  Same outputs as gather_sorted_kernel.
__global__ void __launch_bounds__(kThreads) gather_rebuild_kernel(
    int m, const unsigned long long* __restrict__ sorted_rec, MeshView mesh, int list, int axis,
    BoxArrays out, PrefilterArrays pf, GridParams g)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= m)
        return;
    const unsigned long long r = sorted_rec[j];
    double lo[3], hi[3];
    int4 id;
    make_list_box(mesh, list, (int)(uint32_t)r, lo, hi, id); // (indices were checked by the sender)
    double2 x;
    double4 yz;
    rotate_box(lo, hi, axis, x, yz);
    store_sorted(j, (uint32_t)(r >> 32), x, yz, id, out, pf, g);
}

__global__ void widen_counts_kernel(const uint32_t* hist, int world, unsigned long long* counts)
{
    if ((int)threadIdx.x < world)
        counts[threadIdx.x] = hist[threadIdx.x];
}

// run the digit passes of one sort; returns the buffer that holds the result
unsigned long long* sort_records(
    long long m, int lo_bit, int n_bits, unsigned long long* a, unsigned long long* b, void* temp,
    size_t temp_bytes, cudaStream_t s, LaunchCounter& lc, bool hist_ready = false)
{
    if (m <= 0 || n_bits <= 0)
        return a;
    const int passes = (n_bits + 7) / 8;
    if (passes > kMaxPasses)
        throw std::invalid_argument("sort: more than 32 key bits");
    const int top_bits = n_bits - 8 * (passes - 1);
    const int tiles = sort_tiles(m);
    const size_t words = sort_scratch_words(m, passes);
    if (temp_bytes < words * 4)
        throw std::logic_error("sort: scratch too small");
    uint32_t* hist = (uint32_t*)temp;
    uint32_t* ctr = hist + kMaxPasses * kRadix;
    uint32_t* status = ctr + 64;
    if (!hist_ready) { // (else: launch_sort_prepare + the kernel that made the records did it)
        SCCD_CUDA(cudaMemsetAsync(temp, 0, words * 4, s));
        radix_hist_kernel<<<std::min(tiles, 148 * 8), kThreads, 0, s>>>(
            a, m, lo_bit, passes, top_bits, hist);
        SCCD_CUDA(cudaGetLastError());
        lc.n++;
    }
    for (int p = 0; p < passes; p++) {
        BitsDigit dg;
        dg.shift = lo_bit + 8 * p;
        dg.mask = p == passes - 1 ? ((1u << top_bits) - 1u) : 255u;
        radix_pass_kernel<BitsDigit><<<tiles, kThreads, 0, s>>>(
            a, b, m, dg, hist + p * kRadix, status + (size_t)p * tiles * kRadix, ctr + p, nullptr,
            false);
        SCCD_CUDA(cudaGetLastError());
        lc.n++;
        std::swap(a, b);
    }
    return a;
}
} // namespace

size_t sort_temp_bytes(long long m) { return sort_scratch_words(m > 0 ? m : 1, kMaxPasses) * 4; }

// zeroes the sort scratch and returns where the digit histograms go, for a producer that
// builds them while it writes the records (grid.cu: expand_fill_kernel)
uint32_t* launch_sort_prepare(void* temp, size_t temp_bytes, long long m, cudaStream_t s)
{
    const size_t words = sort_scratch_words(m > 0 ? m : 1, kMaxPasses);
    if (temp_bytes < words * 4)
        throw std::logic_error("sort: scratch too small");
    SCCD_CUDA(cudaMemsetAsync(temp, 0, words * 4, s));
    return (uint32_t*)temp;
}

void launch_sort_and_gather(
    int m, int key_bits, unsigned long long* rec, unsigned long long* rec_tmp, void* temp,
    size_t temp_bytes, BoxArrays unsorted, SortedList out, cudaStream_t s, LaunchCounter& lc,
    cudaEvent_t gather_begin, cudaEvent_t gather_end, bool hist_ready)
{
    // the flag bits are not part of the order: equal (cell, q) records keep element order
    const unsigned long long* sorted = m > 0
        ? sort_records(
              m, 32 + kKeyFlagBits, key_bits, rec, rec_tmp, temp, temp_bytes, s, lc, hist_ready)
        : rec;
    if (gather_begin)
        SCCD_CUDA(cudaEventRecord(gather_begin, s));
    if (m > 0) {
        gather_sorted_kernel<<<(m + kThreads - 1) / kThreads, kThreads, 0, s>>>(
            m, sorted, unsorted, out.box, out.pf, out.grid);
        SCCD_CUDA(cudaGetLastError());
        lc.n++;
    }
    if (gather_end)
        SCCD_CUDA(cudaEventRecord(gather_end, s));
}

// ---- multi-GPU (shard.cu) ---------------------------------------------------------------
size_t partition_temp_bytes(long long m) { return sort_scratch_words(m > 0 ? m : 1, 1) * 4; }

// Stable partition of the records by the rank that owns their cell (one digit pass);
// counts[d] = records of destination d (device, `world` words of 64 bits).
void launch_partition_by_dest(
    long long m, const unsigned long long* rec_in, unsigned long long* rec_out, int cell_shift,
    const unsigned long long* h_first_cell /* world + 1, host */, int world,
    unsigned long long* counts, void* temp, size_t temp_bytes, cudaStream_t s, LaunchCounter& lc)
{
    if (world > 16)
        throw std::invalid_argument("partition: world too large");
    const size_t words = sort_scratch_words(m > 0 ? m : 1, 1);
    if (temp_bytes < words * 4)
        throw std::logic_error("partition: scratch too small");
    uint32_t* hist = (uint32_t*)temp;
    uint32_t* ctr = hist + kMaxPasses * kRadix;
    uint32_t* status = ctr + 64;
    SCCD_CUDA(cudaMemsetAsync(temp, 0, words * 4, s));
    if (m > 0) {
        DestDigit dg;
        dg.cell_shift = cell_shift;
        dg.world = world;
        for (int r = 0; r <= 16; r++) // (first_cell[world] = number of cells: no cell reaches it)
            dg.first_cell[r] = r < world ? (uint32_t)h_first_cell[r] : 0xffffffffu;
        const int tiles = sort_tiles(m);
        digit_hist_kernel<DestDigit><<<std::min(tiles, 148 * 8), kThreads, 0, s>>>(
            rec_in, m, dg, hist, nullptr);
        SCCD_CUDA(cudaGetLastError());
        radix_pass_kernel<DestDigit><<<tiles, kThreads, 0, s>>>(
            rec_in, rec_out, m, dg, hist, status, ctr, nullptr, false);
        SCCD_CUDA(cudaGetLastError());
        lc.n += 2;
    }
    widen_counts_kernel<<<1, 32, 0, s>>>(hist, world, counts);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

// Narrow phase: stable one-pass sort of the cull's survivor records on their lower-bound bucket
// (bits 32..39).  The count is on the device (d_n); n_max bounds it.  rec -> rec_out.
size_t sort_survivors_temp_bytes(long long n_max) { return partition_temp_bytes(n_max); }
void launch_sort_survivors(
    const unsigned long long* rec, unsigned long long* rec_out, const unsigned long long* d_n,
    const uint32_t* hist, long long n_max, void* temp, size_t temp_bytes, cudaStream_t s,
    LaunchCounter& lc)
{
    if (n_max <= 0)
        return;
    const size_t words = sort_scratch_words(n_max, 1);
    if (temp_bytes < words * 4)
        throw std::logic_error("sort_survivors: scratch too small");
    uint32_t* ctr = (uint32_t*)temp + kMaxPasses * kRadix;
    uint32_t* status = ctr + 64;
    SCCD_CUDA(cudaMemsetAsync(temp, 0, words * 4, s));
    BitsDigit dg;
    dg.shift = 32;
    dg.mask = 255u;
    radix_pass_kernel<BitsDigit><<<sort_tiles(n_max), kThreads, 0, s>>>(
        rec, rec_out, n_max, dg, hist, status, ctr, d_n, true);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

size_t sort_records_temp_bytes(long long m) { return sort_temp_bytes(m); }

// sorts m received records on key bits [kKeyFlagBits, kKeyFlagBits + key_bits) of their high
// word (stable: equal keys keep arrival order = global box order) and rebuilds the sorted views
void launch_sort_records_and_rebuild(
    int m, int key_bits, unsigned long long* rec, unsigned long long* rec_tmp, void* temp,
    size_t temp_bytes, const MeshView& mesh, int list, int axis, SortedList out, cudaStream_t s,
    LaunchCounter& lc, cudaEvent_t gather_begin, cudaEvent_t gather_end)
{
    const unsigned long long* sorted = m > 0
        ? sort_records(m, 32 + kKeyFlagBits, key_bits, rec, rec_tmp, temp, temp_bytes, s, lc)
        : rec;
    if (gather_begin)
        SCCD_CUDA(cudaEventRecord(gather_begin, s));
    if (m > 0) {
        gather_rebuild_kernel<<<(m + kThreads - 1) / kThreads, kThreads, 0, s>>>(
            m, sorted, mesh, list, axis, out.box, out.pf, out.grid);
        SCCD_CUDA(cudaGetLastError());
        lc.n++;
    }
    if (gather_end)
        SCCD_CUDA(cudaEventRecord(gather_end, s));
}

size_t scan_temp_bytes(int n)
{
    const long long tiles = ((long long)n + 1 + kScanTile - 1) / kScanTile;
    return (size_t)(tiles + 8) * 8;
}

void launch_scan_u32_to_u64(
    const uint32_t* counts, unsigned long long* offsets, int n, void* temp, size_t temp_bytes,
    cudaStream_t s, LaunchCounter& lc)
{
    // counts has n+1 readable entries (the last one is zero) so that offsets[n] = total.
    const long long n_out = (long long)n + 1;
    const long long tiles = (n_out + kScanTile - 1) / kScanTile;
    if (temp_bytes < (size_t)(tiles + 8) * 8)
        throw std::logic_error("scan: scratch too small");
    SCCD_CUDA(cudaMemsetAsync(temp, 0, (size_t)(tiles + 8) * 8, s));
    unsigned long long* status = (unsigned long long*)temp + 8;
    scan_u32_to_u64_kernel<<<(unsigned)tiles, kThreads, 0, s>>>(
        counts, offsets, n_out, status, (uint32_t*)temp);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

} // namespace sccd
