// Cell-striped sort keys for the sweep (north-star item 2: "major-axis projection").
//
// A 1-axis sweep tests every box against ALL boxes whose x interval starts inside its own --
// on a flat cloth that is ~2,400 candidates per box for ~25 real overlaps, and the sweep is
// bound by those tests, not by HBM.  Here each box is also binned into a uniform (y, z) cell
// grid: it is replicated into every cell its closed yz range touches, records are sorted on
// (cell, xmin), and the sweep window of a record only contains records of the same cell.
// A pair is reported in exactly one cell (see GridParams), so the emitted SET is unchanged.
//
// Kernels: box statistics (grid choice), copies-per-box count, (key, index) expansion,
// per-cell histogram + balanced cell ranges for the multi-GPU split.
#include "common.cuh"

#include <cfloat>

namespace sccd {

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
    box_stats_kernel(BoxArrays boxes, int n, int stride, double* __restrict__ partials)
{
    double r[kNumStats];
    stats_identity(r);
    const long long step = (long long)gridDim.x * kThreads * stride;
    for (long long ii = ((long long)blockIdx.x * kThreads + threadIdx.x) * stride; ii < n;
         ii += step) {
        const int i = (int)ii;
        const double4 b = ldg_d4(&boxes.yz[i]); // (ymin, zmin, ymax, zmax)
        const double2 x = __ldg(&boxes.x[i]);
        const double lo[3] = { x.x, b.x, b.y }, hi[3] = { x.y, b.z, b.w };
        double v[kNumStats];
        stats_of_box(v, lo, hi);
        stats_merge(r, v);
    }
    __shared__ double red[32 * kNumStats];
    stats_block_reduce(r, red);
    if (threadIdx.x == 0)
        for (int k = 0; k < kNumStats; k++)
            partials[blockIdx.x * kNumStats + k] = r[k];
}

// one CTA; fixed reduction tree => deterministic sums
__global__ void __launch_bounds__(kThreads)
    box_stats_final_kernel(const double* __restrict__ partials, int blocks, double scale, double* out)
{
    double r[kNumStats];
    stats_identity(r);
    for (int b = threadIdx.x; b < blocks; b += kThreads)
        stats_merge(r, partials + b * kNumStats);
    __shared__ double red[32 * kNumStats];
    stats_block_reduce(r, red);
    if (threadIdx.x == 0) {
        r[4] *= scale; // sums over a 1-in-stride sample -> estimates for the list
        r[5] *= scale;
        for (int k = 0; k < kNumStats; k++)
            out[k] = r[k];
    }
}

__device__ __forceinline__ void cell_range(
    const double4 yz, const GridParams& g, int& y0, int& y1, int& z0, int& z1)
{
    y0 = cell_index(yz.x, g.y0, g.inv_hy, g.sy);
    y1 = cell_index(yz.z, g.y0, g.inv_hy, g.sy);
    z0 = cell_index(yz.y, g.z0, g.inv_hz, g.sz);
    z1 = cell_index(yz.w, g.z0, g.inv_hz, g.sz);
}

// cells of row cy, columns [z0, z1], that fall in this rank's linear range [cell_lo, cell_hi)
__device__ __forceinline__ void clip_row(const GridParams& g, int cy, int& z0, int& z1)
{
    const long long base = (long long)cy * g.sz;
    const long long lo = (long long)g.cell_lo - base, hi = (long long)g.cell_hi - base - 1;
    if (lo > z0)
        z0 = (int)(lo < (long long)g.sz ? lo : (long long)g.sz);
    if (hi < z1)
        z1 = (int)(hi > -1 ? hi : -1);
}

__global__ void __launch_bounds__(kThreads) expand_count_kernel(
    BoxArrays boxes, int n, GridParams g, const unsigned long long* __restrict__ d_range,
    uint32_t* __restrict__ copies)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n)
        return;
    if (d_range) { // this rank's cell range, still on the device (launch_cell_splits)
        g.cell_lo = (int)d_range[0];
        g.cell_hi = (int)d_range[1];
    }
    int y0, y1, z0, z1;
    cell_range(ldg_d4(&boxes.yz[i]), g, y0, y1, z0, z1);
    uint32_t k = 0;
    for (int cy = y0; cy <= y1; cy++) {
        int a = z0, b = z1;
        clip_row(g, cy, a, b);
        if (b >= a)
            k += (uint32_t)(b - a + 1);
    }
    copies[i] = k;
}

// hist[cell] += 1 for every cell a sampled box (every stride-th) touches: the multi-GPU cell
// ranges only have to be balanced, not exact, and one atomic per record of a 50M-box list
// costs more than sorting this rank's share of it.
__global__ void __launch_bounds__(kThreads) cell_hist_kernel(
    BoxArrays boxes, int n, int stride, GridParams g, uint32_t* __restrict__ hist)
{
    const long long s = ((long long)blockIdx.x * kThreads + threadIdx.x) * stride;
    if (s >= n)
        return;
    const int i = (int)s;
    int y0, y1, z0, z1;
    cell_range(ldg_d4(&boxes.yz[i]), g, y0, y1, z0, z1);
    for (int cy = y0; cy <= y1; cy++)
        for (int cz = z0; cz <= z1; cz++)
            atomicAdd(&hist[cy * g.sz + cz], 1u);
}

// Contiguous cell ranges of ~equal sweep work for `world` ranks.  Work estimate per cell = its
// record count: with the cell grid the windows are short and the measured sweep time is
// proportional to the records (config 4: 0.18-0.20 us per 1000 records on every rank), as are
// sort and gather.  Exact integer arithmetic, so every rank derives the same split points from
// its own replica.  out[0..world] = first cell of each rank (out[world] = cells),
// out[world+1] = total SAMPLED records, out[world+2+r] = sampled records of rank r.  One CTA.
__global__ void __launch_bounds__(1024) cell_splits_kernel(
    const uint32_t* __restrict__ hist, int cells, int world, unsigned long long* __restrict__ out)
{
    __shared__ unsigned long long part[1024];
    __shared__ unsigned long long total_s;
    __shared__ unsigned long long recs[64];
    __shared__ int split_s[65];
    const int t = threadIdx.x;
    const int per = (cells + 1023) / 1024;
    const int c0 = min(cells, t * per), c1 = min(cells, c0 + per);
    unsigned long long s = 0;
    for (int c = c0; c < c1; c++)
        s += hist[c];
    part[t] = s;
    if (t <= world) {
        split_s[t] = t == 0 ? 0 : cells;
        recs[t] = 0;
    }
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) {
            const unsigned long long v = part[i];
            part[i] = run;
            run += v;
        }
        total_s = run;
    }
    __syncthreads();
    const unsigned long long total = total_s;
    unsigned long long run = part[t];
    for (int c = c0; c < c1; c++) {
        const unsigned long long before = run;
        run += hist[c];
        for (int r = 1; r < world; r++) {
            const unsigned long long target = total / (unsigned long long)world * r;
            if (before < target && target <= run)
                split_s[r] = c + 1; // cell c is the last one of rank r - 1
        }
    }
    __syncthreads();
    // a target of 0 (empty list) is never crossed: keep the splits monotone
    if (t == 0)
        for (int r = world - 1; r >= 1; r--)
            split_s[r] = min(split_s[r], split_s[r + 1]);
    __syncthreads();
    unsigned long long mine[16] = {};
    for (int c = c0; c < c1; c++) {
        int r = 0;
        while (r + 1 < world && c >= split_s[r + 1])
            r++;
        if (r < 16)
            mine[r] += hist[c];
    }
    for (int r = 0; r < world && r < 16; r++)
        if (mine[r])
            atomicAdd(&recs[r], mine[r]);
    __syncthreads();
    if (t <= world)
        out[t] = (unsigned long long)split_s[t];
    if (t < world)
        out[world + 2 + t] = recs[t];
    if (t == 0) {
        unsigned long long m = 0;
        for (int r = 0; r < world; r++)
            m += recs[r];
        out[world + 1] = m;
    }
}

template <bool HIST>
__device__ __forceinline__ void expand_fill_one(
    const BoxArrays& boxes, int i, const GridParams& g, const unsigned long long* __restrict__ offsets,
    uint32_t idx_base, unsigned long long* __restrict__ rec, uint32_t* h, int hist_shift,
    int hist_passes, int hist_top)
{
    // multi-GPU (replicated build): most boxes have no record in this rank's cell range -- 16
    // bytes instead of 64.  (Measured and dropped: fusing count + scan + fill into one
    // chained-scan pass; with 256-box tiles the look-back latency made it 40 % slower.)
    if (offsets[i + 1] == offsets[i])
        return;
    int y0, y1, z0, z1;
    cell_range(ldg_d4(&boxes.yz[i]), g, y0, y1, z0, z1);
    // key layout: common.cuh ("32-bit sweep key of a record"); record = key << 32 | box index
    const uint32_t xq = quantize_x(__ldg(&boxes.x[i]).x, g) << kKeyFlagBits;
    const uint32_t type = __ldg(&boxes.id[i]).w < 0 ? kKeyFlagType : 0u;
    const int cell_shift = g.x_bits + kKeyFlagBits;
    const unsigned long long idx = (unsigned long long)(idx_base + (uint32_t)i);
    unsigned long long o = offsets[i];
    for (int cy = y0; cy <= y1; cy++) {
        int a = z0, b = z1;
        clip_row(g, cy, a, b);
        for (int cz = a; cz <= b; cz++) {
            const uint32_t cell = (uint32_t)(cy * g.sz + cz);
            const uint32_t hi = cell_shift >= 32 ? 0u : (cell << cell_shift);
            const uint32_t key =
                hi | xq | type | (cy == y0 ? kKeyFlagY : 0u) | (cz == z0 ? kKeyFlagZ : 0u);
            const unsigned long long r = ((unsigned long long)key << 32) | idx;
            rec[o++] = r;
            if (HIST)
                for (int p = 0; p < hist_passes; p++) {
                    const uint32_t mask = p == hist_passes - 1 ? ((1u << hist_top) - 1u) : 255u;
                    atomicAdd(&h[p * 256 + ((uint32_t)(r >> (hist_shift + 8 * p)) & mask)], 1u);
                }
        }
    }
}

// hist (optional): the digit histograms of the radix sort that follows (sort.cu), built here while
// the keys are in registers -- one launch and one read of the records less on the critical path.
// hist_passes digits of 8 bits from bit hist_shift of the 64-bit record, the top one hist_top wide.
template <bool HIST>
__global__ void __launch_bounds__(kThreads) expand_fill_kernel(
    BoxArrays boxes, int n, GridParams g, const unsigned long long* __restrict__ offsets,
    uint32_t idx_base, unsigned long long* __restrict__ rec, uint32_t* __restrict__ hist,
    int hist_shift, int hist_passes, int hist_top)
{
    __shared__ uint32_t h[HIST ? 4 * 256 : 1];
    if (HIST) {
        for (int j = threadIdx.x; j < 4 * 256; j += kThreads)
            h[j] = 0;
        __syncthreads();
    }
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i < n)
        expand_fill_one<HIST>(boxes, i, g, offsets, idx_base, rec, h, hist_shift, hist_passes, hist_top);
    if (HIST) {
        __syncthreads();
        for (int j = threadIdx.x; j < hist_passes * 256; j += kThreads)
            if (h[j])
                atomicAdd(&hist[j], h[j]);
    }
}

} // namespace

void launch_expand_fill_records(
    const BoxArrays& unsorted, int n, GridParams g, const unsigned long long* offsets,
    uint32_t idx_base, unsigned long long* rec, cudaStream_t s, LaunchCounter& lc)
{
    if (n <= 0)
        return;
    expand_fill_kernel<false><<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        unsorted, n, g, offsets, idx_base, rec, nullptr, 0, 0, 0);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_box_stats(
    const BoxArrays& unsorted, int n, int stride, double* partials, double* stats,
    cudaStream_t s, LaunchCounter& lc, double sum_scale)
{
    if (sum_scale <= 0.0)
        sum_scale = (double)stride;
    int blocks = ((n + stride - 1) / stride + kThreads - 1) / kThreads;
    if (blocks > kStatsBlocks)
        blocks = kStatsBlocks;
    if (blocks < 1)
        blocks = 1;
    box_stats_kernel<<<blocks, kThreads, 0, s>>>(unsorted, n, stride, partials);
    SCCD_CUDA(cudaGetLastError());
    box_stats_final_kernel<<<1, kThreads, 0, s>>>(partials, blocks, sum_scale, stats);
    SCCD_CUDA(cudaGetLastError());
    lc.n += 2;
}

void launch_expand_count(
    const BoxArrays& unsorted, int n, GridParams g, const unsigned long long* d_range,
    uint32_t* copies, cudaStream_t s, LaunchCounter& lc)
{
    if (n <= 0)
        return;
    expand_count_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        unsorted, n, g, d_range, copies);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_cell_splits(
    const BoxArrays& unsorted, int n, int stride, GridParams g, int world, uint32_t* hist,
    unsigned long long* out, cudaStream_t s, LaunchCounter& lc)
{
    const int cells = g.sy * g.sz;
    SCCD_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint32_t) * (size_t)cells, s));
    if (n > 0) {
        const int samples = (n + stride - 1) / stride;
        cell_hist_kernel<<<(samples + kThreads - 1) / kThreads, kThreads, 0, s>>>(
            unsorted, n, stride, g, hist);
        SCCD_CUDA(cudaGetLastError());
        lc.n++;
    }
    cell_splits_kernel<<<1, 1024, 0, s>>>(hist, cells, world, out);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_expand_fill(
    const BoxArrays& unsorted, int n, GridParams g, const unsigned long long* offsets,
    unsigned long long* rec, uint32_t* sort_hist, int key_bits, cudaStream_t s, LaunchCounter& lc)
{
    if (n <= 0)
        return;
    // digits of the sort that follows: key bits [kKeyFlagBits, kKeyFlagBits + key_bits) of the
    // record's high word (sort.cu: launch_sort_and_gather)
    const int passes = (key_bits + 7) / 8;
    expand_fill_kernel<true><<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        unsorted, n, g, offsets, 0u, rec, sort_hist, 32 + kKeyFlagBits, passes,
        key_bits - 8 * (passes - 1));
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

} // namespace sccd
