// Cell-striped sort keys for the sweep (north-star item 2: "major-axis projection").
//
// A 1-axis sweep tests every box against ALL boxes whose x interval starts inside its own --
// on a flat cloth that is ~2,400 candidates per box for ~25 real overlaps, and the sweep is
// bound by those tests, not by HBM.  Here each box is also binned into a uniform (y, z) cell
// grid: it is replicated into every cell its closed yz range touches, records are sorted on
// (cell, xmin), and the sweep window of a record only contains records of the same cell.
// A pair is reported in exactly one cell (see GridParams), so the emitted SET is unchanged.
//
// Kernels: box statistics (grid choice), copies-per-box count, (key, index) expansion.
#include "common.cuh"

#include <cfloat>

namespace sccd {

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
    box_stats_kernel(BoxArrays boxes, int n, double* __restrict__ partials)
{
    double mn_y = DBL_MAX, mx_y = -DBL_MAX, mn_z = DBL_MAX, mx_z = -DBL_MAX, sy = 0.0, sz = 0.0;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const double4 b = ldg_d4(&boxes.yz[i]); // (ymin, zmin, ymax, zmax)
        mn_y = fmin(mn_y, b.x);
        mn_z = fmin(mn_z, b.y);
        mx_y = fmax(mx_y, b.z);
        mx_z = fmax(mx_z, b.w);
        sy += b.z - b.x;
        sz += b.w - b.y;
    }
    __shared__ double red[6][kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn_y = fmin(mn_y, __shfl_xor_sync(0xffffffffu, mn_y, o));
        mx_y = fmax(mx_y, __shfl_xor_sync(0xffffffffu, mx_y, o));
        mn_z = fmin(mn_z, __shfl_xor_sync(0xffffffffu, mn_z, o));
        mx_z = fmax(mx_z, __shfl_xor_sync(0xffffffffu, mx_z, o));
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) {
        red[0][warp] = mn_y;
        red[1][warp] = mx_y;
        red[2][warp] = mn_z;
        red[3][warp] = mx_z;
        red[4][warp] = sy;
        red[5][warp] = sz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; w++) {
            red[0][0] = fmin(red[0][0], red[0][w]);
            red[1][0] = fmax(red[1][0], red[1][w]);
            red[2][0] = fmin(red[2][0], red[2][w]);
            red[3][0] = fmax(red[3][0], red[3][w]);
            red[4][0] += red[4][w];
            red[5][0] += red[5][w];
        }
        for (int k = 0; k < 6; k++)
            partials[blockIdx.x * 6 + k] = red[k][0];
    }
}

// one warp; fixed reduction tree => deterministic sums
__global__ void box_stats_final_kernel(const double* __restrict__ partials, int blocks, double* out)
{
    const int lane = threadIdx.x;
    double r[6] = { DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX, 0.0, 0.0 };
    for (int b = lane; b < blocks; b += 32) {
        r[0] = fmin(r[0], partials[b * 6 + 0]);
        r[1] = fmax(r[1], partials[b * 6 + 1]);
        r[2] = fmin(r[2], partials[b * 6 + 2]);
        r[3] = fmax(r[3], partials[b * 6 + 3]);
        r[4] += partials[b * 6 + 4];
        r[5] += partials[b * 6 + 5];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r[0] = fmin(r[0], __shfl_xor_sync(0xffffffffu, r[0], o));
        r[1] = fmax(r[1], __shfl_xor_sync(0xffffffffu, r[1], o));
        r[2] = fmin(r[2], __shfl_xor_sync(0xffffffffu, r[2], o));
        r[3] = fmax(r[3], __shfl_xor_sync(0xffffffffu, r[3], o));
        r[4] += __shfl_xor_sync(0xffffffffu, r[4], o);
        r[5] += __shfl_xor_sync(0xffffffffu, r[5], o);
    }
    if (lane == 0)
        for (int k = 0; k < 6; k++)
            out[k] = r[k];
}

__device__ __forceinline__ void cell_range(
    const double4 yz, const GridParams& g, int& y0, int& y1, int& z0, int& z1)
{
    y0 = cell_index(yz.x, g.y0, g.inv_hy, g.sy);
    y1 = cell_index(yz.z, g.y0, g.inv_hy, g.sy);
    z0 = cell_index(yz.y, g.z0, g.inv_hz, g.sz);
    z1 = cell_index(yz.w, g.z0, g.inv_hz, g.sz);
}

__global__ void __launch_bounds__(kThreads)
    expand_count_kernel(BoxArrays boxes, int n, GridParams g, uint32_t* __restrict__ copies)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n)
        return;
    int y0, y1, z0, z1;
    cell_range(ldg_d4(&boxes.yz[i]), g, y0, y1, z0, z1);
    copies[i] = (uint32_t)((y1 - y0 + 1) * (z1 - z0 + 1));
}

__global__ void __launch_bounds__(kThreads) expand_fill_kernel(
    BoxArrays boxes, int n, GridParams g, const unsigned long long* __restrict__ offsets,
    unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n)
        return;
    int y0, y1, z0, z1;
    cell_range(ldg_d4(&boxes.yz[i]), g, y0, y1, z0, z1);
    // x part of the key: min.x rounded DOWN to f32 (conservative for the prefilter)
    const unsigned long long xk = float_to_key(__double2float_rd(__ldg(&boxes.x[i]).x));
    unsigned long long o = offsets[i];
    for (int cy = y0; cy <= y1; cy++)
        for (int cz = z0; cz <= z1; cz++) {
            const unsigned long long cell = (unsigned long long)(cy * g.sz + cz);
            keys[o] = (cell << 32) | xk;
            idx[o] = (uint32_t)i;
            o++;
        }
}

} // namespace

void launch_box_stats(
    const BoxArrays& unsorted, int n, double* partials, double* stats, cudaStream_t s,
    LaunchCounter& lc)
{
    int blocks = (n + kThreads - 1) / kThreads;
    if (blocks > kStatsBlocks)
        blocks = kStatsBlocks;
    if (blocks < 1)
        blocks = 1;
    box_stats_kernel<<<blocks, kThreads, 0, s>>>(unsorted, n, partials);
    SCCD_CUDA(cudaGetLastError());
    box_stats_final_kernel<<<1, 32, 0, s>>>(partials, blocks, stats);
    SCCD_CUDA(cudaGetLastError());
    lc.n += 2;
}

void launch_expand_count(
    const BoxArrays& unsorted, int n, GridParams g, uint32_t* copies, cudaStream_t s,
    LaunchCounter& lc)
{
    if (n <= 0)
        return;
    expand_count_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(unsorted, n, g, copies);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_expand_fill(
    const BoxArrays& unsorted, int n, GridParams g, const unsigned long long* offsets,
    unsigned long long* keys, uint32_t* idx, cudaStream_t s, LaunchCounter& lc)
{
    if (n <= 0)
        return;
    expand_fill_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        unsorted, n, g, offsets, keys, idx);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

} // namespace sccd
