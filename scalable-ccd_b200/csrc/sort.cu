// Major-axis sort (north-star item 2a).
//
// Replaces thrust::sort_by_key with a comparator on 16 B keys / 48 B values, the D2D copy,
// flip_element_ids and thrust::merge_by_key of the reference
// (cuda/broad_phase/aabb.cu:107-109, broad_phase.cu:57-101) by:
//   1. an LSD radix sort of (u32 key = [cell | quantised min.x | flags] (common.cuh), u32 box
//      index) over only the key bits in use (3-4 digit passes) -- 8 B per record per pass
//      instead of 64 B;
//   2. ONE gather that moves each 64 B exact record to its sorted position and emits the
//      24 B prefilter view (key, reach, f32 yz) the sweep streams.
// The radix sort is stable, so ties keep element order and the result is deterministic.
// The vertex-face list is sorted as one tagged list (vertex element ids are already
// flipped at build time), so no merge is needed.
#include "boxmake.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

namespace sccd {

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) gather_sorted_kernel(
    int m, const uint32_t* __restrict__ sorted_keys, const uint32_t* __restrict__ sorted_idx,
    BoxArrays in, BoxArrays out, PrefilterArrays pf, GridParams g)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= m)
        return;
    const uint32_t src = sorted_idx[j];
    const uint32_t key = sorted_keys[j];
    const double2 x = __ldg(&in.x[src]);
    const double4 yz = ldg_d4(&in.yz[src]);
    const int4 id = __ldg(&in.id[src]);
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
    const uint32_t key = (uint32_t)(r >> 32);
    double lo[3], hi[3];
    int4 id;
    make_list_box(mesh, list, (int)(uint32_t)r, lo, hi, id); // (indices were checked by the sender)
    double2 x;
    double4 yz;
    rotate_box(lo, hi, axis, x, yz);
    out.x[j] = x;
    out.yz[j] = yz;
    out.id[j] = id;
    pf.key[j] = key;
    const int cell_shift = g.x_bits + kKeyFlagBits;
    const uint32_t cell_part = cell_shift >= 32 ? 0u : (key >> cell_shift) << cell_shift;
    pf.reach[j] = cell_part | (quantize_x(x.y, g) << kKeyFlagBits) | ((1u << kKeyFlagBits) - 1u);
    pf.yz[j] = make_float4(
        __double2float_rd(yz.x), __double2float_ru(yz.z), __double2float_rd(yz.y),
        __double2float_ru(yz.w));
}

struct U32ToU64 {
    __host__ __device__ unsigned long long operator()(uint32_t v) const { return v; }
};
} // namespace

size_t sort_temp_bytes(int n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(
        nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
        (uint32_t*)nullptr, n > 0 ? n : 1, 0, 32);
    return bytes;
}

void launch_sort_and_gather(
    int m, int key_bits, uint32_t* keys_in, uint32_t* keys_out, uint32_t* idx_in,
    uint32_t* idx_out, void* temp, size_t temp_bytes, BoxArrays unsorted, SortedList out,
    cudaStream_t s, LaunchCounter& lc, cudaEvent_t gather_begin, cudaEvent_t gather_end)
{
    if (m <= 0) {
        if (gather_begin)
            SCCD_CUDA(cudaEventRecord(gather_begin, s));
        if (gather_end)
            SCCD_CUDA(cudaEventRecord(gather_end, s));
        return;
    }
    // the flag bits are not part of the order: equal (cell, q) records keep element order
    SCCD_CUDA(cub::DeviceRadixSort::SortPairs(
        temp, temp_bytes, (const uint32_t*)keys_in, keys_out, (const uint32_t*)idx_in, idx_out, m,
        kKeyFlagBits, kKeyFlagBits + key_bits, s));
    lc.n += 2 + (key_bits + 7) / 8; // histogram + exclusive sum + one onesweep pass per digit
    if (gather_begin)
        SCCD_CUDA(cudaEventRecord(gather_begin, s));
    gather_sorted_kernel<<<(m + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        m, keys_out, idx_out, unsorted, out.box, out.pf, out.grid);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
    if (gather_end)
        SCCD_CUDA(cudaEventRecord(gather_end, s));
}

// ---- multi-GPU (shard.cu) ---------------------------------------------------------------
size_t partition_temp_bytes(long long m)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(
        nullptr, bytes, (const uint8_t*)nullptr, (uint8_t*)nullptr,
        (const unsigned long long*)nullptr, (unsigned long long*)nullptr, m > 0 ? m : 1, 0, 5);
    return bytes;
}

// stable partition of the records by destination rank (one digit pass)
void launch_partition_by_dest(
    long long m, const uint8_t* dest_in, uint8_t* dest_out, const unsigned long long* rec_in,
    unsigned long long* rec_out, void* temp, size_t temp_bytes, cudaStream_t s, LaunchCounter& lc)
{
    if (m <= 0)
        return;
    SCCD_CUDA(cub::DeviceRadixSort::SortPairs(
        temp, temp_bytes, dest_in, dest_out, rec_in, rec_out, m, 0, 5, s));
    lc.n += 3;
}

size_t sort_records_temp_bytes(long long m)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortKeys(
        nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
        m > 0 ? m : 1, 32, 64);
    return bytes;
}

// sorts m received records on key bits [kKeyFlagBits, kKeyFlagBits + key_bits) of their high
// word (stable: equal keys keep arrival order = global box order) and rebuilds the sorted views
void launch_sort_records_and_rebuild(
    int m, int key_bits, const unsigned long long* rec_in, unsigned long long* rec_out, void* temp,
    size_t temp_bytes, const MeshView& mesh, int list, int axis, SortedList out, cudaStream_t s,
    LaunchCounter& lc, cudaEvent_t gather_begin, cudaEvent_t gather_end)
{
    if (m > 0) {
        SCCD_CUDA(cub::DeviceRadixSort::SortKeys(
            temp, temp_bytes, rec_in, rec_out, m, 32 + kKeyFlagBits, 32 + kKeyFlagBits + key_bits,
            s));
        lc.n += 2 + (key_bits + 7) / 8;
    }
    if (gather_begin)
        SCCD_CUDA(cudaEventRecord(gather_begin, s));
    if (m > 0) {
        gather_rebuild_kernel<<<(m + kThreads - 1) / kThreads, kThreads, 0, s>>>(
            m, rec_out, mesh, list, axis, out.box, out.pf, out.grid);
        SCCD_CUDA(cudaGetLastError());
        lc.n++;
    }
    if (gather_end)
        SCCD_CUDA(cudaEventRecord(gather_end, s));
}

size_t scan_temp_bytes(int n)
{
    size_t bytes = 0;
    cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> it(
        nullptr, U32ToU64());
    cub::DeviceScan::ExclusiveSum(
        nullptr, bytes, it, (unsigned long long*)nullptr, n > 0 ? n + 1 : 1);
    return bytes;
}

void launch_scan_u32_to_u64(
    const uint32_t* counts, unsigned long long* offsets, int n, void* temp,
    size_t temp_bytes, cudaStream_t s, LaunchCounter& lc)
{
    // counts has n+1 readable entries (the last one is zero) so that offsets[n] = total.
    cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> it(
        counts, U32ToU64());
    SCCD_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, it, offsets, n + 1, s));
    lc.n += 2;
}

} // namespace sccd
