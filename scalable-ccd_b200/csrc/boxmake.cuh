// Device helpers that make the exact box of one entry of a box list from the per-vertex
// boxes (aabb.cu:186-229 build_edge_boxes / build_face_boxes: union of vertex boxes).  Shared by
// the fused mesh kernels (boxes.cu), the sliced multi-GPU build and the gather that REBUILDS the
// exact record of a received (key, box index) record (sort.cu).
#pragma once

#include "common.cuh"

namespace sccd {

// replicated per-vertex data + topology: everything a rank needs to rebuild any box of the mesh
struct MeshView {
    const double* vbox = nullptr; // 6 * nV: min xyz, max xyz
    const int32_t* E = nullptr;
    const int32_t* F = nullptr;
    int nV = 0, nE = 0, nF = 0;
};

#ifdef __CUDACC__
__device__ __forceinline__ void load_vbox(
    const double* __restrict__ vbox, int v, double lo[3], double hi[3])
{
    const double2* vb = reinterpret_cast<const double2*>(vbox + (size_t)6 * v);
    const double2 a = __ldg(vb), b = __ldg(vb + 1), c = __ldg(vb + 2);
    lo[0] = a.x;
    lo[1] = a.y;
    lo[2] = b.x;
    hi[0] = b.y;
    hi[1] = c.x;
    hi[2] = c.y;
}

// The sweep runs along the FIRST coordinate of a record: a list swept along `axis` stores its
// boxes with the axes rotated to (axis, axis + 1, axis + 2) mod 3 (SCCD_OPT_SWEEP_AXIS; the
// reference's GPU path always sorts on x, aabb.cu:86, its CPU path on the caller's axis,
// sort_and_sweep.cpp:78-116).  The overlap set does not depend on it.
__device__ __forceinline__ double pick_axis(const double v[3], int a)
{
    return a == 0 ? v[0] : (a == 1 ? v[1] : v[2]);
}
__device__ __forceinline__ void rotate_box(
    const double lo[3], const double hi[3], int axis, double2& x, double4& yz)
{
    const int ay = axis == 2 ? 0 : axis + 1, az = axis == 0 ? 2 : axis - 1;
    x = make_double2(pick_axis(lo, axis), pick_axis(hi, axis));
    yz = make_double4(pick_axis(lo, ay), pick_axis(lo, az), pick_axis(hi, ay), pick_axis(hi, az));
}
__device__ __forceinline__ void store_record(
    const BoxArrays& out, size_t k, const double lo[3], const double hi[3], int4 id, int axis)
{
    double2 x;
    double4 yz;
    rotate_box(lo, hi, axis, x, yz);
    out.x[k] = x;
    out.yz[k] = yz;
    out.id[k] = id;
}

// Box `idx` of list 0 (vertices [0, nV), then faces) or list 1 (edges), ids as the reference
// sets them (aabb.cu:180-181, 199-203, 221-226; vertex element ids flipped, broad_phase.cu:20-26).
// false: an E / F entry is not a vertex index (the reference would read out of bounds).
__device__ __forceinline__ bool make_list_box(
    const MeshView& m, int list, int idx, double lo[3], double hi[3], int4& id)
{
    if (list == 0 && idx < m.nV) {
        load_vbox(m.vbox, idx, lo, hi);
        id = make_int4(idx, -idx - 1, -idx - 1, -idx - 1);
        return true;
    }
    int v0, v1, v2 = -1;
    if (list == 0) {
        const int f = idx - m.nV;
        v0 = __ldg(m.F + f), v1 = __ldg(m.F + f + (size_t)m.nF), v2 = __ldg(m.F + f + (size_t)2 * m.nF);
        id = make_int4(v0, v1, v2, f);
    } else {
        v0 = __ldg(m.E + idx), v1 = __ldg(m.E + idx + (size_t)m.nE);
        id = make_int4(v0, v1, -v0 - 1, idx);
    }
    if ((unsigned)v0 >= (unsigned)m.nV || (unsigned)v1 >= (unsigned)m.nV
        || (list == 0 && (unsigned)v2 >= (unsigned)m.nV)) {
        lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.0;
        return false;
    }
    double lo1[3], hi1[3];
    load_vbox(m.vbox, v0, lo, hi);
    load_vbox(m.vbox, v1, lo1, hi1);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        lo[k] = fmin(lo[k], lo1[k]);
        hi[k] = fmax(hi[k], hi1[k]);
    }
    if (list == 0) {
        load_vbox(m.vbox, v2, lo1, hi1);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fmin(lo[k], lo1[k]);
            hi[k] = fmax(hi[k], hi1[k]);
        }
    }
    return true;
}
#endif

} // namespace sccd
