// AABB construction on the device (north-star item 1).
//
// Replaces the reference's host/TBB box builders + 64 B/box AoS upload + split_boxes
// kernel (cuda/broad_phase/aabb.cu:26-35, 40-72, 115-229) with two coalesced kernels that
// read the two-frame vertex / edge / face buffers directly and emit, per box, the exact
// 64-byte record (x interval, yz mini-box, ids); the sort keys are made by csrc/grid.cu.
// Results are bit-identical to the reference's boxes: nextafter() in double is exact on
// the device and min/max/add are correctly rounded.
#include "boxmake.cuh"

#include <cfloat>

namespace sccd {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double next_down(double x) { return nextafter(x, -DBL_MAX); }
__device__ __forceinline__ double next_up(double x) { return nextafter(x, DBL_MAX); }
// float build of the reference (scalar.hpp:31-49 with Scalar = float)
__device__ __forceinline__ float next_down(float x) { return nextafterf(x, -FLT_MAX); }
__device__ __forceinline__ float next_up(float x) { return nextafterf(x, FLT_MAX); }

// aabb.cu:146-184 build_vertex_boxes(V0, V1, r) + from_point + conservative_inflation.
// F32: the reference's float build -- the vertices are cast to float first (aabb.cu:124-128,
// round to nearest), the box is made with nextafterf and float adds, and everything downstream
// (vertex table, exact records) holds those float values widened to double, which is exact and
// keeps every comparison of the sweep unchanged.  This file is compiled without -ftz, like the
// reference's host code that makes its boxes.
template <bool F32>
__global__ void __launch_bounds__(kThreads) vertex_boxes_kernel(
    const double* __restrict__ V0, const double* __restrict__ V1, int nV, double radius_up,
    VertexRec* __restrict__ vtab, double* __restrict__ vbox, BoxArrays vf, int axis_vf)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= nV)
        return;
    double p0[3], p1[3], lo[3], hi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        p0[k] = V0[i + (size_t)k * nV];
        p1[k] = V1[i + (size_t)k * nV];
        if (F32) {
            const float q0 = __double2float_rn(p0[k]), q1 = __double2float_rn(p1[k]);
            const float r = (float)radius_up; // already a float value (made on the host)
            const float a_lo = __fsub_rn(next_down(q0), r);
            const float b_lo = __fsub_rn(next_down(q1), r);
            const float a_hi = __fadd_rn(next_up(q0), r);
            const float b_hi = __fadd_rn(next_up(q1), r);
            lo[k] = (double)fminf(a_lo, b_lo);
            hi[k] = (double)fmaxf(a_hi, b_hi);
            p0[k] = (double)q0;
            p1[k] = (double)q1;
        } else {
            // aabb.cu:29-34 on each end point, then AABB(a, b) = componentwise min / max
            const double a_lo = __dsub_rn(next_down(p0[k]), radius_up);
            const double b_lo = __dsub_rn(next_down(p1[k]), radius_up);
            const double a_hi = __dadd_rn(next_up(p0[k]), radius_up);
            const double b_hi = __dadd_rn(next_up(p1[k]), radius_up);
            lo[k] = fmin(a_lo, b_lo);
            hi[k] = fmax(a_hi, b_hi);
        }
    }
    VertexRec r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.p0[k] = p0[k];
        r.p1[k] = p1[k];
    }
    vtab[i] = r;
    double2* vb = reinterpret_cast<double2*>(vbox + (size_t)6 * i);
    vb[0] = make_double2(lo[0], lo[1]);
    vb[1] = make_double2(lo[2], hi[0]);
    vb[2] = make_double2(hi[1], hi[2]);
    // aabb.cu:180-181 ids; element id flipped because vertices are list A of the
    // vertex-face sweep (broad_phase.cu:20-26).
    if (vf.x) // (the sliced multi-GPU build makes its records from vbox: list_boxes_kernel)
        store_record(vf, i, lo, hi, make_int4(i, -i - 1, -i - 1, -i - 1), axis_vf);
}

// aabb.cu:186-229 build_edge_boxes / build_face_boxes (union of vertex boxes).
__global__ void __launch_bounds__(kThreads) element_boxes_kernel(
    const double* __restrict__ vbox, const int32_t* __restrict__ E, int nE,
    const int32_t* __restrict__ F, int nF, int nV, BoxArrays eb, BoxArrays vf, int axis_e,
    int axis_vf, int* __restrict__ bad)
{
    const int t = blockIdx.x * kThreads + threadIdx.x;
    if (t < nE) {
        const int e0 = E[t], e1 = E[t + (size_t)nE];
        double lo[3], hi[3], lo1[3], hi1[3];
        if ((unsigned)e0 >= (unsigned)nV || (unsigned)e1 >= (unsigned)nV) {
            // the reference would index out of bounds; reported as SCCD_ERR_ARG by build_boxes
            *bad = 1;
            const double z[3] = { 0.0, 0.0, 0.0 };
            store_record(eb, t, z, z, make_int4(0, 0, -1, t), axis_e);
            return;
        }
        load_vbox(vbox, e0, lo, hi);
        load_vbox(vbox, e1, lo1, hi1);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fmin(lo[k], lo1[k]);
            hi[k] = fmax(hi[k], hi1[k]);
        }
        store_record(eb, t, lo, hi, make_int4(e0, e1, -e0 - 1, t), axis_e);
    } else if (t < nE + nF) {
        const int f = t - nE;
        const int f0 = F[f], f1 = F[f + (size_t)nF], f2 = F[f + (size_t)2 * nF];
        double lo[3], hi[3], lo1[3], hi1[3], lo2[3], hi2[3];
        if ((unsigned)f0 >= (unsigned)nV || (unsigned)f1 >= (unsigned)nV
            || (unsigned)f2 >= (unsigned)nV) {
            *bad = 1;
            const double z[3] = { 0.0, 0.0, 0.0 };
            store_record(vf, nV + f, z, z, make_int4(0, 0, 0, f), axis_vf);
            return;
        }
        load_vbox(vbox, f0, lo, hi);
        load_vbox(vbox, f1, lo1, hi1);
        load_vbox(vbox, f2, lo2, hi2);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fmin(fmin(lo[k], lo1[k]), lo2[k]);
            hi[k] = fmax(fmax(hi[k], hi1[k]), hi2[k]);
        }
        store_record(vf, nV + f, lo, hi, make_int4(f0, f1, f2, f), axis_vf);
    }
}

// Boxes first, first + stride, ... (count of them) of a list, made from the replicated vertex
// boxes into a compact array: the multi-GPU build's SLICE of a list (stride 1) and the sample
// every rank takes of the whole list for the grid statistics and the cell histogram.
__global__ void __launch_bounds__(kThreads) list_boxes_kernel(
    MeshView m, int list, long long first, int stride, int count, BoxArrays out, int axis,
    int* __restrict__ bad)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= count)
        return;
    double lo[3], hi[3];
    int4 id;
    if (!make_list_box(m, list, (int)(first + (long long)j * stride), lo, hi, id))
        *bad = 1;
    store_record(out, j, lo, hi, id, axis);
}

// ---- the reference's three box builders by name, on caller-made arrays ------------------
// (cuda/broad_phase/aabb.cuh:150-188).  The pipeline uses the fused mesh kernels above; these
// serve callers that build and keep the boxes themselves, as tests/test_broad_phase.cu:88-91 does.
// AoS in and out (the reference's 64-byte cuda::AABB).
template <bool F32>
__global__ void __launch_bounds__(kThreads) vertex_aabb_kernel(
    const double* __restrict__ V0, const double* __restrict__ V1, int nV, double radius_up,
    sccd_aabb* __restrict__ out)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= nV)
        return;
    sccd_aabb b;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double p0 = V0[i + (size_t)k * nV], p1 = V1[i + (size_t)k * nV];
        if (F32) {
            const float q0 = __double2float_rn(p0), q1 = __double2float_rn(p1);
            const float r = (float)radius_up;
            b.min[k] = (double)fminf(__fsub_rn(next_down(q0), r), __fsub_rn(next_down(q1), r));
            b.max[k] = (double)fmaxf(__fadd_rn(next_up(q0), r), __fadd_rn(next_up(q1), r));
        } else {
            b.min[k] = fmin(__dsub_rn(next_down(p0), radius_up), __dsub_rn(next_down(p1), radius_up));
            b.max[k] = fmax(__dadd_rn(next_up(p0), radius_up), __dadd_rn(next_up(p1), radius_up));
        }
    }
    b.vertex_ids[0] = i, b.vertex_ids[1] = -i - 1, b.vertex_ids[2] = -i - 1; // aabb.cu:140-141
    b.element_id = i;
    out[i] = b;
}

// aabb.cu:186-229: union of 2 (edge) or 3 (face) vertex boxes; idx is n x k column-major
__global__ void __launch_bounds__(kThreads) element_aabb_kernel(
    const sccd_aabb* __restrict__ vb, int nV, const int32_t* __restrict__ idx, int n, int k,
    sccd_aabb* __restrict__ out, int* __restrict__ bad)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n)
        return;
    int v[3];
    v[0] = idx[i], v[1] = idx[i + (size_t)n], v[2] = k == 3 ? idx[i + (size_t)2 * n] : v[0];
    if ((unsigned)v[0] >= (unsigned)nV || (unsigned)v[1] >= (unsigned)nV
        || (unsigned)v[2] >= (unsigned)nV) {
        *bad = 1; // the reference would index out of bounds
        return;
    }
    sccd_aabb b;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double lo = fmin(vb[v[0]].min[c], vb[v[1]].min[c]);
        double hi = fmax(vb[v[0]].max[c], vb[v[1]].max[c]);
        if (k == 3) {
            lo = fmin(lo, vb[v[2]].min[c]);
            hi = fmax(hi, vb[v[2]].max[c]);
        }
        b.min[c] = lo, b.max[c] = hi;
    }
    b.vertex_ids[0] = v[0], b.vertex_ids[1] = v[1];
    b.vertex_ids[2] = k == 3 ? v[2] : -v[0] - 1; // aabb.cu:199-203 / :221-226
    b.element_id = i;
    out[i] = b;
}

// FP64 pipe yardstick for the narrow-phase roofline: 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
           x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = __fma_rn(x0, a, b);
        x1 = __fma_rn(x1, a, b);
        x2 = __fma_rn(x2, a, b);
        x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b);
        x5 = __fma_rn(x5, a, b);
        x6 = __fma_rn(x6, a, b);
        x7 = __fma_rn(x7, a, b);
    }
    const double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (r == 123.456) // never true: keeps the chains alive
        out[0] = r;
}

} // namespace

double measure_dfma_per_second(int num_sms, cudaStream_t s)
{
    double* d = nullptr;
    SCCD_CUDA(cudaMalloc(&d, 8));
    cudaEvent_t a, b;
    SCCD_CUDA(cudaEventCreate(&a));
    SCCD_CUDA(cudaEventCreate(&b));
    const int iters = 1 << 14, grid = num_sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        SCCD_CUDA(cudaEventRecord(a, s));
        dfma_peak_kernel<<<grid, 256, 0, s>>>(d, iters, 0.999999, 1e-9);
        SCCD_CUDA(cudaEventRecord(b, s));
        SCCD_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        SCCD_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best)
            best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d);
    // thread-level DFMA per second
    return (double)grid * 256.0 * 8.0 * iters / (best * 1e-3);
}

void launch_vertex_aabbs(
    const double* V0, const double* V1, int nV, double radius_up, bool f32, sccd_aabb* out,
    cudaStream_t s, LaunchCounter& lc)
{
    if (nV <= 0)
        return;
    const int grid = (nV + kThreads - 1) / kThreads;
    if (f32)
        vertex_aabb_kernel<true><<<grid, kThreads, 0, s>>>(V0, V1, nV, radius_up, out);
    else
        vertex_aabb_kernel<false><<<grid, kThreads, 0, s>>>(V0, V1, nV, radius_up, out);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_element_aabbs(
    const sccd_aabb* vb, int nV, const int32_t* idx, int n, int k, sccd_aabb* out, int* bad,
    cudaStream_t s, LaunchCounter& lc)
{
    if (n <= 0)
        return;
    element_aabb_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        vb, nV, idx, n, k, out, bad);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_list_boxes(
    const MeshView& m, int list, long long first, int stride, int count, BoxArrays out, int axis,
    int* bad, cudaStream_t s, LaunchCounter& lc)
{
    if (count <= 0)
        return;
    list_boxes_kernel<<<(count + kThreads - 1) / kThreads, kThreads, 0, s>>>(
        m, list, first, stride, count, out, axis, bad);
    SCCD_CUDA(cudaGetLastError());
    lc.n++;
}

void launch_mesh_boxes(
    const double* V0, const double* V1, int nV, double radius_up, bool f32, VertexRec* vtab,
    double* vbox, const int32_t* E, int nE, const int32_t* F, int nF, BoxArrays e_unsorted,
    BoxArrays vf_unsorted, int axis_e, int axis_vf, int* bad, cudaStream_t s, LaunchCounter& lc)
{
    if (nV > 0) {
        const int grid = (nV + kThreads - 1) / kThreads;
        if (f32)
            vertex_boxes_kernel<true><<<grid, kThreads, 0, s>>>(
                V0, V1, nV, radius_up, vtab, vbox, vf_unsorted, axis_vf);
        else
            vertex_boxes_kernel<false><<<grid, kThreads, 0, s>>>(
                V0, V1, nV, radius_up, vtab, vbox, vf_unsorted, axis_vf);
        SCCD_CUDA(cudaGetLastError());
        lc.n++;
    }
    if (nE + nF > 0) {
        element_boxes_kernel<<<(nE + nF + kThreads - 1) / kThreads, kThreads, 0, s>>>(
            vbox, E, nE, F, nF, nV, e_unsorted, vf_unsorted, axis_e, axis_vf, bad);
        SCCD_CUDA(cudaGetLastError());
        lc.n++;
    }
}

} // namespace sccd
