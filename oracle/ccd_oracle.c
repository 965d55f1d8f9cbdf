/*
 * ccd_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by, or
 * called from the product path).
 *
 * A plain-C, single-threaded-by-default CPU restatement of the Scalable-CCD hot
 * path (AABB build -> sort-and-sweep broad phase -> Tight-Inclusion narrow
 * phase), used as the parity checker for the CUDA kernels in
 * scalable-ccd_b200/csrc and as the "port" CPU baseline in bench.py.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.  Arithmetic follows the contract in SURVEY.md
 * section 8a: IEEE double, FMA exactly where nvcc contracts the reference's
 * expressions (checked against the reference's own SASS), IEEE division.
 * Compile with -ffp-contract=off so that the ONLY fused operations are the
 * explicit fma() calls below.
 *
 * Parity pinning: the broad-phase half is pinned against the unmodified
 * reference CPU sources built by oracle/Makefile into oracle/_ref/ (run in the
 * build container, see tests/test_oracle_vs_ref.py); the narrow-phase half is
 * pinned against the unmodified reference CUDA sources (oracle/_ref/
 * libref_sccd_cuda.so) run on a B200 and frozen as tests/golden/ fixtures.
 *
 * Two builds of this file (oracle/Makefile): liborc.so restates the reference's
 * default double build; liborc_f32.so (-DORC_F32) restates its float build
 * (SCALABLE_CCD_USE_DOUBLE off, scalar.hpp:16-18).  The C interface is the same
 * -- double arrays in and out -- and in the float build every value that crosses
 * it is a float widened to double.  Float narrow phase: the reference compiles
 * its CUDA code with --use_fast_math (CMakeLists.txt:219-230), i.e. flush-to-zero
 * arithmetic and a / b = a * rcp.approx(b).  FTZ is restated exactly; the
 * hardware reciprocal (MUFU.RCP, <= 1 ulp) cannot be, so rcp() below is the
 * correctly rounded reciprocal and the float narrow phase of this oracle is in
 * principle a TOLERANCE-level model of the reference (tolerances and split choices
 * can differ in the last ulp); the float broad phase is exact.  Pinned: on every
 * committed fixture frozen from the reference's float CUDA build on a B200
 * (tests/golden, files named ..._f32...: 1,489 collisions of config 1 with their TOIs, 3,424
 * adversarial queries) the two agree bit for bit (tests/test_f32.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Same 64-byte layout as the reference's cuda::AABB (double build):
 * cuda/broad_phase/aabb.cuh:12-96 -- Scalar3 min, Scalar3 max, int3
 * vertex_ids, int element_id. */
typedef struct {
    double min[3];
    double max[3];
    int32_t vertex_ids[3];
    int32_t element_id;
} orc_aabb;

/* ---- the reference's Scalar (scalar.hpp:13-19) ------------------------------ */
#ifdef ORC_F32
typedef float real;
#define REAL_MAX FLT_MAX
#define REAL_EPSILON FLT_EPSILON
#define r_nextafter nextafterf
#define r_fma fmaf
#define r_abs fabsf
#define r_min fminf
#define r_max fmaxf
/* flush-to-zero of a result, sign kept (every device op of the float build is .FTZ) */
static inline real ftz(real x) { return (x != 0 && r_abs(x) < FLT_MIN) ? copysignf(0.0f, x) : x; }
/* a / b as the float build computes it: a * rcp(b) (MUFU.RCP + FMUL.FTZ) */
static inline real r_div(real a, real b)
{
    const real rcp = ftz((real)(1.0 / (double)ftz(b)));
    return ftz(ftz(a) * rcp);
}
#else
typedef double real;
#define REAL_MAX DBL_MAX
#define REAL_EPSILON DBL_EPSILON
#define r_nextafter nextafter
#define r_fma fma
#define r_abs fabs
#define r_min fmin
#define r_max fmax
static inline real ftz(real x) { return x; }
static inline real r_div(real a, real b) { return a / b; }
#endif
int orc_is_f32(void) { return sizeof(real) == 4; }

/* scalar.hpp:31-49 */
static inline real nextafter_down(real x) { return r_nextafter(x, -REAL_MAX); }
static inline real nextafter_up(real x) { return r_nextafter(x, REAL_MAX); }

/* cuda/broad_phase/aabb.cu:19-35 (from_point + conservative_inflation),
 * CPU twin broad_phase/aabb.cpp:17-36.  Host code in the reference: no FTZ.
 * p is cast to Scalar first in the float build (aabb.cu:124-128). */
static inline void point_box(const double p[3], double r, real mn[3], real mx[3])
{
    const real ru = nextafter_up((real)r);
    for (int k = 0; k < 3; k++) {
        const real q = (real)p[k];
        mn[k] = nextafter_down(q) - ru;
        mx[k] = nextafter_up(q) + ru;
    }
}

/* cuda/broad_phase/aabb.cu:146-184: build_vertex_boxes(V0, V1, r).
 * V0/V1 are nV x 3 COLUMN-major (Eigen default; device_matrix.cuh:77-81). */
void orc_build_vertex_boxes(
    const double* V0, const double* V1, int64_t nV, double r, orc_aabb* out)
{
    for (int64_t i = 0; i < nV; i++) {
        double p0[3] = { V0[i], V0[i + nV], V0[i + 2 * nV] };
        double p1[3] = { V1[i], V1[i + nV], V1[i + 2 * nV] };
        real a0[3], a1[3], b0[3], b1[3];
        point_box(p0, r, a0, a1);
        point_box(p1, r, b0, b1);
        for (int k = 0; k < 3; k++) {
            /* aabb.cuh:18-28: AABB(a, b) = (min of mins, max of maxs) */
            out[i].min[k] = a0[k] < b0[k] ? a0[k] : b0[k];
            out[i].max[k] = a1[k] > b1[k] ? a1[k] : b1[k];
        }
        /* aabb.cu:180-181 */
        out[i].vertex_ids[0] = (int32_t)i;
        out[i].vertex_ids[1] = (int32_t)(-i - 1);
        out[i].vertex_ids[2] = (int32_t)(-i - 1);
        out[i].element_id = (int32_t)i;
    }
}

/* cuda/broad_phase/aabb.cu:186-206.  E is nE x 2 column-major int32. */
void orc_build_edge_boxes(
    const orc_aabb* vb, const int32_t* E, int64_t nE, orc_aabb* out)
{
    for (int64_t i = 0; i < nE; i++) {
        const int32_t e0 = E[i], e1 = E[i + nE];
        for (int k = 0; k < 3; k++) {
            out[i].min[k] = fmin(vb[e0].min[k], vb[e1].min[k]);
            out[i].max[k] = fmax(vb[e0].max[k], vb[e1].max[k]);
        }
        out[i].vertex_ids[0] = e0;
        out[i].vertex_ids[1] = e1;
        out[i].vertex_ids[2] = -e0 - 1;
        out[i].element_id = (int32_t)i;
    }
}

/* cuda/broad_phase/aabb.cu:208-229.  F is nF x 3 column-major int32. */
void orc_build_face_boxes(
    const orc_aabb* vb, const int32_t* F, int64_t nF, orc_aabb* out)
{
    for (int64_t i = 0; i < nF; i++) {
        const int32_t f0 = F[i], f1 = F[i + nF], f2 = F[i + 2 * nF];
        for (int k = 0; k < 3; k++) {
            out[i].min[k] = fmin(fmin(vb[f0].min[k], vb[f1].min[k]), vb[f2].min[k]);
            out[i].max[k] = fmax(fmax(vb[f0].max[k], vb[f1].max[k]), vb[f2].max[k]);
        }
        out[i].vertex_ids[0] = f0;
        out[i].vertex_ids[1] = f1;
        out[i].vertex_ids[2] = f2;
        out[i].element_id = (int32_t)i;
    }
}

/* ------------------------------------------------------------------------- */
/* Broad phase                                                               */

/* cuda/broad_phase/collision.cuh:17-21, broad_phase/sort_and_sweep.cpp:22-28 */
static inline int share_a_vertex(const int32_t* a, const int32_t* b)
{
    return a[0] == b[0] || a[0] == b[1] || a[0] == b[2] || a[1] == b[0]
        || a[1] == b[1] || a[1] == b[2] || a[2] == b[0] || a[2] == b[1]
        || a[2] == b[2];
}

/* broad_phase/aabb.cpp:24-29 and cuda/broad_phase/aabb.cuh:67-72 (closed). */
static inline int intersects(const orc_aabb* a, const orc_aabb* b)
{
    return a->min[0] <= b->max[0] && b->min[0] <= a->max[0]
        && a->min[1] <= b->max[1] && b->min[1] <= a->max[1]
        && a->min[2] <= b->max[2] && b->min[2] <= a->max[2];
}

typedef struct {
    int32_t a, b;
} orc_pair;

static int g_axis;
static int cmp_min_axis(const void* pa, const void* pb)
{
    const double a = ((const orc_aabb*)pa)->min[g_axis];
    const double b = ((const orc_aabb*)pb)->min[g_axis];
    return (a > b) - (a < b);
}

/* broad_phase/sort_and_sweep.cpp:176-195: next sort axis = argmax of
 * sum(c^2) - sum(c)^2 / n over box centres (serial accumulation order). */
static int next_axis(const orc_aabb* boxes, int64_t n)
{
    real s[3] = { 0, 0, 0 }, s2[3] = { 0, 0, 0 };
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            const real c = ((real)boxes[i].min[k] + (real)boxes[i].max[k]) / 2;
            s[k] += c;
            s2[k] += c * c;
        }
    real var[3];
    for (int k = 0; k < 3; k++)
        var[k] = s2[k] - s[k] * s[k] / (real)n;
    int ax = 0;
    if (var[1] > var[0])
        ax = 1;
    if (var[2] > var[ax])
        ax = 2;
    return ax;
}

/* broad_phase/sort_and_sweep.cpp:78-125 (batched_sweep) over boxes already
 * sorted on min[axis]; two-list mode expects list-A ids already flipped
 * (-id-1), as sort_and_sweep.cpp:229-231 / cuda broad_phase.cu:20-26 do.
 * Emits (vertex, face) in two-list mode and (min id, max id) otherwise --
 * identical to cuda/broad_phase/sweep.cu:152-164.
 * Returns the number of overlaps found; writes at most cap of them. */
static int64_t sweep_sorted(
    const orc_aabb* boxes, int64_t n, int axis, int two_lists, orc_pair* out,
    int64_t cap)
{
    int64_t count = 0;
    for (int64_t i = 0; i < n; i++) {
        const orc_aabb* a = &boxes[i];
        for (int64_t j = i + 1; j < n; j++) {
            const orc_aabb* b = &boxes[j];
            if (a->max[axis] < b->min[axis])
                break;
            if (two_lists
                && !((a->element_id >= 0 && b->element_id < 0)
                     || (a->element_id < 0 && b->element_id >= 0)))
                continue;
            if (!intersects(a, b) || share_a_vertex(a->vertex_ids, b->vertex_ids))
                continue;
            if (count < cap) {
                if (two_lists) {
                    out[count].a = a->element_id < 0 ? -a->element_id - 1
                                                     : -b->element_id - 1;
                    out[count].b = a->element_id < 0 ? b->element_id : a->element_id;
                } else {
                    out[count].a = a->element_id < b->element_id ? a->element_id
                                                                 : b->element_id;
                    out[count].b = a->element_id < b->element_id ? b->element_id
                                                                 : a->element_id;
                }
            }
            count++;
        }
    }
    return count;
}

/* broad_phase/sort_and_sweep.cpp:198-211 (single list). */
int64_t orc_sort_and_sweep(
    const orc_aabb* boxes_in, int64_t n, int* sort_axis, orc_pair* out, int64_t cap)
{
    if (n == 0)
        return 0;
    orc_aabb* boxes = (orc_aabb*)malloc(sizeof(orc_aabb) * (size_t)n);
    memcpy(boxes, boxes_in, sizeof(orc_aabb) * (size_t)n);
    g_axis = *sort_axis;
    qsort(boxes, (size_t)n, sizeof(orc_aabb), cmp_min_axis);
    const int64_t c = sweep_sorted(boxes, n, *sort_axis, 0, out, cap);
    *sort_axis = next_axis(boxes, n);
    free(boxes);
    return c;
}

/* broad_phase/sort_and_sweep.cpp:213-239 (two lists: A = vertices, B = faces). */
int64_t orc_sort_and_sweep_two_lists(
    const orc_aabb* A, int64_t nA, const orc_aabb* B, int64_t nB, int* sort_axis,
    orc_pair* out, int64_t cap)
{
    if (nA == 0 || nB == 0)
        return 0;
    const int64_t n = nA + nB;
    orc_aabb* boxes = (orc_aabb*)malloc(sizeof(orc_aabb) * (size_t)n);
    memcpy(boxes, A, sizeof(orc_aabb) * (size_t)nA);
    memcpy(boxes + nA, B, sizeof(orc_aabb) * (size_t)nB);
    for (int64_t i = 0; i < nA; i++)
        boxes[i].element_id = -boxes[i].element_id - 1;
    g_axis = *sort_axis;
    qsort(boxes, (size_t)n, sizeof(orc_aabb), cmp_min_axis);
    const int64_t c = sweep_sorted(boxes, n, *sort_axis, 1, out, cap);
    *sort_axis = next_axis(boxes, n);
    free(boxes);
    return c;
}

/* O(n^2) definition of the overlap set: all closed-
 * intersecting, type-valid, non-incident pairs.  Small n only. */
int64_t orc_brute_force(
    const orc_aabb* A, int64_t nA, const orc_aabb* B, int64_t nB, orc_pair* out,
    int64_t cap)
{
    int64_t count = 0;
    if (B) {
        for (int64_t i = 0; i < nA; i++)
            for (int64_t j = 0; j < nB; j++)
                if (intersects(&A[i], &B[j])
                    && !share_a_vertex(A[i].vertex_ids, B[j].vertex_ids)) {
                    if (count < cap) {
                        out[count].a = A[i].element_id;
                        out[count].b = B[j].element_id;
                    }
                    count++;
                }
    } else {
        for (int64_t i = 0; i < nA; i++)
            for (int64_t j = i + 1; j < nA; j++)
                if (intersects(&A[i], &A[j])
                    && !share_a_vertex(A[i].vertex_ids, A[j].vertex_ids)) {
                    if (count < cap) {
                        const int32_t x = A[i].element_id, y = A[j].element_id;
                        out[count].a = x < y ? x : y;
                        out[count].b = x < y ? y : x;
                    }
                    count++;
                }
    }
    return count;
}

/* ------------------------------------------------------------------------- */
/* Narrow phase (Tight-Inclusion, Scalable-CCD GPU variant)                  */

/* Per-query data as it crosses the C interface: the first 192 bytes of the
 * reference's CCDData in the double build (cuda/narrow_phase/ccd_data.cuh:8-26):
 * v0s v1s v2s v3s v0e v1e v2e v3e. */
typedef struct {
    double s[4][3]; /* vertices at t=0 */
    double e[4][3]; /* vertices at t=1 */
} orc_query;

/* ... and in the reference's Scalar: DeviceMatrix<Scalar> / CCDData hold the
 * vertices cast to Scalar (device_matrix.cuh:21-27); the device code flushes
 * subnormal inputs at their first use. */
typedef struct {
    real s[4][3];
    real e[4][3];
} rquery;

static inline void to_rquery(const orc_query* q, rquery* r)
{
    for (int v = 0; v < 4; v++)
        for (int k = 0; k < 3; k++) {
            r->s[v][k] = ftz((real)q->s[v][k]);
            r->e[v][k] = ftz((real)q->e[v][k]);
        }
}

typedef struct {
    real tol[3];
    real err[3];
} orc_bounds;

static inline real linf3(const real* a, const real* b)
{
    /* (b - a).lpNorm<Infinity>() */
    const real x = r_abs(ftz(b[0] - a[0])), y = r_abs(ftz(b[1] - a[1])),
               z = r_abs(ftz(b[2] - a[2]));
    return r_max(r_max(x, y), z);
}

/* cuda/narrow_phase/root_finder.cu:31-46 */
static inline real max_linf_4(
    const real* p1, const real* p2, const real* p3, const real* p4, const real* p1e,
    const real* p2e, const real* p3e, const real* p4e)
{
    return r_max(
        r_max(linf3(p1, p1e), linf3(p2, p2e)), r_max(linf3(p3, p3e), linf3(p4, p4e)));
}

static inline void sub3(const real* a, const real* b, real* r)
{
    r[0] = ftz(a[0] - b[0]);
    r[1] = ftz(a[1] - b[1]);
    r[2] = ftz(a[2] - b[2]);
}

/* cuda/narrow_phase/root_finder.cu:48-88 (tolerances) and :90-135 (error). */
static void compute_bounds(
    const rquery* q, int is_vf, real co_domain_tol, int use_ms, orc_bounds* out)
{
    real p000[3], p001[3], p011[3], p010[3], p100[3], p101[3], p111[3], p110[3];
    if (is_vf) {
        /* root_finder.cu:50-59 */
        real tmp[3];
        sub3(q->s[0], q->s[1], p000);
        sub3(q->s[0], q->s[3], p001);
        for (int k = 0; k < 3; k++)
            tmp[k] = ftz(ftz(q->s[2][k] + q->s[3][k]) - q->s[1][k]);
        sub3(q->s[0], tmp, p011);
        sub3(q->s[0], q->s[2], p010);
        sub3(q->e[0], q->e[1], p100);
        sub3(q->e[0], q->e[3], p101);
        for (int k = 0; k < 3; k++)
            tmp[k] = ftz(ftz(q->e[2][k] + q->e[3][k]) - q->e[1][k]);
        sub3(q->e[0], tmp, p111);
        sub3(q->e[0], q->e[2], p110);
        /* root_finder.cu:61-66 */
        out->tol[0] = r_div(
            co_domain_tol, ftz(3 * max_linf_4(p000, p001, p011, p010, p100, p101, p111, p110)));
        out->tol[1] = r_div(
            co_domain_tol, ftz(3 * max_linf_4(p000, p100, p101, p001, p010, p110, p111, p011)));
        out->tol[2] = r_div(
            co_domain_tol, ftz(3 * max_linf_4(p000, p100, p110, p010, p001, p101, p111, p011)));
    } else {
        /* root_finder.cu:73-80 */
        sub3(q->s[0], q->s[2], p000);
        sub3(q->s[0], q->s[3], p001);
        sub3(q->s[1], q->s[2], p010);
        sub3(q->s[1], q->s[3], p011);
        sub3(q->e[0], q->e[2], p100);
        sub3(q->e[0], q->e[3], p101);
        sub3(q->e[1], q->e[2], p110);
        sub3(q->e[1], q->e[3], p111);
        /* root_finder.cu:82-87 -- tol[1] deliberately equals tol[0] */
        out->tol[0] = r_div(
            co_domain_tol, ftz(3 * max_linf_4(p000, p001, p011, p010, p100, p101, p111, p110)));
        out->tol[1] = out->tol[0];
        out->tol[2] = r_div(
            co_domain_tol, ftz(3 * max_linf_4(p000, p100, p101, p001, p010, p110, p111, p011)));
    }
    /* root_finder.cu:93-122 */
    real filter;
#ifdef ORC_F32
    if (!use_ms)
        filter = is_vf ? 3.576279e-06f : 3.337861e-06f;
    else
        filter = is_vf ? 4.053116e-06f : 3.814698e-06f;
#else
    if (!use_ms)
        filter = is_vf ? 6.661338147750939e-15 : 6.217248937900877e-15;
    else
        filter = is_vf ? 7.549516567451064e-15 : 7.105427357601002e-15;
#endif
    /* root_finder.cu:124-134 */
    for (int k = 0; k < 3; k++) {
        real m = 1;
        for (int v = 0; v < 4; v++) {
            m = r_max(m, r_abs(q->s[v][k]));
            m = r_max(m, r_abs(q->e[v][k]));
        }
        out->err[k] = ftz(ftz(ftz(m * m) * m) * filter);
    }
}

/* cuda/narrow_phase/root_finder.cu:137-155: F at one (t,u,v) corner with the
 * FMA contraction nvcc applies to the reference's expressions. */
static inline void eval_corner(
    const rquery* q, int is_vf, real t, real u, real v, real* r)
{
    for (int k = 0; k < 3; k++) {
        const real a0 = ftz(r_fma(ftz(q->e[0][k] - q->s[0][k]), t, q->s[0][k]));
        const real a1 = ftz(r_fma(ftz(q->e[1][k] - q->s[1][k]), t, q->s[1][k]));
        const real a2 = ftz(r_fma(ftz(q->e[2][k] - q->s[2][k]), t, q->s[2][k]));
        const real a3 = ftz(r_fma(ftz(q->e[3][k] - q->s[3][k]), t, q->s[3][k]));
        if (is_vf) {
            /* v - (t1-t0)*u - (t2-t0)*v - t0, root_finder.cu:144 */
            real x = ftz(r_fma(-ftz(a2 - a1), u, a0));
            x = ftz(r_fma(-ftz(a3 - a1), v, x));
            r[k] = ftz(x - a1);
        } else {
            /* ((ea1-ea0)*u+ea0) - ((eb1-eb0)*v+eb0), root_finder.cu:154 */
            const real x = ftz(r_fma(ftz(a1 - a0), u, a0));
            const real y = ftz(r_fma(ftz(a3 - a2), v, a2));
            r[k] = ftz(x - y);
        }
    }
}

typedef struct {
    real lo[3], hi[3];
} orc_box;

/* cuda/narrow_phase/root_finder.cu:157-198 */
static inline int origin_in_inclusion(
    const rquery* q, const orc_bounds* b, int is_vf, real ms, const orc_box* box,
    real* true_tol, int* box_in)
{
    real cmin[3] = { REAL_MAX, REAL_MAX, REAL_MAX };
    real cmax[3] = { -REAL_MAX, -REAL_MAX, -REAL_MAX };
    for (int c = 0; c < 8; c++) {
        /* interval.cuh:52-57: bit0 -> t, bit1 -> u, bit2 -> v */
        const real t = (c & 1) ? box->hi[0] : box->lo[0];
        const real u = (c & 2) ? box->hi[1] : box->lo[1];
        const real v = (c & 4) ? box->hi[2] : box->lo[2];
        real r[3];
        eval_corner(q, is_vf, t, u, v, r);
        for (int k = 0; k < 3; k++) {
            cmin[k] = r_min(cmin[k], r[k]);
            cmax[k] = r_max(cmax[k], r[k]);
        }
    }
    const real w0 = ftz(cmax[0] - cmin[0]), w1 = ftz(cmax[1] - cmin[1]),
               w2 = ftz(cmax[2] - cmin[2]);
    *true_tol = r_max(0, r_max(r_max(w0, w1), w2));
    *box_in = 1;
    for (int k = 0; k < 3; k++)
        if (ftz(cmin[k] - ms) > b->err[k] || ftz(cmax[k] + ms) < -b->err[k])
            return 0;
    for (int k = 0; k < 3; k++)
        if (ftz(cmin[k] + ms) < -b->err[k] || ftz(cmax[k] - ms) > b->err[k])
            *box_in = 0;
    return 1;
}

/* cuda/narrow_phase/root_finder.cu:200-211 */
static inline int split_dimension(const orc_bounds* b, const real* w)
{
    const real r0 = r_div(w[0], b->tol[0]), r1 = r_div(w[1], b->tol[1]),
               r2 = r_div(w[2], b->tol[2]);
    if (r0 >= r1 && r0 >= r2)
        return 0;
    if (r1 >= r0 && r1 >= r2)
        return 1;
    return 2;
}

/* Statistics the bench reports (box checks = ccd_kernel invocations that
 * reach origin_in_inclusion_function). */
typedef struct {
    int64_t box_checks;
    int64_t max_stack;
    int64_t capped_queries;
} orc_np_stats;

#define ORC_STACK_MAX 4096

/*
 * One query, depth-first, earliest-t child first.  Restates
 * cuda/narrow_phase/root_finder.cu:277-370 (ccd_kernel) + :213-254 (bisect).
 * With max_iter < 0 the accepted minimum is independent of traversal order
 * (SURVEY.md 8a "arithmetic contract"), so DFS == the reference's BFS.
 *
 * prune_toi: pointer to the bound used for pruning and lowered on accept
 *   (the per-query toi in TOI_PER_QUERY mode, narrow_phase.cu:69-71 /
 *   root_finder.cu:296-297; the global toi otherwise, root_finder.cu:295).
 * max_iter >= 0: the reference silently DROPS boxes once the racy per-query
 *   counter passes the cap (root_finder.cu:288-305).  cap_mode 0 restates
 *   that (serialised, depth-first: non-conservative, like the reference);
 *   cap_mode 1 is the conservative rule the B200 path uses: a box popped after
 *   the cap is ACCEPTED at its t_lo, which can only make the answer earlier.
 */
static int64_t solve_query(
    const orc_query* q_in, int is_vf, double ms_in, int max_iter, double co_tol_in,
    int allow_zero_toi, int cap_mode, double* prune_toi, double* global_toi,
    orc_np_stats* st)
{
    /* Scalar parameters of the reference's entry points (narrow_phase.cuh:30-46) */
    const real ms = ftz((real)ms_in), co_tol = ftz((real)co_tol_in);
    rquery rq;
    to_rquery(q_in, &rq);
    const rquery* q = &rq;
    orc_bounds bd;
    compute_bounds(q, is_vf, co_tol, ms > 0, &bd);

    static __thread orc_box stack[ORC_STACK_MAX];
    int sp = 0;
    for (int k = 0; k < 3; k++) {
        stack[0].lo[k] = 0;
        stack[0].hi[k] = 1;
    }
    sp = 1;
    int64_t checks = 0;
    const real one_plus = (real)1 / ((real)1 - REAL_EPSILON); /* root_finder.cu:24 */
    int capped = 0;

    while (sp > 0) {
        if (sp > st->max_stack)
            st->max_stack = sp;
        const orc_box box = stack[--sp];
        const real min_t = box.lo[0];
        const int64_t seen = checks++; /* root_finder.cu:288-289 */
        if (min_t >= *prune_toi) /* root_finder.cu:295-300 */
            continue;
        if (max_iter >= 0 && seen > max_iter) { /* root_finder.cu:303-305 */
            capped = 1;
            if (cap_mode == 1) {
                if (min_t < *prune_toi)
                    *prune_toi = min_t;
                if (min_t < *global_toi)
                    *global_toi = min_t;
            }
            continue;
        }
        st->box_checks++;
        real true_tol;
        int box_in;
        if (!origin_in_inclusion(q, &bd, is_vf, ms, &box, &true_tol, &box_in))
            continue;
        const real w[3] = { ftz(box.hi[0] - box.lo[0]), ftz(box.hi[1] - box.lo[1]),
                            ftz(box.hi[2] - box.lo[2]) };
        int accept = 0;
        /* Condition 1, root_finder.cu:322 */
        if (w[0] <= bd.tol[0] && w[1] <= bd.tol[1] && w[2] <= bd.tol[2])
            accept = 1;
        /* Condition 2, root_finder.cu:331 */
        else if (box_in && (allow_zero_toi || min_t > 0))
            accept = 1;
        /* Condition 3, root_finder.cu:340-341 */
        else if (true_tol <= co_tol && (allow_zero_toi || min_t > 0))
            accept = 1;
        if (!accept) {
            const int split = split_dimension(&bd, w);
            /* interval.cuh:18-27 */
            const real mid = ftz(ftz(box.lo[split] + box.hi[split]) / 2);
            /* Condition 4, root_finder.cu:222-225,362 */
            if (box.lo[split] >= mid || mid >= box.hi[split]) {
                accept = 1;
            } else {
                orc_box first = box, second = box;
                first.hi[split] = mid;
                second.lo[split] = mid;
                int push_second;
                if (split == 0) /* root_finder.cu:229-232 */
                    push_second = mid <= *prune_toi;
                else if (is_vf) /* root_finder.cu:234-247 */
                    push_second = ftz(mid + box.lo[split == 1 ? 2 : 1]) <= one_plus;
                else /* root_finder.cu:249 */
                    push_second = 1;
                if (sp + 2 > ORC_STACK_MAX)
                    abort();
                if (push_second)
                    stack[sp++] = second;
                stack[sp++] = first; /* popped first: earliest / lower half */
            }
        }
        if (accept) {
            if (min_t < *prune_toi)
                *prune_toi = min_t;
            if (min_t < *global_toi)
                *global_toi = min_t;
        }
    }
    if (capped)
        st->capped_queries++;
    return checks;
}

/*
 * cuda/narrow_phase/narrow_phase.cu:108-206 + root_finder.cu:372-457 on direct
 * query arrays.
 *   toi_per_query == NULL : reference default build -- one shared running toi
 *                           (in/out, ccd.cu:125 starts it at 1.0).
 *   toi_per_query != NULL : SCALABLE_CCD_TOI_PER_QUERY build -- per-query toi
 *                           initialised to INFINITY (narrow_phase.cu:70), hit
 *                           <=> toi < 1 (narrow_phase.cu:77-82); *toi still
 *                           receives the global minimum.
 */
void orc_narrow_phase(
    const orc_query* queries, int64_t n, int is_vf, double ms, int max_iter,
    double tol, int allow_zero_toi, int cap_mode, double* toi, double* toi_per_query,
    orc_np_stats* stats, int64_t* checks_per_query)
{
    orc_np_stats total = { 0, 0, 0 };
    *toi = (double)(real)*toi; /* Scalar& toi */
    if (toi_per_query) {
        double g = *toi;
#pragma omp parallel
        {
            orc_np_stats st = { 0, 0, 0 };
            double lg = g;
#pragma omp for schedule(dynamic, 256)
            for (int64_t i = 0; i < n; i++) {
                double tq = INFINITY;
                const int64_t nc = solve_query(
                    &queries[i], is_vf, ms, max_iter, tol, allow_zero_toi, cap_mode,
                    &tq, &lg, &st);
                toi_per_query[i] = tq;
                if (checks_per_query)
                    checks_per_query[i] = nc;
            }
#pragma omp critical
            {
                if (lg < g)
                    g = lg;
                total.box_checks += st.box_checks;
                total.capped_queries += st.capped_queries;
                if (st.max_stack > total.max_stack)
                    total.max_stack = st.max_stack;
            }
        }
        *toi = g;
    } else {
        /* shared running toi: sequential so the pruning bound is well defined;
         * the final minimum is order-independent for max_iter < 0. */
        double g = *toi;
        for (int64_t i = 0; i < n && g > 0; i++) { /* narrow_phase.cu:136 */
            double dummy = g;
            const int64_t nc = solve_query(
                &queries[i], is_vf, ms, max_iter, tol, allow_zero_toi, cap_mode, &g,
                &dummy, &total);
            if (checks_per_query)
                checks_per_query[i] = nc;
        }
        *toi = g;
    }
    if (stats)
        *stats = total;
}

/* cuda/narrow_phase/narrow_phase.cu:24-74 (add_data): gather the 8 vertices of
 * each overlap into the query array.  V0/V1 column-major nV x 3; E nE x 2,
 * F nF x 3 column-major int32; pairs = (vertex, face) or (edge a, edge b). */
void orc_gather_queries(
    const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, const orc_pair* pairs, int64_t n, int is_vf,
    orc_query* out)
{
    for (int64_t i = 0; i < n; i++) {
        int32_t v[4];
        if (is_vf) {
            v[0] = pairs[i].a;
            v[1] = F[pairs[i].b];
            v[2] = F[pairs[i].b + nF];
            v[3] = F[pairs[i].b + 2 * nF];
        } else {
            v[0] = E[pairs[i].a];
            v[1] = E[pairs[i].a + nE];
            v[2] = E[pairs[i].b];
            v[3] = E[pairs[i].b + nE];
        }
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 3; k++) {
                out[i].s[j][k] = V0[v[j] + k * nV];
                out[i].e[j][k] = V1[v[j] + k * nV];
            }
    }
}

int orc_sizeof_aabb(void) { return (int)sizeof(orc_aabb); }
int orc_sizeof_query(void) { return (int)sizeof(orc_query); }
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* Level-synchronous BFS over ALL queries at once, exactly the traversal of the
 * reference's host loop (root_finder.cu:431-447: one ccd_kernel launch per level over
 * the ring buffer).  Same accepted minimum as the DFS above; what differs is the WORK:
 * BFS cannot prune by toi until some box reaches an accepting depth, so whole fronts are
 * expanded.  Used to size the reference's queue when freezing goldens and to report the
 * reference's box-check count next to ours.  Returns the number of levels; *max_level
 * receives the widest level (= ring-buffer entries the reference needs). */
typedef struct {
    orc_box box;
    int32_t query;
} orc_bfs_item;

int64_t orc_narrow_phase_bfs(
    const orc_query* queries_in, int64_t n, int is_vf, double ms_in, double tol_in,
    int allow_zero_toi, double* toi_per_query, int64_t* total_checks, int64_t* max_level,
    int64_t cap_items)
{
    const real ms = ftz((real)ms_in), tol = ftz((real)tol_in);
    orc_bounds* bd = (orc_bounds*)malloc(sizeof(orc_bounds) * (size_t)(n > 0 ? n : 1));
    rquery* queries = (rquery*)malloc(sizeof(rquery) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) {
        to_rquery(&queries_in[i], &queries[i]);
        compute_bounds(&queries[i], is_vf, tol, ms > 0, &bd[i]);
        toi_per_query[i] = INFINITY;
    }
    int64_t cap = n > 1024 ? 2 * n : 2048, cur_n = n, levels = 0;
    orc_bfs_item* cur = (orc_bfs_item*)malloc(sizeof(orc_bfs_item) * (size_t)cap);
    orc_bfs_item* nxt = (orc_bfs_item*)malloc(sizeof(orc_bfs_item) * (size_t)cap);
    int64_t nxt_cap = cap;
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) {
            cur[i].box.lo[k] = 0;
            cur[i].box.hi[k] = 1;
        }
        cur[i].query = (int32_t)i;
    }
    const real one_plus = (real)1 / ((real)1 - REAL_EPSILON);
    *total_checks = 0;
    *max_level = n;
    while (cur_n > 0) {
        levels++;
        int64_t nn = 0;
        for (int64_t b = 0; b < cur_n; b++) {
            const orc_box box = cur[b].box;
            const int32_t q = cur[b].query;
            const real min_t = box.lo[0];
            if (min_t >= toi_per_query[q])
                continue;
            (*total_checks)++;
            real true_tol;
            int box_in;
            if (!origin_in_inclusion(&queries[q], &bd[q], is_vf, ms, &box, &true_tol, &box_in))
                continue;
            const real w[3] = { ftz(box.hi[0] - box.lo[0]), ftz(box.hi[1] - box.lo[1]),
                                ftz(box.hi[2] - box.lo[2]) };
            int accept = 0;
            if (w[0] <= bd[q].tol[0] && w[1] <= bd[q].tol[1] && w[2] <= bd[q].tol[2])
                accept = 1;
            else if (box_in && (allow_zero_toi || min_t > 0))
                accept = 1;
            else if (true_tol <= tol && (allow_zero_toi || min_t > 0))
                accept = 1;
            if (!accept) {
                const int split = split_dimension(&bd[q], w);
                const real mid = ftz(ftz(box.lo[split] + box.hi[split]) / 2);
                if (box.lo[split] >= mid || mid >= box.hi[split]) {
                    accept = 1;
                } else {
                    if (nn + 2 > nxt_cap) {
                        nxt_cap *= 2;
                        if (cap_items > 0 && nxt_cap > 4 * cap_items) {
                            free(bd);
                            free(queries);
                            free(cur);
                            free(nxt);
                            *max_level = nxt_cap;
                            return -1; /* gave up: the front exceeds the caller's bound */
                        }
                        nxt = (orc_bfs_item*)realloc(nxt, sizeof(orc_bfs_item) * (size_t)nxt_cap);
                    }
                    nxt[nn].box = box;
                    nxt[nn].box.hi[split] = mid;
                    nxt[nn].query = q;
                    nn++;
                    int push_second;
                    if (split == 0)
                        push_second = mid <= toi_per_query[q];
                    else if (is_vf)
                        push_second = ftz(mid + box.lo[split == 1 ? 2 : 1]) <= one_plus;
                    else
                        push_second = 1;
                    if (push_second) {
                        nxt[nn].box = box;
                        nxt[nn].box.lo[split] = mid;
                        nxt[nn].query = q;
                        nn++;
                    }
                }
            }
            if (accept && min_t < toi_per_query[q])
                toi_per_query[q] = min_t;
        }
        if (nn > *max_level)
            *max_level = nn;
        if (cap < nxt_cap) {
            cap = nxt_cap;
            cur = (orc_bfs_item*)realloc(cur, sizeof(orc_bfs_item) * (size_t)cap);
        }
        orc_bfs_item* t = cur;
        cur = nxt;
        nxt = t;
        const int64_t tc = cap;
        cap = nxt_cap;
        nxt_cap = tc;
        cur_n = nn;
    }
    free(bd);
    free(queries);
    free(cur);
    free(nxt);
    return levels;
}
