// extern "C" driver around the UNMODIFIED reference CUDA path
// (src/scalable_ccd/cuda/**), built into oracle/_ref/libref_sccd_cuda[_pq].so
// by oracle/Makefile (the _pq variant defines SCALABLE_CCD_TOI_PER_QUERY).
// TEST INFRASTRUCTURE ONLY: used to freeze golden fixtures on a B200 and as the
// "reference CUDA" timing arm.  Never linked into the product.
#include <scalable_ccd/config.hpp>
#include <scalable_ccd/cuda/ccd.cuh>
#include <scalable_ccd/cuda/ipc_ccd_strategy.hpp>
#include <scalable_ccd/cuda/broad_phase/broad_phase.cuh>
#include <scalable_ccd/cuda/broad_phase/aabb.cuh>
#include <scalable_ccd/cuda/narrow_phase/ccd_data.cuh>
#include <scalable_ccd/cuda/narrow_phase/root_finder.cuh>
#include <scalable_ccd/cuda/memory_handler.hpp>
#include <scalable_ccd/utils/logger.hpp>

#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>

using namespace scalable_ccd;
using namespace scalable_ccd::cuda;

namespace {
void to_eigen(
    const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, Eigen::MatrixXd& v0, Eigen::MatrixXd& v1,
    Eigen::MatrixXi& e, Eigen::MatrixXi& f)
{
    v0.resize(nV, 3);
    v1.resize(nV, 3);
    e.resize(nE, 2);
    f.resize(nF, 3);
    std::memcpy(v0.data(), V0, sizeof(double) * nV * 3);
    std::memcpy(v1.data(), V1, sizeof(double) * nV * 3);
    std::memcpy(e.data(), E, sizeof(int32_t) * nE * 2);
    std::memcpy(f.data(), F, sizeof(int32_t) * nF * 3);
}
double now_ms()
{
    return std::chrono::duration<double, std::milli>(
               std::chrono::steady_clock::now().time_since_epoch())
        .count();
}
} // namespace

extern "C" {

int ref_cuda_toi_per_query()
{
#ifdef SCALABLE_CCD_TOI_PER_QUERY
    return 1;
#else
    return 0;
#endif
}

void ref_cuda_quiet() { logger().set_level(spdlog::level::warn); }

// cuda/ccd.cuh:26-38.  In the TOI_PER_QUERY build the collisions
// (aid, bid, toi) are returned through coll_ids / coll_toi (up to coll_cap).
double ref_cuda_ccd(
    const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, double ms, int max_iter, double tol,
    int allow_zero_toi, int32_t* coll_ids, double* coll_toi, int64_t coll_cap,
    int64_t* n_coll, double* elapsed_ms)
{
    Eigen::MatrixXd v0, v1;
    Eigen::MatrixXi e, f;
    to_eigen(V0, V1, nV, E, nE, F, nF, v0, v1, e, f);
    cudaDeviceSynchronize();
    const double t0 = now_ms();
#ifdef SCALABLE_CCD_TOI_PER_QUERY
    std::vector<std::tuple<int, int, Scalar>> collisions;
    const Scalar toi =
        ccd(v0, v1, e, f, ms, max_iter, tol, allow_zero_toi != 0, collisions, 0);
#else
    const Scalar toi = ccd(v0, v1, e, f, ms, max_iter, tol, allow_zero_toi != 0, 0);
#endif
    cudaDeviceSynchronize();
    if (elapsed_ms)
        *elapsed_ms = now_ms() - t0;
#ifdef SCALABLE_CCD_TOI_PER_QUERY
    if (n_coll)
        *n_coll = int64_t(collisions.size());
    for (size_t i = 0; i < collisions.size() && int64_t(i) < coll_cap; i++) {
        coll_ids[2 * i] = std::get<0>(collisions[i]);
        coll_ids[2 * i + 1] = std::get<1>(collisions[i]);
        coll_toi[i] = std::get<2>(collisions[i]);
    }
#else
    if (n_coll)
        *n_coll = -1;
#endif
    return toi;
}

// cuda/ipc_ccd_strategy.hpp:17-24
double ref_cuda_ipc_ccd_strategy(
    const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, double min_distance, int max_iter, double tol,
    double* elapsed_ms)
{
    Eigen::MatrixXd v0, v1;
    Eigen::MatrixXi e, f;
    to_eigen(V0, V1, nV, E, nE, F, nF, v0, v1, e, f);
    cudaDeviceSynchronize();
    const double t0 = now_ms();
    const Scalar toi = ipc_ccd_strategy(v0, v1, e, f, min_distance, max_iter, tol);
    cudaDeviceSynchronize();
    if (elapsed_ms)
        *elapsed_ms = now_ms() - t0;
    return toi;
}

// Broad phase exactly as tests/test_broad_phase.cu:88-104 drives it.
void ref_cuda_broad_phase(
    const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, double r, int32_t* vf_out, int64_t vf_cap,
    int32_t* ee_out, int64_t ee_cap, int64_t* counts, double* elapsed_ms)
{
    Eigen::MatrixXd v0, v1;
    Eigen::MatrixXi e, f;
    to_eigen(V0, V1, nV, E, nE, F, nF, v0, v1, e, f);
    cudaDeviceSynchronize();
    const double t0 = now_ms();
    std::vector<AABB> vb, eb, fb;
    build_vertex_boxes(v0, v1, vb, r);
    build_edge_boxes(vb, e, eb);
    build_face_boxes(vb, f, fb);

    BroadPhase broad_phase;
    broad_phase.build(
        std::make_shared<DeviceAABBs>(vb), std::make_shared<DeviceAABBs>(fb));
    std::vector<std::pair<int, int>> vf = broad_phase.detect_overlaps();
    broad_phase.build(std::make_shared<DeviceAABBs>(eb));
    std::vector<std::pair<int, int>> ee = broad_phase.detect_overlaps();
    cudaDeviceSynchronize();
    if (elapsed_ms)
        *elapsed_ms = now_ms() - t0;

    counts[0] = int64_t(vf.size());
    counts[1] = int64_t(ee.size());
    for (size_t i = 0; i < vf.size() && int64_t(i) < vf_cap; i++) {
        vf_out[2 * i] = vf[i].first;
        vf_out[2 * i + 1] = vf[i].second;
    }
    for (size_t i = 0; i < ee.size() && int64_t(i) < ee_cap; i++) {
        ee_out[2 * i] = ee[i].first;
        ee_out[2 * i + 1] = ee[i].second;
    }
}

// Root finder on direct query arrays: the body of narrow_phase<is_vf>
// (cuda/narrow_phase/narrow_phase.cu:108-206) with add_data replaced by a host
// fill of CCDData, then root_finder.cuh:41-50 ccd<is_vf>().
// queries: n x 24 doubles (v0s v1s v2s v3s v0e v1e v2e v3e).
// Returns the number of overflow reruns, or -1 on error.
int ref_cuda_narrow_queries(
    const double* queries, int64_t n, int is_vf, double ms, int max_iter,
    double tol, int allow_zero_toi, double* toi_inout, double* toi_per_query,
    double* elapsed_ms, int64_t min_queue_units)
{
    thrust::host_vector<CCDData> h(n);
    for (int64_t i = 0; i < n; i++) {
        const double* q = queries + 24 * i;
        CCDData& d = h[i];
        for (int k = 0; k < 3; k++) {
            d.v0s[k] = q[0 + k];
            d.v1s[k] = q[3 + k];
            d.v2s[k] = q[6 + k];
            d.v3s[k] = q[9 + k];
            d.v0e[k] = q[12 + k];
            d.v1e[k] = q[15 + k];
            d.v2e[k] = q[18 + k];
            d.v3e[k] = q[21 + k];
        }
        d.ms = ms;
#ifdef SCALABLE_CCD_TOI_PER_QUERY
        d.toi = INFINITY;
        d.aid = int(i);
        d.bid = int(i);
#endif
        d.nbr_checks = 0;
    }
    auto mh = std::make_shared<MemoryHandler>();
    mh->MAX_QUERIES = size_t(n);
    size_t nq = size_t(n);
    Scalar toi = *toi_inout;
    int reruns = 0;
    bool overflowed = false;
    thrust::device_vector<CCDData> d_data;
    cudaDeviceSynchronize();
    const double t0 = now_ms();
    do {
        if (!overflowed)
            mh->handleNarrowPhase(nq);
        else
            mh->handleOverflow(nq);
        // The reference sizes its ring queue at 2x the query count
        // (memory_handler.cpp:116) and its full-check is not atomic with the push
        // (ccd_buffer.cuh:25-34): when almost every query is deep, as in these
        // adversarial sets, whole BFS levels wrap over unprocessed entries WITHOUT raising
        // the overflow flag and hits are silently lost.  Give it the queue a realistic
        // (mostly shallow) batch would have had.
        if ((int64_t)mh->MAX_UNIT_SIZE < min_queue_units)
            mh->MAX_UNIT_SIZE = (size_t)min_queue_units;
        d_data = h;
        if (is_vf)
            overflowed = ccd<true>(
                d_data, mh, 64, max_iter, tol, ms > 0, allow_zero_toi != 0, toi);
        else
            overflowed = ccd<false>(
                d_data, mh, 64, max_iter, tol, ms > 0, allow_zero_toi != 0, toi);
        cudaDeviceSynchronize();
        if (overflowed)
            reruns++;
        if (reruns > 8)
            return -1;
    } while (overflowed);
    if (elapsed_ms)
        *elapsed_ms = now_ms() - t0;
    *toi_inout = toi;
#ifdef SCALABLE_CCD_TOI_PER_QUERY
    if (toi_per_query) {
        h = d_data;
        for (int64_t i = 0; i < n; i++)
            toi_per_query[i] = h[i].toi;
    }
#else
    (void)toi_per_query;
#endif
    return reruns;
}
}
