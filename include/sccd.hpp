// sccd.hpp -- header-only C++17 shim that gives the C ABI of sccd.h the names, argument
// meaning and error behaviour of the reference's C++ API (paths relative to
// /root/reference/src/scalable_ccd):
//
//   scalable_ccd::cuda::ccd(V0, V1, E, F, ms, max_iter, tol, allow_zero_toi[, collisions][, GB])
//                                                                  cuda/ccd.cuh:26-38
//   scalable_ccd::cuda::ipc_ccd_strategy(V0, V1, E, F, min_distance, max_iter, tol)
//                                                                  cuda/ipc_ccd_strategy.hpp:17-24
//   scalable_ccd::cuda::build_vertex_boxes (x2) / build_edge_boxes / build_face_boxes
//                                                                  cuda/broad_phase/aabb.cuh:150-188
//   scalable_ccd::cuda::narrow_phase<is_vf>                        narrow_phase.cuh:30-46
//   scalable_ccd::cuda::DeviceAABBs, BroadPhase                    aabb.cuh:122-148, broad_phase.cuh:15-92
//   scalable_ccd::sort_and_sweep (both overloads)                  broad_phase/sort_and_sweep.hpp:24-42
//
// Matrix arguments are templates over anything with Eigen's dense interface (rows(), cols(),
// data(), column-major storage), so this header compiles without Eigen and accepts
// Eigen::MatrixXd / Eigen::MatrixXi unchanged where Eigen is installed.  Errors surface as
// std::runtime_error, like the reference's gpuErrchk (cuda/utils/assert.cuh:12-28).
// Link with -lsccd_b200.  There is no CPU fallback.
//
// Scalar: double, as in the reference's default build.  Define SCCD_SHIM_USE_FLOAT before
// including this header for the reference's float build (SCALABLE_CCD_USE_DOUBLE off,
// scalar.hpp:16-18): Scalar becomes float and every context is switched to SCCD_F32.
#pragma once

#include "sccd.h"

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace scalable_ccd {

#ifdef SCCD_SHIM_USE_FLOAT
using Scalar = float; // scalar.hpp:16-18
#else
using Scalar = double; // SCALABLE_CCD_USE_DOUBLE build (scalar.hpp:13-15)
#endif

namespace cuda {

    using AABB = ::sccd_aabb; // same 64-byte layout as cuda::AABB

    namespace detail {
        struct Ctx {
            sccd_ctx* h = nullptr;
            explicit Ctx(int device = 0, void* stream = nullptr)
            {
                if (sccd_create(device, stream, &h) != SCCD_OK)
                    throw std::runtime_error("sccd_create failed: no usable sm_100 CUDA device");
#ifdef SCCD_SHIM_USE_FLOAT
                sccd_set_scalar_type(h, SCCD_F32);
#endif
            }
            ~Ctx() { sccd_destroy(h); }
            Ctx(const Ctx&) = delete;
            Ctx& operator=(const Ctx&) = delete;
            void check(int rc) const
            {
                if (rc < 0)
                    throw std::runtime_error(sccd_last_error(h));
            }
        };
        // one lazily created context per thread (the reference is single-threaded, device 0:
        // cuda/broad_phase/broad_phase.cuh:82)
        inline Ctx& default_ctx()
        {
            static thread_local Ctx ctx(0, nullptr);
            return ctx;
        }
        template <typename VMat, typename IMat>
        void upload(Ctx& c, const VMat& V0, const VMat& V1, const IMat& E, const IMat& F)
        {
            if (V0.rows() != V1.rows() || V0.cols() != 3 || V1.cols() != 3 || E.cols() != 2
                || F.cols() != 3) // asserted in ccd.cu:94-98
                throw std::runtime_error("ccd: expected V (n x 3), E (m x 2), F (k x 3)");
            c.check(sccd_upload_mesh(
                c.h, V0.data(), V1.data(), (int64_t)V0.rows(), E.data(), (int64_t)E.rows(),
                F.data(), (int64_t)F.rows(), /*on_device=*/0));
        }
    } // namespace detail

    // max_iter >= 0: 0 (default) = boxes of a query that reaches the cap are accepted at their
    // t_lo (never later than the exact answer); 1 = dropped, the reference's rule
    // (root_finder.cu:303-305), which can miss a collision.  No effect for max_iter < 0.
    inline void set_max_iter_mode(int mode)
    {
        auto& c = detail::default_ctx();
        c.check(sccd_set_option(c.h, SCCD_OPT_MAX_ITER_MODE, mode));
    }

    // ---- cuda/ccd.cuh:26-38 -----------------------------------------------------------
    template <typename VMat, typename IMat>
    Scalar
    ccd(const VMat& vertices_t0, const VMat& vertices_t1, const IMat& edges, const IMat& faces,
        const Scalar minimum_separation_distance, const int max_iterations,
        const Scalar tolerance, const bool allow_zero_toi, const int memory_limit_GB = 0)
    {
        auto& c = detail::default_ctx();
        c.check(sccd_set_memory_limit(c.h, (size_t)memory_limit_GB << 30));
        detail::upload(c, vertices_t0, vertices_t1, edges, faces);
        double toi = 1;
        c.check(sccd_ccd(
            c.h, minimum_separation_distance, max_iterations, tolerance, allow_zero_toi, &toi));
        return (Scalar)toi;
    }

    // SCALABLE_CCD_TOI_PER_QUERY overload (ccd.cuh:35-37)
    template <typename VMat, typename IMat>
    Scalar
    ccd(const VMat& vertices_t0, const VMat& vertices_t1, const IMat& edges, const IMat& faces,
        const Scalar minimum_separation_distance, const int max_iterations,
        const Scalar tolerance, const bool allow_zero_toi,
        std::vector<std::tuple<int, int, Scalar>>& collisions, const int memory_limit_GB = 0)
    {
        auto& c = detail::default_ctx();
        c.check(sccd_set_memory_limit(c.h, (size_t)memory_limit_GB << 30));
        detail::upload(c, vertices_t0, vertices_t1, edges, faces);
        double toi = 1;
        int64_t nvf = 0, nee = 0;
        c.check(sccd_ccd_collisions(
            c.h, minimum_separation_distance, max_iterations, tolerance, allow_zero_toi, &toi,
            nullptr, nullptr, 0, &nvf, &nee));
        std::vector<sccd_pair> ids((size_t)(nvf + nee));
        std::vector<double> tois(ids.size());
        c.check(sccd_get_collisions(c.h, ids.data(), tois.data(), (int64_t)ids.size(), &nvf, &nee));
        for (size_t i = 0; i < ids.size(); i++)
            collisions.emplace_back(ids[i].a, ids[i].b, (Scalar)tois[i]);
        return (Scalar)toi;
    }

    // ---- cuda/ipc_ccd_strategy.hpp:17-24 ------------------------------------------------
    template <typename VMat, typename IMat>
    Scalar ipc_ccd_strategy(
        const VMat& V0, const VMat& V1, const IMat& E, const IMat& F, const Scalar min_distance,
        const int max_iterations, const Scalar tolerance)
    {
        auto& c = detail::default_ctx();
        detail::upload(c, V0, V1, E, F);
        double toi = 1;
        c.check(sccd_ipc_ccd_strategy(c.h, min_distance, max_iterations, tolerance, &toi));
        return (Scalar)toi;
    }

    // ---- cuda/narrow_phase/narrow_phase.cuh:30-46 ------------------------------------------
    // The reference passes its DeviceMatrix / thrust containers; here the mesh is the one the
    // (per-thread) context holds -- upload_mesh() below -- and the overlaps are a device array
    // of pairs, e.g. what BroadPhase::detect_overlaps_partial() returned.  `toi` is lowered in
    // place, exactly as the reference's Scalar& toi.
    template <typename VMat, typename IMat>
    void upload_mesh(const VMat& V0, const VMat& V1, const IMat& E, const IMat& F,
                     double inflation_radius = 0)
    {
        auto& c = detail::default_ctx();
        detail::upload(c, V0, V1, E, F);
        c.check(sccd_build_boxes(c.h, inflation_radius));
    }
    template <bool is_vf>
    void narrow_phase(
        const sccd_pair* d_overlaps, const int64_t n_overlaps, const int max_iter, const Scalar tol,
        const Scalar minimum_separation_distance, const bool allow_zero_toi, Scalar& toi)
    {
        auto& c = detail::default_ctx();
        double t = toi;
        c.check(sccd_narrow_phase(
            c.h, is_vf ? SCCD_VF : SCCD_EE, d_overlaps, n_overlaps, minimum_separation_distance,
            max_iter, tol, allow_zero_toi, &t, nullptr));
        toi = (Scalar)t;
    }

    // ---- cuda/broad_phase/aabb.cuh:150-188 ------------------------------------------------
    // The reference builds vertex boxes on the host and derives edge / face boxes from them;
    // here all three come from one device pass, so the mesh-level call is the primitive and
    // the three reference-named functions are views of it.
    template <typename VMat, typename IMat>
    void build_boxes(
        const VMat& V0, const VMat& V1, const IMat& E, const IMat& F,
        std::vector<AABB>& vertex_boxes, std::vector<AABB>& edge_boxes,
        std::vector<AABB>& face_boxes, const double inflation_radius = 0)
    {
        auto& c = detail::default_ctx();
        detail::upload(c, V0, V1, E, F);
        c.check(sccd_build_boxes(c.h, inflation_radius));
        vertex_boxes.resize((size_t)V0.rows());
        edge_boxes.resize((size_t)E.rows());
        face_boxes.resize((size_t)F.rows());
        c.check(sccd_get_boxes(c.h, 0, vertex_boxes.data()));
        c.check(sccd_get_boxes(c.h, 1, edge_boxes.data()));
        c.check(sccd_get_boxes(c.h, 2, face_boxes.data()));
    }

    // The reference's own three builders (aabb.cuh:150-188), computed on the device from
    // host arrays, as tests/test_broad_phase.cu:88-91 calls them.
    template <typename VMat>
    void build_vertex_boxes(
        const VMat& vertices_t0, const VMat& vertices_t1, std::vector<AABB>& vertex_boxes,
        double inflation_radius = 0)
    {
        if (vertices_t0.rows() != vertices_t1.rows() || vertices_t0.cols() != 3
            || vertices_t1.cols() != 3) // asserted in aabb.cu:152-153
            throw std::runtime_error("build_vertex_boxes: expected two (n x 3) matrices");
        auto& c = detail::default_ctx();
        vertex_boxes.resize((size_t)vertices_t0.rows());
        c.check(sccd_build_vertex_boxes(
            c.h, vertices_t0.data(), vertices_t1.data(), (int64_t)vertices_t0.rows(),
            inflation_radius, vertex_boxes.data()));
    }
    template <typename VMat>
    void build_vertex_boxes(
        const VMat& vertices, std::vector<AABB>& vertex_boxes, const double inflation_radius = 0)
    {
        if (vertices.cols() != 3) // aabb.cu:121
            throw std::runtime_error("build_vertex_boxes: expected an (n x 3) matrix");
        auto& c = detail::default_ctx();
        vertex_boxes.resize((size_t)vertices.rows());
        c.check(sccd_build_vertex_boxes(
            c.h, vertices.data(), nullptr, (int64_t)vertices.rows(), inflation_radius,
            vertex_boxes.data()));
    }
    template <typename IMat>
    void build_edge_boxes(
        const std::vector<AABB>& vertex_boxes, const IMat& edges, std::vector<AABB>& edge_boxes)
    {
        if (edges.rows() > 0 && edges.cols() != 2)
            throw std::runtime_error("build_edge_boxes: expected an (n x 2) matrix");
        auto& c = detail::default_ctx();
        edge_boxes.resize((size_t)edges.rows());
        c.check(sccd_build_element_boxes(
            c.h, vertex_boxes.data(), (int64_t)vertex_boxes.size(), edges.data(),
            (int64_t)edges.rows(), 2, edge_boxes.data()));
    }
    template <typename IMat>
    void build_face_boxes(
        const std::vector<AABB>& vertex_boxes, const IMat& faces, std::vector<AABB>& face_boxes)
    {
        if (faces.rows() > 0 && faces.cols() != 3)
            throw std::runtime_error("build_face_boxes: expected an (n x 3) matrix");
        auto& c = detail::default_ctx();
        face_boxes.resize((size_t)faces.rows());
        c.check(sccd_build_element_boxes(
            c.h, vertex_boxes.data(), (int64_t)vertex_boxes.size(), faces.data(),
            (int64_t)faces.rows(), 3, face_boxes.data()));
    }

    /// A list of caller-made boxes (cuda/broad_phase/aabb.cuh:122-148).  Sorting happens on
    /// the device when the list is handed to BroadPhase::build.
    struct DeviceAABBs {
        DeviceAABBs() = default;
        explicit DeviceAABBs(const std::vector<AABB>& b) : boxes(b) { }
        size_t size() const { return boxes.size(); }
        void clear() { boxes.clear(); }
        std::vector<AABB> boxes;
    };

    /// cuda/broad_phase/broad_phase.cuh:15-92
    class BroadPhase {
    public:
        BroadPhase() = default;

        void clear()
        {
            m_built = false;
            m_num_boxes = 0;
            m_overlaps = { nullptr, 0 };
        }

        void build(const std::shared_ptr<DeviceAABBs> boxes)
        {
            auto& c = detail::default_ctx();
            c.check(sccd_set_boxes(
                c.h, boxes->boxes.data(), (int64_t)boxes->size(), nullptr, 0, 0, nullptr));
            c.check(sccd_broad_phase_begin(c.h, SCCD_BOXES));
            m_built = true;
            m_num_boxes = boxes->size();
        }

        void build(const std::shared_ptr<DeviceAABBs> boxesA, const std::shared_ptr<DeviceAABBs> boxesB)
        {
            auto& c = detail::default_ctx();
            c.check(sccd_set_boxes(
                c.h, boxesA->boxes.data(), (int64_t)boxesA->size(), boxesB->boxes.data(),
                (int64_t)boxesB->size(), 0, nullptr));
            c.check(sccd_broad_phase_begin(c.h, SCCD_BOXES));
            m_built = true;
            m_num_boxes = boxesA->size() + boxesB->size();
        }

        /// Next chunk of overlaps; device pointer + count, valid until the next call.
        const std::pair<const sccd_pair*, int64_t>& detect_overlaps_partial()
        {
            require_built();
            auto& c = detail::default_ctx();
            c.check(sccd_broad_phase_partial(c.h, &m_overlaps.first, &m_overlaps.second));
            return m_overlaps;
        }

        std::vector<std::pair<int, int>> detect_overlaps()
        {
            require_built();
            auto& c = detail::default_ctx();
            int64_t n = 0;
            c.check(sccd_broad_phase(c.h, SCCD_BOXES, nullptr, 0, &n));
            std::vector<sccd_pair> tmp((size_t)n);
            c.check(sccd_broad_phase(c.h, SCCD_BOXES, tmp.data(), n, &n));
            std::vector<std::pair<int, int>> out((size_t)n);
            for (int64_t i = 0; i < n; i++)
                out[(size_t)i] = { tmp[(size_t)i].a, tmp[(size_t)i].b };
            return out;
        }

        bool is_complete() const
        {
            if (!m_built)
                return true; // broad_phase.cuh:50 with no boxes: 0 >= 0
            return sccd_broad_phase_is_complete(detail::default_ctx().h) == 1;
        }

        size_t num_boxes() const { return m_num_boxes; }
        const std::pair<const sccd_pair*, int64_t>& overlaps() const { return m_overlaps; }

        int threads_per_block = 32; // kept for source compatibility; the tiled sweep ignores it

    private:
        void require_built() const
        {
            if (!m_built) // broad_phase.cu:123-126
                throw std::runtime_error(
                    "Must initialize build broad phase before detecting overlaps!");
        }
        bool m_built = false;
        size_t m_num_boxes = 0;
        std::pair<const sccd_pair*, int64_t> m_overlaps { nullptr, 0 };
    };

} // namespace cuda

// ---- broad_phase/sort_and_sweep.hpp:24-42 (GPU-backed drop-ins for the CPU entry points) ----
using AABB = ::sccd_aabb;

inline void sort_and_sweep(
    const std::vector<AABB>& boxes, int& sort_axis, std::vector<std::pair<int, int>>& overlaps)
{
    overlaps.clear();
    if (boxes.empty()) // sort_and_sweep.cpp:205-207
        return;
    auto& c = cuda::detail::default_ctx();
    int next = 0;
    c.check(sccd_set_boxes(c.h, boxes.data(), (int64_t)boxes.size(), nullptr, 0, sort_axis, &next));
    int64_t n = 0;
    c.check(sccd_broad_phase(c.h, SCCD_BOXES, nullptr, 0, &n));
    std::vector<sccd_pair> tmp((size_t)n);
    c.check(sccd_broad_phase(c.h, SCCD_BOXES, tmp.data(), n, &n));
    overlaps.reserve((size_t)n);
    for (const auto& p : tmp)
        overlaps.emplace_back(p.a, p.b);
    sort_axis = next;
}

inline void sort_and_sweep(
    const std::vector<AABB>& boxesA, const std::vector<AABB>& boxesB, int& sort_axis,
    std::vector<std::pair<int, int>>& overlaps)
{
    overlaps.clear();
    if (boxesA.empty() || boxesB.empty()) // sort_and_sweep.cpp:221-223
        return;
    auto& c = cuda::detail::default_ctx();
    int next = 0;
    c.check(sccd_set_boxes(
        c.h, boxesA.data(), (int64_t)boxesA.size(), boxesB.data(), (int64_t)boxesB.size(),
        sort_axis, &next));
    int64_t n = 0;
    c.check(sccd_broad_phase(c.h, SCCD_BOXES, nullptr, 0, &n));
    std::vector<sccd_pair> tmp((size_t)n);
    c.check(sccd_broad_phase(c.h, SCCD_BOXES, tmp.data(), n, &n));
    overlaps.reserve((size_t)n);
    for (const auto& p : tmp)
        overlaps.emplace_back(p.a, p.b);
    sort_axis = next;
}

} // namespace scalable_ccd
