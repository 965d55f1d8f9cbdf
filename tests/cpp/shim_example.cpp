// Reads like the reference's tests/test_narrow_phase.cu:41-65 and tests/test_broad_phase.cu:88-104,
// written against include/sccd.hpp.  Built and run by tests/test_cpp_shim.py.
#include <sccd.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

// minimal column-major matrix with Eigen's dense interface
template <typename T> struct Mat {
    long r = 0, c = 0;
    std::vector<T> v;
    Mat(long rows, long cols) : r(rows), c(cols), v((size_t)(rows * cols)) { }
    long rows() const { return r; }
    long cols() const { return c; }
    const T* data() const { return v.data(); }
    T& operator()(long i, long j) { return v[(size_t)(i + j * r)]; }
};

int main(int argc, char** argv)
{
    using namespace scalable_ccd;
    using namespace scalable_ccd::cuda;
    // a vertex falling through a static triangle: earliest contact at t = 0.5
    Mat<double> V0(4, 3), V1(4, 3);
    const double tri[3][3] = { { 0, 0, 0 }, { 1, 0, 0 }, { 0, 1, 0 } };
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++)
            V0(i, k) = V1(i, k) = tri[i][k];
    V0(3, 0) = V1(3, 0) = 0.25;
    V0(3, 1) = V1(3, 1) = 0.25;
    V0(3, 2) = 1.0;
    V1(3, 2) = -1.0;
    Mat<int> E(3, 2), F(1, 3);
    E(0, 0) = 0; E(0, 1) = 1; E(1, 0) = 1; E(1, 1) = 2; E(2, 0) = 0; E(2, 1) = 2;
    F(0, 0) = 0; F(0, 1) = 1; F(0, 2) = 2;

    constexpr bool allow_zero_toi = true;
    constexpr Scalar min_distance = 0;
    constexpr int max_iterations = -1;
    constexpr Scalar tolerance = 1e-6;
    const Scalar toi = ccd(V0, V1, E, F, min_distance, max_iterations, tolerance, allow_zero_toi);
    std::vector<std::tuple<int, int, Scalar>> collisions;
    const Scalar toi2 =
        ccd(V0, V1, E, F, min_distance, max_iterations, tolerance, allow_zero_toi, collisions);
    const Scalar toi3 = ipc_ccd_strategy(V0, V1, E, F, min_distance, max_iterations, tolerance);
    // the reference's max_iter rule (drop) on request; without a cap both rules are the same
    set_max_iter_mode(1);
    const Scalar toi_drop = ccd(V0, V1, E, F, min_distance, max_iterations, tolerance, allow_zero_toi);
    set_max_iter_mode(0);
    if (toi_drop != toi) {
        std::fprintf(stderr, "max_iter mode changed an uncapped result\n");
        return 1;
    }

    std::vector<AABB> vb, eb, fb;
    build_boxes(V0, V1, E, F, vb, eb, fb);
    // the reference's own three builders (tests/test_broad_phase.cu:88-91) give the same boxes
    std::vector<AABB> vb2, eb2, fb2, vb_static;
    build_vertex_boxes(V0, V1, vb2);
    build_edge_boxes(vb2, E, eb2);
    build_face_boxes(vb2, F, fb2);
    build_vertex_boxes(V0, vb_static);
    auto same = [](const std::vector<AABB>& a, const std::vector<AABB>& b) {
        return a.size() == b.size()
            && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(AABB)) == 0);
    };
    const bool builders_ok = same(vb, vb2) && same(eb, eb2) && same(fb, fb2)
        && vb_static.size() == vb.size() && vb_static[3].max[2] >= 1.0 && vb_static[3].min[2] > 0.9;
    BroadPhase broad_phase;
    bool threw = false;
    try {
        broad_phase.detect_overlaps_partial();
    } catch (const std::runtime_error&) {
        threw = true;
    }
    broad_phase.build(std::make_shared<DeviceAABBs>(vb), std::make_shared<DeviceAABBs>(fb));
    const auto vf = broad_phase.detect_overlaps();
    int axis = 0;
    std::vector<std::pair<int, int>> ee;
    sort_and_sweep(eb, axis, ee);
    // step-wise: partial overlaps on the device -> narrow_phase<true> (ccd.cu:55-76)
    upload_mesh(V0, V1, E, F);
    BroadPhase bp2;
    bp2.build(std::make_shared<DeviceAABBs>(vb), std::make_shared<DeviceAABBs>(fb));
    Scalar toi4 = 1;
    while (!bp2.is_complete()) {
        const auto& ov = bp2.detect_overlaps_partial();
        narrow_phase<true>(ov.first, ov.second, max_iterations, tolerance, min_distance,
                           allow_zero_toi, toi4);
    }

    std::printf("toi=%.17g toi_pq=%.17g ipc=%.17g stepwise=%.17g collisions=%zu vf=%zu ee=%zu "
                "threw=%d builders=%d\n", (double)toi, (double)toi2, (double)toi3, (double)toi4,
                collisions.size(), vf.size(), ee.size(), (int)threw, (int)builders_ok);
    const bool ok = toi <= 0.5 && 0.5 - toi < 1e-5 && toi2 == toi && toi3 == toi && toi4 == toi
        && builders_ok
        && collisions.size() == 1 && std::get<0>(collisions[0]) == 3 && vf.size() == 1
        && vf[0].first == 3 && vf[0].second == 0 && ee.empty() && threw;
    return ok ? 0 : 1;
}
