// Reads like the reference's tests/test_narrow_phase.cu:41-65 and tests/test_broad_phase.cu:88-104,
// written against include/sccd.hpp.  Built and run by tests/test_cpp_shim.py.
#include <sccd.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
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

    std::vector<AABB> vb, eb, fb;
    build_boxes(V0, V1, E, F, vb, eb, fb);
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

    std::printf("toi=%.17g toi_pq=%.17g ipc=%.17g collisions=%zu vf=%zu ee=%zu threw=%d\n", toi,
                toi2, toi3, collisions.size(), vf.size(), ee.size(), (int)threw);
    const bool ok = toi <= 0.5 && 0.5 - toi < 1e-5 && toi2 == toi && toi3 == toi
        && collisions.size() == 1 && std::get<0>(collisions[0]) == 3 && vf.size() == 1
        && vf[0].first == 3 && vf[0].second == 0 && ee.empty() && threw;
    return ok ? 0 : 1;
}
