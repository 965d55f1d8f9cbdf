// extern "C" driver around the UNMODIFIED reference CPU broad phase
// (src/scalable_ccd/broad_phase/{aabb,sort_and_sweep}.cpp), built into
// oracle/_ref/libref_sccd_cpu.so by oracle/Makefile.  TEST INFRASTRUCTURE ONLY.
#include <scalable_ccd/broad_phase/aabb.hpp>
#include <scalable_ccd/broad_phase/sort_and_sweep.hpp>

#include <omp.h>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

using namespace scalable_ccd;

namespace {
struct FlatBox {
    double min[3], max[3];
    int32_t vids[3];
    int32_t elem;
};
void flatten(const std::vector<AABB>& in, FlatBox* out)
{
    for (size_t i = 0; i < in.size(); i++) {
        for (int k = 0; k < 3; k++) {
            out[i].min[k] = in[i].min[k];
            out[i].max[k] = in[i].max[k];
            out[i].vids[k] = int32_t(in[i].vertex_ids[k]);
        }
        out[i].elem = int32_t(in[i].element_id);
    }
}
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
} // namespace

extern "C" {

int ref_cpu_num_threads() { return omp_get_max_threads(); }

// aabb.hpp:79-112
void ref_cpu_build_boxes(
    const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, double r, void* vb, void* eb, void* fb)
{
    Eigen::MatrixXd v0, v1;
    Eigen::MatrixXi e, f;
    to_eigen(V0, V1, nV, E, nE, F, nF, v0, v1, e, f);
    std::vector<AABB> vbox, ebox, fbox;
    build_vertex_boxes(v0, v1, vbox, r);
    build_edge_boxes(vbox, e, ebox);
    build_face_boxes(vbox, f, fbox);
    flatten(vbox, (FlatBox*)vb);
    flatten(ebox, (FlatBox*)eb);
    flatten(fbox, (FlatBox*)fb);
}

// sort_and_sweep.hpp:24-42: VF (two lists) then EE (single list), as
// tests/test_broad_phase.cpp:44-55 drives it.  Returns wall seconds of the
// timed region (boxes + both sweeps).  counts = {nVF, nEE}; axes = in: sort
// axis used for both, out: {next axis after VF, next axis after EE}.
double ref_cpu_broad_phase(
    const double* V0, const double* V1, int64_t nV, const int32_t* E, int64_t nE,
    const int32_t* F, int64_t nF, double r, int sort_axis, int32_t* vf_out,
    int64_t vf_cap, int32_t* ee_out, int64_t ee_cap, int64_t* counts, int* axes_out)
{
    Eigen::MatrixXd v0, v1;
    Eigen::MatrixXi e, f;
    to_eigen(V0, V1, nV, E, nE, F, nF, v0, v1, e, f);

    const auto t0 = std::chrono::steady_clock::now();
    std::vector<AABB> vbox, ebox, fbox;
    build_vertex_boxes(v0, v1, vbox, r);
    build_edge_boxes(vbox, e, ebox);
    build_face_boxes(vbox, f, fbox);

    int axis = sort_axis;
    std::vector<std::pair<int, int>> vf;
    sort_and_sweep(vbox, fbox, axis, vf);
    axes_out[0] = axis;

    axis = sort_axis;
    std::vector<std::pair<int, int>> ee;
    sort_and_sweep(ebox, axis, ee);
    axes_out[1] = axis;
    const auto t1 = std::chrono::steady_clock::now();

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
    return std::chrono::duration<double>(t1 - t0).count();
}
}
