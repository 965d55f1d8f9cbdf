#pragma once
#include <tbb/blocked_range.h>
#include <omp.h>
#include <algorithm>
#include <exception>
namespace tbb {
template <typename T, typename F>
void parallel_for(const blocked_range<T>& range, const F& f)
{
    const long long b = (long long)range.begin(), e = (long long)range.end();
    if (e <= b)
        return;
    const long long n = e - b;
    const long long nthreads = omp_get_max_threads();
    // ~16 chunks per thread, dynamically scheduled (TBB's auto partitioner
    // also over-decomposes and steals).
    long long chunk = std::max<long long>(1, n / (nthreads * 16));
    const long long nchunks = (n + chunk - 1) / chunk;
    std::exception_ptr err = nullptr;
#pragma omp parallel for schedule(dynamic, 1)
    for (long long c = 0; c < nchunks; c++) {
        const long long cb = b + c * chunk;
        const long long ce = std::min(e, cb + chunk);
        try {
            f(blocked_range<T>(T(cb), T(ce)));
        } catch (...) {
#pragma omp critical
            err = std::current_exception();
        }
    }
    if (err)
        std::rethrow_exception(err);
}
} // namespace tbb
