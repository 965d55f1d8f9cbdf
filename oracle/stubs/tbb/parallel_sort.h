#pragma once
#include <parallel/algorithm>
namespace tbb {
template <typename It, typename Cmp> void parallel_sort(It b, It e, const Cmp& cmp)
{
    __gnu_parallel::sort(b, e, cmp);
}
template <typename It> void parallel_sort(It b, It e) { __gnu_parallel::sort(b, e); }
} // namespace tbb
