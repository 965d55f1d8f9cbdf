#pragma once
#include <omp.h>
#include <cstddef>
namespace tbb {
class global_control {
public:
    enum parameter { max_allowed_parallelism, thread_stack_size };
    global_control(parameter p, std::size_t v)
    {
        if (p == max_allowed_parallelism)
            omp_set_num_threads(int(v));
    }
};
} // namespace tbb
