#pragma once
#include <omp.h>
namespace tbb { namespace info {
inline int default_concurrency() { return omp_get_max_threads(); }
} }
