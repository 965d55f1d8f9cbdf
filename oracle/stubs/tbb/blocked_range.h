// Stand-in for oneTBB (not installable offline).  TEST INFRASTRUCTURE ONLY --
// used to compile the unmodified reference sources into oracle/_ref/.
// tbb::parallel_for -> OpenMP, tbb::parallel_sort -> __gnu_parallel::sort.
#pragma once
#include <cstddef>
namespace tbb {
template <typename T> class blocked_range {
public:
    blocked_range(T b, T e, std::size_t grain = 1) : m_b(b), m_e(e), m_g(grain) { }
    T begin() const { return m_b; }
    T end() const { return m_e; }
    std::size_t size() const { return std::size_t(m_e - m_b); }
    std::size_t grainsize() const { return m_g; }
    bool empty() const { return !(m_b < m_e); }
private:
    T m_b, m_e;
    std::size_t m_g;
};
} // namespace tbb
