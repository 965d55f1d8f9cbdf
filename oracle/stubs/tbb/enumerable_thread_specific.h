#pragma once
#include <omp.h>
#include <vector>
namespace tbb {
template <typename T> class enumerable_thread_specific {
    struct alignas(64) Slot {
        T value;
    };
public:
    enumerable_thread_specific() : m_slots(size_t(omp_get_max_threads())) { }
    T& local() { return m_slots[size_t(omp_get_thread_num())].value; }
    class const_iterator {
    public:
        const_iterator(const Slot* p) : m_p(p) { }
        const T& operator*() const { return m_p->value; }
        const T* operator->() const { return &m_p->value; }
        const_iterator& operator++()
        {
            ++m_p;
            return *this;
        }
        bool operator!=(const const_iterator& o) const { return m_p != o.m_p; }
        bool operator==(const const_iterator& o) const { return m_p == o.m_p; }
    private:
        const Slot* m_p;
    };
    const_iterator begin() const { return const_iterator(m_slots.data()); }
    const_iterator end() const { return const_iterator(m_slots.data() + m_slots.size()); }
    size_t size() const { return m_slots.size(); }
private:
    std::vector<Slot> m_slots;
};
} // namespace tbb
