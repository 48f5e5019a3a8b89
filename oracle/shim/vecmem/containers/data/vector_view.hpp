#pragma once
#include <cstddef>
#include <type_traits>
namespace vecmem::data {
template <typename T>
struct vector_view {
    using size_type = unsigned int;
    size_type m_capacity = 0;
    size_type* m_size = nullptr;
    T* m_ptr = nullptr;
    vector_view() = default;
    vector_view(size_type n, T* p) : m_capacity(n), m_ptr(p) {}
    template <typename U, std::enable_if_t<std::is_same_v<std::remove_cv_t<T>, std::remove_cv_t<U>>, bool> = true>
    vector_view(const vector_view<U>& o) : m_capacity(o.m_capacity), m_size(o.m_size), m_ptr(o.m_ptr) {}
    size_type size() const { return m_size ? *m_size : m_capacity; }
    size_type capacity() const { return m_capacity; }
    T* ptr() const { return m_ptr; }
};
}
