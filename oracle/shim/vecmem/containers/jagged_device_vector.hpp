#pragma once
#include "vecmem/containers/data/jagged_vector_view.hpp"
#include "vecmem/containers/device_vector.hpp"
namespace vecmem {
template <typename T>
class jagged_device_vector {
    public:
    using size_type = unsigned int;
    using value_type = device_vector<T>;
    using reference = device_vector<T>;
    using const_reference = device_vector<T>;
    jagged_device_vector(const data::jagged_vector_view<T>& v) : m_size(v.m_size), m_ptr(v.m_ptr) {}
    size_type size() const { return m_size; }
    bool empty() const { return m_size == 0; }
    device_vector<T> at(size_type i) const { return device_vector<T>(m_ptr[i]); }
    device_vector<T> operator[](size_type i) const { return device_vector<T>(m_ptr[i]); }
    private:
    size_type m_size;
    data::vector_view<T>* m_ptr;
};
}
