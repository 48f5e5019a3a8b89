#pragma once
#include "vecmem/containers/data/vector_view.hpp"
#include "vecmem/containers/vector.hpp"
namespace vecmem {
template <typename T>
class device_vector {
    public:
    using size_type = unsigned int;
    using value_type = T;
    using reference = T&;
    using const_reference = const T&;
    device_vector(const data::vector_view<T>& v) : m_size(v.size()), m_ptr(v.ptr()) {}
    size_type size() const { return m_size; }
    bool empty() const { return m_size == 0; }
    T& at(size_type i) const { return m_ptr[i]; }
    T& operator[](size_type i) const { return m_ptr[i]; }
    T* begin() const { return m_ptr; }
    T* end() const { return m_ptr + m_size; }
    private:
    size_type m_size;
    T* m_ptr;
};
template <typename T, typename A>
data::vector_view<T> get_data(std::vector<T, A>& v) { return {static_cast<unsigned int>(v.size()), v.data()}; }
template <typename T, typename A>
data::vector_view<const T> get_data(const std::vector<T, A>& v) { return {static_cast<unsigned int>(v.size()), v.data()}; }
}
