#pragma once
#include "vecmem/containers/data/vector_view.hpp"
namespace vecmem::data {
template <typename T>
struct jagged_vector_view {
    using size_type = unsigned int;
    size_type m_size = 0;
    vector_view<T>* m_ptr = nullptr;
    jagged_vector_view() = default;
    template <typename O>
    jagged_vector_view(const O& o) : m_size(o.m_size), m_ptr(reinterpret_cast<vector_view<T>*>(o.m_ptr)) {}
};
}
#include "vecmem/containers/data/jagged_vector_data.hpp"
