#pragma once
#include "vecmem/containers/vector.hpp"
#include "vecmem/containers/data/jagged_vector_data.hpp"
namespace vecmem {
template <typename T, typename A1, typename A2>
data::jagged_vector_data<T> get_data(std::vector<std::vector<T, A1>, A2>& v, memory_resource* = nullptr) {
    data::jagged_vector_data<T> d(static_cast<unsigned int>(v.size()));
    for (std::size_t i = 0; i < v.size(); ++i) d.m_ptr[i] = data::vector_view<T>(static_cast<unsigned int>(v[i].size()), v[i].data());
    return d;
}
template <typename T, typename A1, typename A2>
data::jagged_vector_data<const T> get_data(const std::vector<std::vector<T, A1>, A2>& v, memory_resource* = nullptr) {
    data::jagged_vector_data<const T> d(static_cast<unsigned int>(v.size()));
    for (std::size_t i = 0; i < v.size(); ++i) d.m_ptr[i] = data::vector_view<const T>(static_cast<unsigned int>(v[i].size()), v[i].data());
    return d;
}
}
