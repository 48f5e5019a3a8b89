#pragma once
#include <memory>
#include <vector>
#include "vecmem/containers/data/vector_view.hpp"
namespace vecmem::data {
// Owning array of per-row views over a host jagged vector (what vecmem::get_data returns).
template <typename T>
struct jagged_vector_data {
    using size_type = unsigned int;
    size_type m_size = 0;
    vector_view<T>* m_ptr = nullptr;
    std::shared_ptr<std::vector<vector_view<T>>> m_rows;
    jagged_vector_data() = default;
    explicit jagged_vector_data(size_type n)
        : m_size(n), m_rows(std::make_shared<std::vector<vector_view<T>>>(n)) {
        m_ptr = m_rows->data();
    }
    template <typename O>
    jagged_vector_data(const O& o) : m_size(o.m_size), m_ptr(reinterpret_cast<vector_view<T>*>(o.m_ptr)) {}
};
}
