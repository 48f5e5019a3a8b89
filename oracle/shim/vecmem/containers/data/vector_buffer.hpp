#pragma once
#include "vecmem/containers/data/vector_view.hpp"
namespace vecmem::data {
template <typename T>
struct vector_buffer : vector_view<T> {};
}
