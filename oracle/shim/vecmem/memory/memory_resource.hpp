#pragma once
#include <memory_resource>
#include <vector>
namespace vecmem {
using memory_resource = std::pmr::memory_resource;
// (the real headers reach the container aliases through other includes)
template <typename T>
using vector = std::vector<T, std::pmr::polymorphic_allocator<T>>;
template <typename T>
using jagged_vector = vector<vector<T>>;
}
