#pragma once
#include <memory_resource>
#include <vector>
#include "vecmem/memory/memory_resource.hpp"
namespace vecmem {
template <typename T>
using vector = std::vector<T, std::pmr::polymorphic_allocator<T>>;
}
