#pragma once
#include "vecmem/containers/vector.hpp"
namespace vecmem {
template <typename T>
using jagged_vector = vector<vector<T>>;
}
