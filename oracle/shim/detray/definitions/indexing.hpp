#pragma once
#include "vecmem/containers/vector.hpp"
namespace detray {
using dindex = unsigned int;
template <typename T>
using dvector = vecmem::vector<T>;
using dindex_sequence = dvector<dindex>;
}
