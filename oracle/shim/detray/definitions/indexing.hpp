#pragma once
#include <vector>
namespace detray {
using dindex = unsigned int;
using dindex_sequence = std::vector<dindex>;
}
