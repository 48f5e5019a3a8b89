#pragma once
#include <memory_resource>
namespace vecmem {
using memory_resource = std::pmr::memory_resource;
}
