#pragma once
#include "vecmem/memory/memory_resource.hpp"
