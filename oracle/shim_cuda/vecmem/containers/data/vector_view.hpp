#pragma once
// forwards to the single-header CUDA-capable vecmem stand-in (see shim_all.hpp)
#include "vecmem/shim_all.hpp"
