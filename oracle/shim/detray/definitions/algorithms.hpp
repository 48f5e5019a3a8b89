#pragma once
#include <algorithm>
namespace detray {
template <typename It>
void sequential_sort(It b, It e) { std::sort(b, e); }
template <typename It, typename T>
It find(It b, It e, const T& v) { return std::find(b, e, v); }
}
