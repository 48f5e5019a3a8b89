#pragma once
#include <tuple>
namespace detray {
template <typename... T>
using tuple = std::tuple<T...>;
}
