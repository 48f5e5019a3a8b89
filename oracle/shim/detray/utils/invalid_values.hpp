#pragma once
#include <limits>
namespace detray::detail {
template <typename T>
constexpr T invalid_value() { return std::numeric_limits<T>::max(); }
template <typename T>
constexpr bool is_invalid_value(const T& v) { return v == invalid_value<T>(); }
}
