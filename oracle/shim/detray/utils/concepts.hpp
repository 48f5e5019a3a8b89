#pragma once
#include <concepts>
#include <type_traits>
namespace detray::concepts {
template <typename T>
concept scalar = std::is_arithmetic_v<std::remove_cvref_t<T>>;
template <typename T, typename U>
concept same_as_no_const = std::same_as<std::remove_cv_t<T>, std::remove_cv_t<U>>;
template <typename T>
concept algebra = true;
}  // namespace detray::concepts
namespace detray::concepts {
template <typename T>
concept point2D = requires(const T& p) { p[0]; p[1]; };
template <typename T>
concept point3D = requires(const T& p) { p[0]; p[1]; p[2]; };
}  // namespace detray::concepts
