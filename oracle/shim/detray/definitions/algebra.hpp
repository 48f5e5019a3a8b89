#pragma once
// Stand-in for detray's "array" algebra plugin (algebra-plugins cmath/array): std::array
// points/vectors with the published definitions of the vector helpers.
#include <array>
#include <cmath>
#include <cstddef>
#include <ostream>

#include "detray/definitions/indexing.hpp"

#ifndef DETRAY_CUSTOM_SCALARTYPE
#define DETRAY_CUSTOM_SCALARTYPE float
#endif
#ifndef DETRAY_HOST_DEVICE
#if defined(__CUDACC__)  // the reference's CUDA sources (oracle/ref_cuda_seeding.cu)
#define DETRAY_HOST_DEVICE __host__ __device__
#define DETRAY_HOST __host__
#define DETRAY_DEVICE __device__
#else
#define DETRAY_HOST_DEVICE
#define DETRAY_HOST
#define DETRAY_DEVICE
#endif
#endif

namespace detray {
template <typename T>
struct array {
    using scalar = T;
};
template <typename A>
using dscalar = typename A::scalar;
template <typename A>
using dpoint2D = std::array<typename A::scalar, 2>;
template <typename A>
using dpoint3D = std::array<typename A::scalar, 3>;
template <typename A>
using dvector3D = std::array<typename A::scalar, 3>;
template <typename A>
struct dtransform3D {};

namespace algebra {
namespace array {
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline std::array<T, N> operator+(const std::array<T, N>& a, const std::array<T, N>& b) {
    std::array<T, N> r;
    for (std::size_t i = 0; i < N; ++i) r[i] = a[i] + b[i];
    return r;
}
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline std::array<T, N> operator-(const std::array<T, N>& a, const std::array<T, N>& b) {
    std::array<T, N> r;
    for (std::size_t i = 0; i < N; ++i) r[i] = a[i] - b[i];
    return r;
}
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline std::array<T, N> operator*(T s, const std::array<T, N>& a) {
    std::array<T, N> r;
    for (std::size_t i = 0; i < N; ++i) r[i] = s * a[i];
    return r;
}
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline std::array<T, N> operator*(const std::array<T, N>& a, T s) {
    return s * a;
}
}  // namespace array
template <typename T, std::size_t N>
inline std::ostream& operator<<(std::ostream& os, const std::array<T, N>& a) {
    for (std::size_t i = 0; i < N; ++i) os << (i ? ", " : "[") << a[i];
    return os << "]";
}
}  // namespace algebra

namespace vector {
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline T dot(const std::array<T, N>& a, const std::array<T, N>& b) {
    T r = a[0] * b[0];
    for (std::size_t i = 1; i < N; ++i) r += a[i] * b[i];
    return r;
}
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline T norm(const std::array<T, N>& a) {
    return std::sqrt(dot(a, a));
}
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline T perp(const std::array<T, N>& a) {
    return std::sqrt(a[0] * a[0] + a[1] * a[1]);
}
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline T phi(const std::array<T, N>& a) {
    return std::atan2(a[1], a[0]);
}
template <typename T>
DETRAY_HOST_DEVICE inline T theta(const std::array<T, 3>& a) {
    return std::atan2(perp(a), a[2]);
}
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline std::array<T, N> normalize(const std::array<T, N>& a) {
    return algebra::array::operator*(static_cast<T>(1) / norm(a), a);
}
template <typename T>
DETRAY_HOST_DEVICE inline std::array<T, 3> cross(const std::array<T, 3>& a, const std::array<T, 3>& b) {
    return {a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]};
}
}  // namespace vector
namespace getter {}
namespace matrix {}
}  // namespace detray
