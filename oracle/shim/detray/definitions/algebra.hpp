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
    using value_type = T;
};
template <typename A>
using dscalar = typename A::scalar;
template <typename A>
using dpoint2D = std::array<typename A::scalar, 2>;
template <typename A>
using dpoint3D = std::array<typename A::scalar, 3>;
template <typename A>
using dvector3D = std::array<typename A::scalar, 3>;
// Column-major matrices of the cmath/array plugin: m[col][row]; getter::element(m, row, col).
template <typename A, std::size_t ROWS, std::size_t COLS>
using dmatrix = std::array<std::array<typename A::scalar, ROWS>, COLS>;

// algebra-plugins cmath `transform3` (restated from its published implementation — the library
// itself is not available here): a 4x4 matrix whose columns are the x, y, z axes and the
// translation, plus its inverse obtained by cofactor expansion at construction;
// point_to_local(p) = rotate(inverse, p) + (translation column of the inverse).
template <typename A>
struct dtransform3D {
    using scalar_type = typename A::scalar;
    using point3 = std::array<scalar_type, 3>;
    using vector3 = std::array<scalar_type, 3>;
    using matrix44 = std::array<std::array<scalar_type, 4>, 4>;
    matrix44 _data{};
    matrix44 _data_inv{};

    DETRAY_HOST_DEVICE dtransform3D() {
        for (std::size_t i = 0; i < 4; ++i) _data[i][i] = _data_inv[i][i] = scalar_type(1);
    }
    DETRAY_HOST_DEVICE dtransform3D(const vector3& t, const vector3& x, const vector3& y, const vector3& z,
                                    bool get_inverse = true) {
        for (std::size_t r = 0; r < 3; ++r) {
            _data[0][r] = x[r];
            _data[1][r] = y[r];
            _data[2][r] = z[r];
            _data[3][r] = t[r];
        }
        _data[0][3] = _data[1][3] = _data[2][3] = scalar_type(0);
        _data[3][3] = scalar_type(1);
        if (get_inverse) _data_inv = invert(_data);
    }
    // determinant from the first row of cofactors
    DETRAY_HOST_DEVICE static scalar_type determinant(const matrix44& m) {
        return m[0][3] * m[1][2] * m[2][1] * m[3][0] - m[0][2] * m[1][3] * m[2][1] * m[3][0] -
               m[0][3] * m[1][1] * m[2][2] * m[3][0] + m[0][1] * m[1][3] * m[2][2] * m[3][0] +
               m[0][2] * m[1][1] * m[2][3] * m[3][0] - m[0][1] * m[1][2] * m[2][3] * m[3][0] -
               m[0][3] * m[1][2] * m[2][0] * m[3][1] + m[0][2] * m[1][3] * m[2][0] * m[3][1] +
               m[0][3] * m[1][0] * m[2][2] * m[3][1] - m[0][0] * m[1][3] * m[2][2] * m[3][1] -
               m[0][2] * m[1][0] * m[2][3] * m[3][1] + m[0][0] * m[1][2] * m[2][3] * m[3][1] +
               m[0][3] * m[1][1] * m[2][0] * m[3][2] - m[0][1] * m[1][3] * m[2][0] * m[3][2] -
               m[0][3] * m[1][0] * m[2][1] * m[3][2] + m[0][0] * m[1][3] * m[2][1] * m[3][2] +
               m[0][1] * m[1][0] * m[2][3] * m[3][2] - m[0][0] * m[1][1] * m[2][3] * m[3][2] -
               m[0][2] * m[1][1] * m[2][0] * m[3][3] + m[0][1] * m[1][2] * m[2][0] * m[3][3] +
               m[0][2] * m[1][0] * m[2][1] * m[3][3] - m[0][0] * m[1][2] * m[2][1] * m[3][3] -
               m[0][1] * m[1][0] * m[2][2] * m[3][3] + m[0][0] * m[1][1] * m[2][2] * m[3][3];
    }
    DETRAY_HOST_DEVICE static matrix44 invert(const matrix44& m) {
        matrix44 i;
        i[0][0] = m[1][2] * m[2][3] * m[3][1] - m[1][3] * m[2][2] * m[3][1] + m[1][3] * m[2][1] * m[3][2] -
                  m[1][1] * m[2][3] * m[3][2] - m[1][2] * m[2][1] * m[3][3] + m[1][1] * m[2][2] * m[3][3];
        i[0][1] = m[0][3] * m[2][2] * m[3][1] - m[0][2] * m[2][3] * m[3][1] - m[0][3] * m[2][1] * m[3][2] +
                  m[0][1] * m[2][3] * m[3][2] + m[0][2] * m[2][1] * m[3][3] - m[0][1] * m[2][2] * m[3][3];
        i[0][2] = m[0][2] * m[1][3] * m[3][1] - m[0][3] * m[1][2] * m[3][1] + m[0][3] * m[1][1] * m[3][2] -
                  m[0][1] * m[1][3] * m[3][2] - m[0][2] * m[1][1] * m[3][3] + m[0][1] * m[1][2] * m[3][3];
        i[0][3] = m[0][3] * m[1][2] * m[2][1] - m[0][2] * m[1][3] * m[2][1] - m[0][3] * m[1][1] * m[2][2] +
                  m[0][1] * m[1][3] * m[2][2] + m[0][2] * m[1][1] * m[2][3] - m[0][1] * m[1][2] * m[2][3];
        i[1][0] = m[1][3] * m[2][2] * m[3][0] - m[1][2] * m[2][3] * m[3][0] - m[1][3] * m[2][0] * m[3][2] +
                  m[1][0] * m[2][3] * m[3][2] + m[1][2] * m[2][0] * m[3][3] - m[1][0] * m[2][2] * m[3][3];
        i[1][1] = m[0][2] * m[2][3] * m[3][0] - m[0][3] * m[2][2] * m[3][0] + m[0][3] * m[2][0] * m[3][2] -
                  m[0][0] * m[2][3] * m[3][2] - m[0][2] * m[2][0] * m[3][3] + m[0][0] * m[2][2] * m[3][3];
        i[1][2] = m[0][3] * m[1][2] * m[3][0] - m[0][2] * m[1][3] * m[3][0] - m[0][3] * m[1][0] * m[3][2] +
                  m[0][0] * m[1][3] * m[3][2] + m[0][2] * m[1][0] * m[3][3] - m[0][0] * m[1][2] * m[3][3];
        i[1][3] = m[0][2] * m[1][3] * m[2][0] - m[0][3] * m[1][2] * m[2][0] + m[0][3] * m[1][0] * m[2][2] -
                  m[0][0] * m[1][3] * m[2][2] - m[0][2] * m[1][0] * m[2][3] + m[0][0] * m[1][2] * m[2][3];
        i[2][0] = m[1][1] * m[2][3] * m[3][0] - m[1][3] * m[2][1] * m[3][0] + m[1][3] * m[2][0] * m[3][1] -
                  m[1][0] * m[2][3] * m[3][1] - m[1][1] * m[2][0] * m[3][3] + m[1][0] * m[2][1] * m[3][3];
        i[2][1] = m[0][3] * m[2][1] * m[3][0] - m[0][1] * m[2][3] * m[3][0] - m[0][3] * m[2][0] * m[3][1] +
                  m[0][0] * m[2][3] * m[3][1] + m[0][1] * m[2][0] * m[3][3] - m[0][0] * m[2][1] * m[3][3];
        i[2][2] = m[0][1] * m[1][3] * m[3][0] - m[0][3] * m[1][1] * m[3][0] + m[0][3] * m[1][0] * m[3][1] -
                  m[0][0] * m[1][3] * m[3][1] - m[0][1] * m[1][0] * m[3][3] + m[0][0] * m[1][1] * m[3][3];
        i[2][3] = m[0][3] * m[1][1] * m[2][0] - m[0][1] * m[1][3] * m[2][0] - m[0][3] * m[1][0] * m[2][1] +
                  m[0][0] * m[1][3] * m[2][1] + m[0][1] * m[1][0] * m[2][3] - m[0][0] * m[1][1] * m[2][3];
        i[3][0] = m[1][2] * m[2][1] * m[3][0] - m[1][1] * m[2][2] * m[3][0] - m[1][2] * m[2][0] * m[3][1] +
                  m[1][0] * m[2][2] * m[3][1] + m[1][1] * m[2][0] * m[3][2] - m[1][0] * m[2][1] * m[3][2];
        i[3][1] = m[0][1] * m[2][2] * m[3][0] - m[0][2] * m[2][1] * m[3][0] + m[0][2] * m[2][0] * m[3][1] -
                  m[0][0] * m[2][2] * m[3][1] - m[0][1] * m[2][0] * m[3][2] + m[0][0] * m[2][1] * m[3][2];
        i[3][2] = m[0][2] * m[1][1] * m[3][0] - m[0][1] * m[1][2] * m[3][0] - m[0][2] * m[1][0] * m[3][1] +
                  m[0][0] * m[1][2] * m[3][1] + m[0][1] * m[1][0] * m[3][2] - m[0][0] * m[1][1] * m[3][2];
        i[3][3] = m[0][1] * m[1][2] * m[2][0] - m[0][2] * m[1][1] * m[2][0] + m[0][2] * m[1][0] * m[2][1] -
                  m[0][0] * m[1][2] * m[2][1] - m[0][1] * m[1][0] * m[2][2] + m[0][0] * m[1][1] * m[2][2];
        const scalar_type s = scalar_type(1) / determinant(m);
        for (std::size_t c = 0; c < 4; ++c)
            for (std::size_t r = 0; r < 4; ++r) i[c][r] *= s;
        return i;
    }
    DETRAY_HOST_DEVICE static vector3 rotate(const matrix44& m, const vector3& v) {
        return {m[0][0] * v[0] + m[1][0] * v[1] + m[2][0] * v[2], m[0][1] * v[0] + m[1][1] * v[1] + m[2][1] * v[2],
                m[0][2] * v[0] + m[1][2] * v[1] + m[2][2] * v[2]};
    }
    DETRAY_HOST_DEVICE point3 point_to_global(const point3& v) const {
        const vector3 rg = rotate(_data, v);
        return {rg[0] + _data[3][0], rg[1] + _data[3][1], rg[2] + _data[3][2]};
    }
    DETRAY_HOST_DEVICE point3 point_to_local(const point3& v) const {
        const vector3 rg = rotate(_data_inv, v);
        return {rg[0] + _data_inv[3][0], rg[1] + _data_inv[3][1], rg[2] + _data_inv[3][2]};
    }
};

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
namespace getter {
// element i of a point / vector
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline T& element(std::array<T, N>& v, std::size_t i) {
    return v[i];
}
template <typename T, std::size_t N>
DETRAY_HOST_DEVICE inline T element(const std::array<T, N>& v, std::size_t i) {
    return v[i];
}
// element (row, col) of a column-major matrix; element (row, 0) of a column vector stored as one
template <typename T, std::size_t ROWS, std::size_t COLS>
DETRAY_HOST_DEVICE inline T& element(std::array<std::array<T, ROWS>, COLS>& m, std::size_t row, std::size_t col) {
    return m[col][row];
}
template <typename T, std::size_t ROWS, std::size_t COLS>
DETRAY_HOST_DEVICE inline T element(const std::array<std::array<T, ROWS>, COLS>& m, std::size_t row,
                                    std::size_t col) {
    return m[col][row];
}
}  // namespace getter
namespace matrix {
template <typename M>
DETRAY_HOST_DEVICE inline M zero() {
    return M{};
}
}  // namespace matrix
}  // namespace detray
