#pragma once
// Stand-in for detray/tracks/tracks.hpp: bound / free track parameters of the array plugin with
// the accessors the seeding path uses. Layout of bound_track_parameters: 64-bit surface
// identifier, 6 x 1 parameter vector, 6 x 6 covariance (zero-initialised), like the published type.
#include <array>
#include <cstddef>

#include "detray/definitions/algebra.hpp"
#include "detray/definitions/track_parametrization.hpp"
#include "detray/geometry/identifier.hpp"
#include "detray/utils/concepts.hpp"

namespace detray {
template <typename A>
using bound_matrix = dmatrix<A, e_bound_size, e_bound_size>;

template <typename A>
struct bound_parameters_vector {
    using scalar_type = typename A::scalar;
    using vector_type = dmatrix<A, e_bound_size, 1>;
    vector_type m_vector{};
    DETRAY_HOST_DEVICE const vector_type& vector() const { return m_vector; }
    DETRAY_HOST_DEVICE vector_type& vector() { return m_vector; }
    DETRAY_HOST_DEVICE scalar_type operator[](std::size_t i) const { return m_vector[0][i]; }
    DETRAY_HOST_DEVICE std::array<scalar_type, 2> bound_local() const { return {m_vector[0][0], m_vector[0][1]}; }
    DETRAY_HOST_DEVICE void set_bound_local(const std::array<scalar_type, 2>& p) {
        m_vector[0][e_bound_loc0] = p[0];
        m_vector[0][e_bound_loc1] = p[1];
    }
    DETRAY_HOST_DEVICE scalar_type phi() const { return m_vector[0][e_bound_phi]; }
    DETRAY_HOST_DEVICE void set_phi(scalar_type v) { m_vector[0][e_bound_phi] = v; }
    DETRAY_HOST_DEVICE scalar_type theta() const { return m_vector[0][e_bound_theta]; }
    DETRAY_HOST_DEVICE void set_theta(scalar_type v) { m_vector[0][e_bound_theta] = v; }
    DETRAY_HOST_DEVICE scalar_type qop() const { return m_vector[0][e_bound_qoverp]; }
    DETRAY_HOST_DEVICE void set_qop(scalar_type v) { m_vector[0][e_bound_qoverp] = v; }
    DETRAY_HOST_DEVICE scalar_type time() const { return m_vector[0][e_bound_time]; }
    DETRAY_HOST_DEVICE void set_time(scalar_type v) { m_vector[0][e_bound_time] = v; }
};

template <typename A>
struct bound_track_parameters : public bound_parameters_vector<A> {
    using covariance_type = bound_matrix<A>;
    geometry::identifier m_barcode{};
    covariance_type m_covariance{};
    DETRAY_HOST_DEVICE const geometry::identifier& surface_link() const { return m_barcode; }
    DETRAY_HOST_DEVICE void set_surface_link(geometry::identifier b) { m_barcode = b; }
    DETRAY_HOST_DEVICE const covariance_type& covariance() const { return m_covariance; }
    DETRAY_HOST_DEVICE covariance_type& covariance() { return m_covariance; }
    DETRAY_HOST_DEVICE void set_covariance(const covariance_type& c) { m_covariance = c; }
};

template <typename A>
struct free_track_parameters {
    using vector_type = dmatrix<A, e_free_size, 1>;
    vector_type m_vector{};
};
}  // namespace detray
