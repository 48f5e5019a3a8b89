#pragma once
// detray::unit / detray::constant (detray/definitions/units.hpp), published values.
namespace detray {
template <typename scalar_t>
struct unit {
    static constexpr scalar_t um{static_cast<scalar_t>(1e-3)};
    static constexpr scalar_t mm{static_cast<scalar_t>(1.0)};
    static constexpr scalar_t cm{static_cast<scalar_t>(10.0)};
    static constexpr scalar_t m{static_cast<scalar_t>(1000.0)};
    static constexpr scalar_t s{static_cast<scalar_t>(299792458000.0)};
    static constexpr scalar_t ns{static_cast<scalar_t>(1e-9 * 299792458000.0)};
    static constexpr scalar_t rad{static_cast<scalar_t>(1.0)};
    static constexpr scalar_t degree{static_cast<scalar_t>(0.017453292519943295)};
    static constexpr scalar_t eV{static_cast<scalar_t>(1e-9)};
    static constexpr scalar_t keV{static_cast<scalar_t>(1e-6)};
    static constexpr scalar_t MeV{static_cast<scalar_t>(1e-3)};
    static constexpr scalar_t GeV{static_cast<scalar_t>(1.0)};
    static constexpr scalar_t TeV{static_cast<scalar_t>(1e3)};
    static constexpr scalar_t e{static_cast<scalar_t>(1.0)};
    static constexpr scalar_t T{static_cast<scalar_t>(0.000299792458)};
};
template <typename scalar_t>
struct constant {
    static constexpr scalar_t pi{static_cast<scalar_t>(3.14159265358979323846)};
    static constexpr scalar_t pi_2{static_cast<scalar_t>(1.57079632679489661923)};
    static constexpr scalar_t pi_4{static_cast<scalar_t>(0.785398163397448309616)};
};
}  // namespace detray
