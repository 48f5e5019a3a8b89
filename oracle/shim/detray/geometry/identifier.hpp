#pragma once
// Stand-in for detray::geometry::identifier (detray/geometry/identifier.hpp; earlier releases:
// geometry::barcode): a 64-bit word with an encoded (volume, id, index, ...) value. The seeding
// path only copies it (measurement::surface_link / identifier columns), so the stand-in keeps
// the raw value.
#include <cstdint>
#include <ostream>

#include "detray/definitions/algebra.hpp"

namespace detray::geometry {
class identifier {
    public:
    using value_t = std::uint64_t;
    constexpr identifier() = default;
    DETRAY_HOST_DEVICE constexpr explicit identifier(value_t v) : m_value(v) {}
    DETRAY_HOST_DEVICE constexpr value_t value() const { return m_value; }
    DETRAY_HOST_DEVICE constexpr bool is_invalid() const { return m_value == ~static_cast<value_t>(0); }
    DETRAY_HOST_DEVICE friend constexpr bool operator==(const identifier& a, const identifier& b) {
        return a.m_value == b.m_value;
    }
    DETRAY_HOST_DEVICE friend constexpr auto operator<=>(const identifier& a, const identifier& b) {
        return a.m_value <=> b.m_value;
    }
    friend std::ostream& operator<<(std::ostream& os, const identifier& c) { return os << c.m_value; }

    private:
    value_t m_value = ~static_cast<value_t>(0);
};
using barcode = identifier;
}  // namespace detray::geometry
