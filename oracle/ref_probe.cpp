// ref_probe.cpp — the REFERENCE's own seeding helper headers, compiled verbatim from
// /root/reference (never copied) against the stand-in third-party headers in oracle/shim,
// exported through a tiny C API so tests can check the oracle's restatement bit for bit.
// TEST INFRASTRUCTURE: built into oracle/_ref/libtraccc_ref.so by `make -C oracle ref`.
#include <array>
#include <cstdint>
#include <cstring>
#include <memory_resource>

#include "traccc/seeding/detail/seeding_config.hpp"       // reference
#include "traccc/seeding/doublet_finding_helper.hpp"      // reference
#include "traccc/seeding/grids/axis.hpp"                  // reference
#include "traccc/seeding/seed_selecting_helper.hpp"       // reference
#include "traccc/seeding/triplet_finding_helper.hpp"      // reference

#include "../include/b200seed.h"

namespace {

// BASE of the reference's edm::spacepoint<BASE> interface: a one-element "proxy".
struct sp_base {
    unsigned int m1 = 0, m2 = 0;
    std::array<float, 3> g{};
    float vz = 0.f, vr = 0.f;
    sp_base() = default;
    sp_base(const float* p) : g{p[0], p[1], p[2]}, vz(p[3]), vr(p[4]) {}
    template <std::size_t I>
    auto& get() {
        if constexpr (I == 0) return m1;
        else if constexpr (I == 1) return m2;
        else if constexpr (I == 2) return g;
        else if constexpr (I == 3) return vz;
        else return vr;
    }
    template <std::size_t I>
    const auto& get() const {
        if constexpr (I == 0) return m1;
        else if constexpr (I == 1) return m2;
        else if constexpr (I == 2) return g;
        else if constexpr (I == 3) return vz;
        else return vr;
    }
};
using ref_sp = traccc::edm::spacepoint<sp_base>;

// The C-ABI config structs must be byte-for-byte the reference PODs.
static_assert(sizeof(traccc::seedfinder_config) == sizeof(b200seed_finder_cfg));
static_assert(sizeof(traccc::spacepoint_grid_config) == sizeof(b200seed_grid_cfg));
static_assert(sizeof(traccc::seedfilter_config) == sizeof(b200seed_filter_cfg));
static_assert(offsetof(traccc::seedfinder_config, maxSeedsPerSpM) == offsetof(b200seed_finder_cfg, maxSeedsPerSpM));
static_assert(offsetof(traccc::seedfinder_config, highland) == offsetof(b200seed_finder_cfg, highland));
static_assert(offsetof(traccc::seedfinder_config, neighbor_scope) == offsetof(b200seed_finder_cfg, neighbor_scope));
static_assert(offsetof(traccc::seedfilter_config, compatSeedLimit) == offsetof(b200seed_filter_cfg, compatSeedLimit));
static_assert(offsetof(traccc::seedfilter_config, spB_min_radius) == offsetof(b200seed_filter_cfg, spB_min_radius));
static_assert(offsetof(traccc::spacepoint_grid_config, phiBinDeflectionCoverage) == offsetof(b200seed_grid_cfg, phiBinDeflectionCoverage));

traccc::seedfinder_config to_ref(const b200seed_finder_cfg* c) {
    traccc::seedfinder_config r;
    std::memcpy(static_cast<void*>(&r), c, sizeof(r));
    return r;
}
traccc::seedfilter_config to_ref(const b200seed_filter_cfg* c) {
    traccc::seedfilter_config r;
    std::memcpy(static_cast<void*>(&r), c, sizeof(r));
    return r;
}
std::pmr::memory_resource& mr() { return *std::pmr::new_delete_resource(); }

}  // namespace

extern "C" {

void ref_finder_cfg_defaults(b200seed_finder_cfg* out) {
    traccc::seedfinder_config c;  // in-class defaults + setup()
    std::memcpy(out, &c, sizeof(c));
}
void ref_finder_cfg_setup(b200seed_finder_cfg* io) {
    traccc::seedfinder_config c = to_ref(io);
    c.setup();
    std::memcpy(io, &c, sizeof(c));
}
void ref_grid_cfg_from_finder(const b200seed_finder_cfg* f, b200seed_grid_cfg* out) {
    traccc::spacepoint_grid_config g(to_ref(f));
    std::memcpy(out, &g, sizeof(g));
}
void ref_filter_cfg_defaults(b200seed_filter_cfg* out) {
    traccc::seedfilter_config c;
    std::memcpy(out, &c, sizeof(c));
}

int ref_doublet_is_compatible(int bottom, const float m[5], const float o[5],
                              const b200seed_finder_cfg* c) {
    const ref_sp sp1(m), sp2(o);
    const traccc::seedfinder_config cfg = to_ref(c);
    if (bottom)
        return traccc::doublet_finding_helper::isCompatible<traccc::details::spacepoint_type::bottom>(sp1, sp2, cfg);
    return traccc::doublet_finding_helper::isCompatible<traccc::details::spacepoint_type::top>(sp1, sp2, cfg);
}
void ref_transform_coordinates(int bottom, const float m[5], const float o[5], float lc[6]) {
    const ref_sp sp1(m), sp2(o);
    const traccc::lin_circle l =
        bottom ? traccc::doublet_finding_helper::transform_coordinates<traccc::details::spacepoint_type::bottom>(sp1, sp2)
               : traccc::doublet_finding_helper::transform_coordinates<traccc::details::spacepoint_type::top>(sp1, sp2);
    lc[0] = l.Zo(), lc[1] = l.cotTheta(), lc[2] = l.iDeltaR(), lc[3] = l.Er(), lc[4] = l.U(), lc[5] = l.V();
}
// iSinTheta2 / scatteringInRegion2 as core/src/seeding/triplet_finding.hpp:77-82 computes them
int ref_triplet_is_compatible(const float m[5], const float lb[6], const float lt[6],
                              const b200seed_finder_cfg* c, float out[2]) {
    const ref_sp spM(m);
    const traccc::seedfinder_config cfg = to_ref(c);
    const traccc::lin_circle b{lb[0], lb[1], lb[2], lb[3], lb[4], lb[5]};
    const traccc::lin_circle t{lt[0], lt[1], lt[2], lt[3], lt[4], lt[5]};
    const traccc::scalar iSinTheta2 = 1.f + b.cotTheta() * b.cotTheta();
    traccc::scalar scatteringInRegion2 = cfg.maxScatteringAngle2 * iSinTheta2;
    scatteringInRegion2 *= cfg.sigmaScattering * cfg.sigmaScattering;
    traccc::scalar curvature = 0.f, impact = 0.f;
    const bool ok = traccc::triplet_finding_helper::isCompatible(spM, b, t, cfg, iSinTheta2,
                                                                scatteringInRegion2, curvature, impact);
    out[0] = curvature;
    out[1] = impact;
    return ok;
}
// seed_selecting_helper: returns the updated weight; flags[0] = single_seed_cut, flags[1] = cut_per_middle_sp
float ref_seed_select(const b200seed_filter_cfg* c, const float m[5], const float b[5], const float t[5],
                      float weight, int flags[2]) {
    const traccc::seedfilter_config cfg = to_ref(c);
    const ref_sp spM(m), spB(b), spT(t);
    traccc::scalar w = weight;
    traccc::seed_selecting_helper::seed_weight(cfg, spM, spB, spT, w);
    flags[0] = traccc::seed_selecting_helper::single_seed_cut(cfg, spM, spB, spT, w);
    flags[1] = traccc::seed_selecting_helper::cut_per_middle_sp(cfg, spB, w);
    return w;
}
float ref_sp_radius(const float p[5]) { return ref_sp(p).radius(); }
float ref_sp_phi(const float p[5]) { return ref_sp(p).phi(); }

uint32_t ref_axis_regular_bin(uint32_t n, float mn, float mx, float v) {
    return traccc::axis2::regular<>{n, mn, mx, mr()}.bin(v);
}
uint32_t ref_axis_circular_bin(uint32_t n, float mn, float mx, float v) {
    return traccc::axis2::circular<>{n, mn, mx, mr()}.bin(v);
}
uint32_t ref_axis_circular_remap(uint32_t n, float mn, float mx, uint32_t ibin, int shood) {
    return traccc::axis2::circular<>{n, mn, mx, mr()}.remap(ibin, shood);
}
void ref_axis_regular_range(uint32_t n, float mn, float mx, float v, uint32_t n0, uint32_t n1, uint32_t out[2]) {
    const auto r = traccc::axis2::regular<>{n, mn, mx, mr()}.range(v, std::array<unsigned int, 2>{n0, n1});
    out[0] = r[0], out[1] = r[1];
}
void ref_axis_circular_range(uint32_t n, float mn, float mx, float v, uint32_t n0, uint32_t n1, uint32_t out[2]) {
    const auto r = traccc::axis2::circular<>{n, mn, mx, mr()}.range(v, std::array<unsigned int, 2>{n0, n1});
    out[0] = r[0], out[1] = r[1];
}
uint32_t ref_axis_zone(int circular, uint32_t n, float mn, float mx, float v, uint32_t n0, uint32_t n1,
                       uint32_t* out, uint32_t cap) {
    const std::array<unsigned int, 2> nh{n0, n1};
    const auto z = circular ? traccc::axis2::circular<>{n, mn, mx, mr()}.zone(v, nh)
                            : traccc::axis2::regular<>{n, mn, mx, mr()}.zone(v, nh);
    for (uint32_t i = 0; i < z.size() && i < cap; ++i) out[i] = z[i];
    return static_cast<uint32_t>(z.size());
}

}  // extern "C"
