// ref_tpe.cpp — the REFERENCE's own seed -> bound-track-parameter code, compiled verbatim from the
// sources under /root/reference (never copied):
//   core/include/traccc/seeding/track_params_estimation_helper.hpp   (seed_to_bound_param_vector)
//   core/src/seeding/track_params_estimation.cpp                    (host algorithm + covariance)
//   device/common/include/traccc/seeding/device/impl/estimate_track_params.ipp
//                                                                   (the device function, built for
//                                                                    the host; field sampled at the
//                                                                    bottom spacepoint)
// against the stand-in third-party headers in oracle/shim (vecmem containers; detray
// bound_track_parameters / geometry::identifier; the array plugin's vector helpers and its
// transform3 with the cofactor-expansion inverse). The traccc logic — frame construction, conformal
// fit, which accessor feeds which parameter, the covariance loop — is 100 % the reference's code;
// the arithmetic INSIDE the third-party helpers is a restatement (the libraries are absent).
// TEST INFRASTRUCTURE: built into oracle/_ref/libtraccc_ref_tpe.so by `make -C oracle ref`.
#include <cstdint>
#include <cstring>
#include <memory_resource>

#include <vecmem/memory/host_memory_resource.hpp>

// reference sources, verbatim
#include "traccc/seeding/track_params_estimation.hpp"
#include "track_params_estimation.cpp"
#include "traccc/seeding/device/impl/estimate_track_params.ipp"

#include "../include/b200seed.h"

namespace traccc {
const Logger& getDummyLogger() {
    static const Logger l;
    return l;
}
}  // namespace traccc

namespace {
// what covfie's constant field view returns everywhere
struct const_field {
    float b[3];
    std::array<float, 3> at(float, float, float) const { return {b[0], b[1], b[2]}; }
};

struct inputs {
    vecmem::host_memory_resource mr;
    traccc::edm::measurement_collection::host meas{mr};
    traccc::edm::spacepoint_collection::host sps{mr};
    traccc::edm::seed_collection::host seeds{mr};
    inputs(uint32_t n_sp, const float* xyz, const uint32_t* sp_meas, uint32_t n_meas,
           const float* meas_local, const uint64_t* meas_surface, uint32_t n_seeds, const uint32_t* b,
           const uint32_t* m, const uint32_t* t) {
        meas.resize(n_meas);
        for (uint32_t i = 0; i < n_meas; ++i) {
            auto x = meas.at(i);
            x.local_position() = {meas_local[2 * i], meas_local[2 * i + 1]};
            x.surface_link() = detray::geometry::identifier{meas_surface[i]};
        }
        sps.resize(n_sp);
        for (uint32_t i = 0; i < n_sp; ++i) {
            auto sp = sps.at(i);
            sp.measurement_index_1() = sp_meas ? sp_meas[i] : i;
            sp.measurement_index_2() = 0xFFFFFFFFu;
            sp.global() = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
            sp.z_variance() = 0.f;
            sp.radius_variance() = 0.f;
        }
        seeds.resize(n_seeds);
        for (uint32_t i = 0; i < n_seeds; ++i) {
            auto s = seeds.at(i);
            s.bottom_index() = b[i];
            s.middle_index() = m[i];
            s.top_index() = t[i];
        }
    }
};

void store(const traccc::bound_track_parameters<>& p, b200seed_bound_params& o) {
    o.surface_link = p.surface_link().value();
    for (unsigned k = 0; k < 6; ++k) o.vec[k] = p[k];
    for (unsigned r = 0; r < 6; ++r)
        for (unsigned c = 0; c < 6; ++c) o.cov[r * 6 + c] = traccc::getter::element(p.covariance(), r, c);
}

traccc::track_params_estimation_config cfg_of(const b200seed_tpe_cfg* c) {
    traccc::track_params_estimation_config r;
    static_assert(sizeof(r) == sizeof(*c));
    std::memcpy(static_cast<void*>(&r), c, sizeof(r));
    return r;
}
}  // namespace

extern "C" {

// traccc::host::track_params_estimation on n_seeds seeds; device_variant != 0: the device function
// traccc::device::estimate_track_params (field sampled at the bottom spacepoint), run on the host.
int ref_tpe_run(const b200seed_tpe_cfg* cfg, uint32_t n_sp, const float* xyz, const uint32_t* sp_meas,
                uint32_t n_meas, const float* meas_local, const uint64_t* meas_surface, uint32_t n_seeds,
                const uint32_t* bottom, const uint32_t* middle, const uint32_t* top, const float bfield[3],
                int device_variant, b200seed_bound_params* out) {
    inputs in(n_sp, xyz, sp_meas, n_meas, meas_local, meas_surface, n_seeds, bottom, middle, top);
    const traccc::edm::measurement_collection::const_view mv = vecmem::get_data(in.meas);
    const traccc::edm::spacepoint_collection::const_view sv = vecmem::get_data(in.sps);
    const traccc::edm::seed_collection::const_view dv = vecmem::get_data(in.seeds);
    const traccc::track_params_estimation_config c = cfg_of(cfg);
    if (!device_variant) {
        traccc::host::track_params_estimation alg(c, in.mr);
        const auto res = alg(mv, sv, dv, traccc::vector3{bfield[0], bfield[1], bfield[2]});
        for (uint32_t i = 0; i < n_seeds; ++i) store(res.at(i), out[i]);
        return int(res.size());
    }
    traccc::bound_track_parameters_collection_types::host res(n_seeds, &in.mr);
    traccc::bound_track_parameters_collection_types::view pv = vecmem::get_data(res);
    const const_field f{{bfield[0], bfield[1], bfield[2]}};
    for (uint32_t i = 0; i < n_seeds; ++i) traccc::device::estimate_track_params(i, c, mv, sv, dv, f, pv);
    for (uint32_t i = 0; i < n_seeds; ++i) store(res.at(i), out[i]);
    return int(n_seeds);
}

}  // extern "C"
