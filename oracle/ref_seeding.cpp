// ref_seeding.cpp — the REFERENCE's own host seeding algorithm
// (traccc::host::seeding_algorithm = spacepoint_binning + seed_finding: doublet_finding,
// triplet_finding, seed_filtering), compiled verbatim from the sources under /root/reference
// (never copied) against the stand-in third-party headers in oracle/shim (vecmem containers,
// detray algebra / units, Acts logger). Exported through a tiny C API so that the tests can
// compare the oracle's restatement — and the CUDA path — with the reference's own loops on
// whole events, and so that bench.py can time the reference's code on the host cores.
// TEST INFRASTRUCTURE: built into oracle/_ref/libtraccc_ref_seeding.so by `make -C oracle ref`.
#include <cstdint>
#include <cstring>
#include <memory_resource>

#include <vecmem/memory/host_memory_resource.hpp>

// reference sources, verbatim
#include "traccc/seeding/seeding_algorithm.hpp"
#include "spacepoint_binning.cpp"
#include "seed_filtering.cpp"
#include "seed_finding.cpp"
#include "seeding_algorithm.cpp"

#include "../include/b200seed.h"

// traccc::getDummyLogger lives in core/src/utils/logging.cpp, which builds real Acts log
// writers; with the stand-in logger it is just a static object.
namespace traccc {
const Logger& getDummyLogger() {
    static const Logger l;
    return l;
}
}  // namespace traccc

namespace {
template <typename R, typename C>
R cfg_cast(const C* c) {
    static_assert(sizeof(R) == sizeof(C));
    R r;
    std::memcpy(static_cast<void*>(&r), c, sizeof(r));
    return r;
}
}  // namespace

extern "C" {

// Runs traccc::host::seeding_algorithm on n spacepoints. Seeds are written to the caller's
// arrays (capacity cap); returns the number of seeds found, or -1 if get_axes threw.
long ref_seeding_run(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                     const b200seed_filter_cfg* filter, uint32_t n, const float* xyz,
                     const float* var_z, const float* var_r, uint32_t cap, uint32_t* bottom,
                     uint32_t* middle, uint32_t* top, float* quality) {
    try {
        vecmem::host_memory_resource mr;
        const auto f = cfg_cast<traccc::seedfinder_config>(finder);
        traccc::spacepoint_grid_config g(f);  // no default constructor: overwrite a copy
        static_assert(sizeof(g) == sizeof(*grid));
        std::memcpy(static_cast<void*>(&g), grid, sizeof(g));
        const auto fl = cfg_cast<traccc::seedfilter_config>(filter);
        traccc::host::seeding_algorithm alg(f, g, fl, mr);
        traccc::edm::spacepoint_collection::host sps{mr};
        sps.resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            auto sp = sps.at(i);
            sp.measurement_index_1() = i;
            sp.measurement_index_2() = 0xFFFFFFFFu;
            sp.global() = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
            sp.z_variance() = var_z ? var_z[i] : 0.f;
            sp.radius_variance() = var_r ? var_r[i] : 0.f;
        }
        const traccc::edm::spacepoint_collection::const_view view = vecmem::get_data(sps);
        const auto seeds = alg(view);
        const std::size_t ns = seeds.size();
        for (std::size_t i = 0; i < ns && i < cap; ++i) {
            bottom[i] = seeds.bottom_index()[i];
            middle[i] = seeds.middle_index()[i];
            top[i] = seeds.top_index()[i];
            quality[i] = seeds.quality()[i];
        }
        return static_cast<long>(ns);
    } catch (const std::domain_error&) {
        return -1;
    }
}

}  // extern "C"
