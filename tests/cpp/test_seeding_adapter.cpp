// The reference's own seeding unit tests (tests/cpu/test_seeding.cpp:35-181 and
// tests/cpu/test_track_params_estimation.cpp:34-144), written against the C++ adapter
// (include/traccc_b200/seeding.hpp) the way the reference writes them against
// traccc::host::seeding_algorithm / track_params_estimation — but running on the GPU.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "traccc_b200/seeding.hpp"

using namespace traccc::b200;

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                      \
        }                                                                  \
    } while (0)

static const float unit_T = 0.000299792458f;

struct device_event {
    cuda_device_memory_resource mr;
    device_allocation xyz, meas_idx, local, surf;
    spacepoint_const_view sp;
    measurement_const_view meas;
    explicit device_event(const std::vector<float>& pts) {
        const std::uint32_t n = static_cast<std::uint32_t>(pts.size() / 3);
        std::vector<std::uint32_t> mi(n);
        std::vector<float> loc(2 * n, 0.f);
        std::vector<std::uint64_t> sf(n);
        for (std::uint32_t i = 0; i < n; ++i) mi[i] = i, sf[i] = 100 + i, loc[2 * i] = 0.5f * i;
        xyz = device_allocation(mr, pts.size() * 4);
        meas_idx = device_allocation(mr, n * 4);
        local = device_allocation(mr, n * 8);
        surf = device_allocation(mr, n * 8);
        cudaMemcpy(xyz.get(), pts.data(), pts.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(meas_idx.get(), mi.data(), n * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(local.get(), loc.data(), n * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(surf.get(), sf.data(), n * 8, cudaMemcpyHostToDevice);
        sp.size = n;
        sp.global = static_cast<const float*>(xyz.get());
        sp.measurement_index_1 = static_cast<const std::uint32_t*>(meas_idx.get());
        meas.size = n;
        meas.local_position = static_cast<const float*>(local.get());
        meas.surface_link = static_cast<const std::uint64_t*>(surf.get());
    }
};

// TEST(seeding, case1) / TEST(seeding, case2)
static int seeding_case(const std::vector<float>& pts) {
    seedfinder_config finder_config;
    spacepoint_grid_config grid_config(finder_config);
    seedfilter_config filter_config;
    finder_config.deltaRMax = 100.f;       // adjusted AFTER the grid config was built
    finder_config.maxPtScattering = 0.5f;

    cudaStream_t s;
    CHECK(cudaStreamCreate(&s) == cudaSuccess);
    stream_wrapper stream(s);
    device_event ev(pts);
    triplet_seeding_algorithm sa(finder_config, grid_config, filter_config, ev.mr, stream);
    auto seeds = sa(ev.sp);
    track_params_estimation_config tpe_config;
    seed_parameter_estimation_algorithm tp(tpe_config, ev.mr, stream);
    const float B[3] = {0.f, 0.f, 2.f * unit_T};
    auto params = tp(B, ev.meas, ev.sp, seeds);
    stream.synchronize();

    std::uint32_t n_seeds = 99;
    cudaMemcpy(&n_seeds, seeds.size, 4, cudaMemcpyDeviceToHost);
    CHECK(n_seeds == 1u);  // ASSERT_EQ(seeds.size(), 1u)
    std::uint32_t b = 9, m = 9, t = 9;
    cudaMemcpy(&b, seeds.bottom_index, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&m, seeds.middle_index, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&t, seeds.top_index, 4, cudaMemcpyDeviceToHost);
    CHECK(b == 0 && m == 1 && t == 2);
    b200seed_bound_params p;
    cudaMemcpy(&p, params.params, sizeof(p), cudaMemcpyDeviceToHost);
    CHECK(p.surface_link == 100u);  // surface of the bottom spacepoint's measurement
    CHECK(std::isfinite(p.vec[4]) && p.vec[4] != 0.f && p.vec[5] == 0.f);
    CHECK(p.cov[0] == 1.f && p.cov[7] == 1.f && p.cov[35] > 0.f);
    cudaStreamDestroy(s);
    return 0;
}

// TEST(track_params_estimation, helix_negative_charge / helix_positive_charge)
static int helix_case(float q) {
    // detray helix from the origin, momentum (1,0,1) GeV, B = (0,0,2 T), s = 50/100/150 mm
    const double R = 1.0 / (0.000299792458 * 2.0);
    const double h = (q < 0) ? 1.0 : -1.0;
    std::vector<float> pts;
    for (double s : {50.0, 100.0, 150.0}) {
        const double tt = (s / std::sqrt(2.0)) / R;
        pts.push_back(static_cast<float>(h * R * std::sin(h * tt)));
        pts.push_back(static_cast<float>(-h * R * (std::cos(h * tt) - 1.0)));
        pts.push_back(static_cast<float>(R * tt));
    }
    cudaStream_t s;
    CHECK(cudaStreamCreate(&s) == cudaSuccess);
    stream_wrapper stream(s);
    device_event ev(pts);
    // seeds.push_back({0, 1, 2, 0.0f})
    seed_buffer seeds;
    seeds.capacity = 1;
    seeds.memory = device_allocation(ev.mr, 1024);
    auto* base = static_cast<unsigned char*>(seeds.memory.get());
    const std::uint32_t host[5] = {1u, 0u, 1u, 2u, 0u};
    cudaMemcpy(base, host, sizeof(host), cudaMemcpyHostToDevice);
    seeds.size = reinterpret_cast<std::uint32_t*>(base);
    seeds.bottom_index = seeds.size + 1;
    seeds.middle_index = seeds.size + 2;
    seeds.top_index = seeds.size + 3;
    seeds.quality = reinterpret_cast<float*>(seeds.size + 4);
    track_params_estimation_config cfg;
    seed_parameter_estimation_algorithm tp(cfg, ev.mr, stream);
    const float B[3] = {0.f, 0.f, 2.f * unit_T};
    auto params = tp(B, ev.meas, ev.sp, seeds);
    stream.synchronize();
    b200seed_bound_params p;
    cudaMemcpy(&p, params.params, sizeof(p), cudaMemcpyDeviceToHost);
    const float mom = 1.f / std::fabs(p.vec[4]);  // bound_params[0].p(q)
    CHECK(std::fabs(mom - std::sqrt(2.f)) < 2.f * 1e-4f);  // ASSERT_NEAR(..., 2e-4)
    CHECK((p.vec[4] < 0.f) == (q < 0.f));
    cudaStreamDestroy(s);
    return 0;
}

static int config_errors() {
    // get_axes throws std::domain_error for minHelixRadius < rMax / 2
    seedfinder_config f;
    spacepoint_grid_config g(f);
    g.minPt = 0.01f;
    cuda_device_memory_resource mr;
    bool thrown = false;
    try {
        triplet_seeding_algorithm sa(f, g, seedfilter_config(), mr, stream_wrapper(nullptr));
    } catch (const std::domain_error&) {
        thrown = true;
    }
    CHECK(thrown);
    // empty input => empty output (triplet_seeding_algorithm.cpp:75-77)
    spacepoint_grid_config g2(f);
    triplet_seeding_algorithm sa(f, g2, seedfilter_config(), mr, stream_wrapper(nullptr));
    auto seeds = sa(spacepoint_const_view{});
    cudaDeviceSynchronize();
    std::uint32_t n = 7;
    cudaMemcpy(&n, seeds.size, 4, cudaMemcpyDeviceToHost);
    CHECK(n == 0u);
    return 0;
}

// TEST(spacepoint_formation, cpu) — tests/cpu/test_spacepoint_formation.cpp:24-102: a telescope
// of nine planes along x (normal = x, local x = global y, local y = global z), a measurement
// (7, 2) on the first and (10, 15) on the last plane -> spacepoints (20, 7, 2), (180, 10, 15);
// a 1D measurement in between must be skipped.
static int formation_case() {
    cudaStream_t s;
    CHECK(cudaStreamCreate(&s) == cudaSuccess);
    stream_wrapper stream(s);
    cuda_device_memory_resource mr;
    const float pos[9] = {20.f, 40.f, 60.f, 80.f, 100.f, 120.f, 140.f, 160.f, 180.f};
    std::vector<b200seed_surface> planes(9);
    for (int i = 0; i < 9; ++i) {
        std::memset(&planes[i], 0, sizeof(b200seed_surface));
        planes[i].translation[0] = pos[i];
        planes[i].x_axis[1] = 1.f, planes[i].y_axis[2] = 1.f, planes[i].z_axis[0] = 1.f;
    }
    const float local[6] = {7.f, 2.f, 3.f, 0.f, 10.f, 15.f};
    const std::uint32_t dims[3] = {2u, 1u, 2u}, sidx[3] = {0u, 4u, 8u};
    device_allocation d_planes(mr, sizeof(b200seed_surface) * 9), d_local(mr, sizeof(local)),
        d_dims(mr, sizeof(dims)), d_sidx(mr, sizeof(sidx));
    cudaMemcpy(d_planes.get(), planes.data(), sizeof(b200seed_surface) * 9, cudaMemcpyHostToDevice);
    cudaMemcpy(d_local.get(), local, sizeof(local), cudaMemcpyHostToDevice);
    cudaMemcpy(d_dims.get(), dims, sizeof(dims), cudaMemcpyHostToDevice);
    cudaMemcpy(d_sidx.get(), sidx, sizeof(sidx), cudaMemcpyHostToDevice);
    detector_view det;
    det.n_surfaces = 9, det.surfaces = static_cast<const b200seed_surface*>(d_planes.get());
    measurement_const_view meas;
    meas.size = 3, meas.local_position = static_cast<const float*>(d_local.get());
    meas.dimensions = static_cast<const std::uint32_t*>(d_dims.get());
    meas.surface_index = static_cast<const std::uint32_t*>(d_sidx.get());
    silicon_pixel_spacepoint_formation_algorithm sp_formation(mr, stream);
    auto spacepoints = sp_formation(det, meas);
    stream.synchronize();
    std::uint32_t n = 0, mi[2] = {9, 9}, mi2[2] = {0, 0};
    float g[6], vz[2] = {1.f, 1.f};
    cudaMemcpy(&n, spacepoints.size, 4, cudaMemcpyDeviceToHost);
    CHECK(n == 2u);  // EXPECT_EQ(spacepoints.size(), 2u)
    cudaMemcpy(g, spacepoints.global, sizeof(g), cudaMemcpyDeviceToHost);
    cudaMemcpy(mi, spacepoints.measurement_index_1, sizeof(mi), cudaMemcpyDeviceToHost);
    cudaMemcpy(mi2, spacepoints.measurement_index_2, sizeof(mi2), cudaMemcpyDeviceToHost);
    cudaMemcpy(vz, spacepoints.z_variance, sizeof(vz), cudaMemcpyDeviceToHost);
    CHECK(g[0] == 20.f && g[1] == 7.f && g[2] == 2.f);
    CHECK(g[3] == 180.f && g[4] == 10.f && g[5] == 15.f);
    CHECK(mi[0] == 0u && mi[1] == 2u && mi2[0] == 0xFFFFFFFFu && vz[0] == 0.f && vz[1] == 0.f);
    // the resizable buffer feeds the seeding algorithm without a size read-back
    seedfinder_config f;
    spacepoint_grid_config gcfg(f);
    triplet_seeding_algorithm sa(f, gcfg, seedfilter_config(), mr, stream);
    auto seeds = sa(spacepoints);
    stream.synchronize();
    std::uint32_t ns = 7;
    cudaMemcpy(&ns, seeds.size, 4, cudaMemcpyDeviceToHost);
    CHECK(ns == 0u);
    // no measurements -> default-constructed buffer
    auto none = sp_formation(det, measurement_const_view{});
    CHECK(none.capacity == 0u && none.size == nullptr);
    cudaStreamDestroy(s);
    return 0;
}

int main() {
    const std::vector<float> case1 = {36.6706f, 10.6472f, 104.131f, 94.2191f, 29.6699f, 113.628f,
                                      149.805f, 47.9518f, 122.979f, 218.514f, 70.3049f, 134.029f,
                                      275.359f, 88.668f,  143.378f};
    const std::vector<float> case2 = {36.301f,  13.1197f, 106.83f,  93.9366f, 33.7101f, 120.978f,
                                      149.192f, 52.0562f, 134.678f, 218.398f, 73.1025f, 151.979f,
                                      275.322f, 89.0663f, 166.229f};
    if (seeding_case(case1)) return 1;
    std::printf("[ OK ] seeding.case1\n");
    if (seeding_case(case2)) return 1;
    std::printf("[ OK ] seeding.case2\n");
    if (helix_case(-1.f)) return 1;
    std::printf("[ OK ] track_params_estimation.helix_negative_charge\n");
    if (helix_case(1.f)) return 1;
    std::printf("[ OK ] track_params_estimation.helix_positive_charge\n");
    if (config_errors()) return 1;
    std::printf("[ OK ] config errors / empty input\n");
    if (formation_case()) return 1;
    std::printf("[ OK ] spacepoint_formation.telescope\n");
    return 0;
}
