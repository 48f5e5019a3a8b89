// ref_adapter.cu — the drop-in classes of include/traccc_b200/traccc_adapter.hpp compiled against
// the REFERENCE's real types (traccc::edm::spacepoint_collection / seed_collection /
// measurement_collection views and buffers, traccc::algorithm<>, traccc::memory_resource,
// traccc::cuda::stream_wrapper, bound_track_parameters_collection_types) from /root/reference, with
// the vecmem stand-in of oracle/shim_cuda — and run next to the reference's own
// traccc::cuda::triplet_seeding_algorithm (compiled verbatim, see ref_cuda_seeding.cu) in the call
// sequence of examples/run/cuda/apps/seeding_example_cuda.cpp:184-196,281-324:
//   spacepoints host -> device buffer; seeds = sa_cuda(view); params = tp_cuda(field, meas, sp, seeds);
//   stream.synchronize(); copy back; compare.
// TEST INFRASTRUCTURE: built into oracle/_ref/libtraccc_ref_adapter.so by `make -C oracle ref_adapter`
// (links ../traccc_b200/libb200seed.so); used by tests/test_ref_adapter.py.
#include "ref_cuda_seeding.cu"  // the reference's CUDA seeding sources + caching_resource

// the reference's host parameter estimation, verbatim (the comparator for the parameters)
#include "traccc/seeding/track_params_estimation.hpp"
#include "core/src/seeding/track_params_estimation.cpp"

#include "../include/traccc_b200/traccc_adapter.hpp"

namespace {
template <typename T>
void up(cudaStream_t s, T* d, const T* h, std::size_t n) {
    if (n) SHIM_CUDA_CHECK(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
}
}  // namespace

extern "C" {

/// Runs, on the same device-resident spacepoints:
///   which = 0: traccc::cuda::triplet_seeding_algorithm (the reference's own CUDA code)
///   which = 1: traccc::b200::triplet_seeding_algorithm + traccc::b200::seed_parameter_estimation_algorithm
/// and copies seeds (and, for which = 1, the bound_track_parameters read through their public
/// accessors) back. Returns the number of seeds, negative on error.
long ref_adapter_run(int which, const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                     const b200seed_filter_cfg* filter, const b200seed_tpe_cfg* tpe, uint32_t n,
                     const float* xyz, const float* var_z, const float* var_r, const uint32_t* sp_meas,
                     uint32_t n_meas, const float* meas_local, const uint64_t* meas_surface,
                     const float bfield[3], int resizable_input, uint32_t cap, uint32_t* bottom,
                     uint32_t* middle, uint32_t* top, float* quality, b200seed_bound_params* params) {
    try {
        cudaStream_t stream = nullptr;
        SHIM_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        caching_resource device_mr(false, true), host_mr(true, true);
        vecmem::copy copy(stream);
        const traccc::memory_resource mr{device_mr, &host_mr};
        const traccc::cuda::stream_wrapper str(stream);
        const auto f = cfg_cast<traccc::seedfinder_config>(finder);
        traccc::spacepoint_grid_config g(f);
        std::memcpy(static_cast<void*>(&g), grid, sizeof(g));
        const auto fl = cfg_cast<traccc::seedfilter_config>(filter);
        traccc::track_params_estimation_config tc;
        std::memcpy(static_cast<void*>(&tc), tpe, sizeof(tc));

        // ---- the event in device buffers (seeding_example_cuda.cpp:264-279) ----
        traccc::edm::spacepoint_collection::buffer sps(
            n, device_mr,
            resizable_input ? vecmem::data::buffer_type::resizable : vecmem::data::buffer_type::fixed_size);
        copy.setup(sps)->ignore();
        {
            std::vector<unsigned int> m1(n), m2(n, 0xFFFFFFFFu);
            std::vector<float> zero(n, 0.f);
            for (uint32_t i = 0; i < n; ++i) m1[i] = sp_meas ? sp_meas[i] : i;
            up(stream, sps.get<0>().ptr(), m1.data(), n);
            up(stream, sps.get<1>().ptr(), m2.data(), n);
            up(stream, reinterpret_cast<float*>(sps.get<2>().ptr()), xyz, 3ul * n);
            up(stream, sps.get<3>().ptr(), var_z ? var_z : zero.data(), n);
            up(stream, sps.get<4>().ptr(), var_r ? var_r : zero.data(), n);
            if (resizable_input) up(stream, sps.get<0>().size_ptr(), &n, 1);
            SHIM_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
        traccc::edm::measurement_collection::buffer meas(n_meas, device_mr);
        copy.setup(meas)->ignore();
        up(stream, reinterpret_cast<float*>(meas.get<0>().ptr()), meas_local, 2ul * n_meas);
        up(stream, reinterpret_cast<uint64_t*>(meas.get<6>().ptr()), meas_surface, n_meas);
        SHIM_CUDA_CHECK(cudaStreamSynchronize(stream));
        const traccc::edm::spacepoint_collection::const_view sp_view(sps);
        const traccc::edm::measurement_collection::const_view meas_view(meas);

        // ---- the algorithms, as the example constructs them (:184-196) ----
        traccc::edm::seed_collection::buffer seeds;
        traccc::bound_track_parameters_collection_types::buffer pars;
        if (which == 0) {
            traccc::cuda::triplet_seeding_algorithm sa_cuda(f, g, fl, mr, copy, str);
            seeds = sa_cuda(sp_view);
            str.synchronize();
        } else {
            traccc::b200::triplet_seeding_algorithm sa_b200(f, g, fl, mr, copy, str);
            traccc::b200::seed_parameter_estimation_algorithm tp_b200(tc, mr, copy, str);
            seeds = sa_b200(sp_view);                                         // (:288)
            pars = tp_b200(traccc::vector3{bfield[0], bfield[1], bfield[2]}, meas_view, sp_view,
                           traccc::edm::seed_collection::const_view(seeds));  // (:312)
            str.synchronize();                                                // (:291)
            sa_b200.check_complete();
        }
        // ---- back to the host (:329-340) ----
        const unsigned int ns = copy.get_size(seeds);
        const unsigned int m = ns < cap ? ns : cap;
        auto down = [&](void* d, const void* s, std::size_t b) {
            if (b) SHIM_CUDA_CHECK(cudaMemcpyAsync(d, s, b, cudaMemcpyDeviceToHost, stream));
        };
        down(bottom, seeds.get<0>().ptr(), m * 4ul);
        down(middle, seeds.get<1>().ptr(), m * 4ul);
        down(top, seeds.get<2>().ptr(), m * 4ul);
        down(quality, seeds.get<3>().ptr(), m * 4ul);
        std::vector<traccc::bound_track_parameters<>> hp(which == 1 ? m : 0);
        if (which == 1 && params) down(hp.data(), pars.ptr(), m * sizeof(traccc::bound_track_parameters<>));
        SHIM_CUDA_CHECK(cudaStreamSynchronize(stream));
        for (unsigned int i = 0; i < hp.size(); ++i) {
            const auto& p = hp[i];
            b200seed_bound_params& o = params[i];
            o.surface_link = p.surface_link().value();
            for (unsigned k = 0; k < 6; ++k) o.vec[k] = p[k];
            for (unsigned r = 0; r < 6; ++r)
                for (unsigned c = 0; c < 6; ++c) o.cov[r * 6 + c] = traccc::getter::element(p.covariance(), r, c);
        }
        seeds = {};
        pars = {};
        cudaStreamDestroy(stream);
        return static_cast<long>(ns);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_adapter_run: %s\n", e.what());
        return -1;
    }
}

}  // extern "C"
