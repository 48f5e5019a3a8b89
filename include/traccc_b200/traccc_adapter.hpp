// traccc_adapter.hpp — the drop-in classes written against traccc's REAL types.
//
// traccc::b200::triplet_seeding_algorithm and traccc::b200::seed_parameter_estimation_algorithm
// have the constructor and operator() shapes of traccc::cuda::triplet_seeding_algorithm
// (device/cuda/include/traccc/cuda/seeding/triplet_seeding_algorithm.hpp:23-111 over
// device/common/.../triplet_seeding_algorithm.hpp:37-69) and
// traccc::cuda::seed_parameter_estimation_algorithm (…/seed_parameter_estimation_algorithm.hpp:19-58
// over device/common/.../seed_parameter_estimation_algorithm.hpp:31-67): they take
// edm::spacepoint_collection::const_view / edm::measurement_collection::const_view /
// edm::seed_collection::const_view, allocate their outputs from traccc::memory_resource::main and
// return edm::seed_collection::buffer / bound_track_parameters_collection_types::buffer without
// synchronising. Inside they hand the column pointers of the SoA views to libb200seed.so.
//
// This header needs traccc's and vecmem's headers, i.e. it is compiled inside a traccc tree (nvcc,
// C++20: it contains one small kernel). In this repository it is compiled against /root/reference
// with the vecmem stand-in of oracle/shim_cuda by oracle/ref_adapter.cu, which runs the call
// sequence of examples/run/cuda/apps/seeding_example_cuda.cpp:184-196,281-324 with the reference's
// own CUDA algorithm and with these classes side by side (tests/test_ref_adapter.py).
//
// Only public vecmem API is used: view.get<I>() -> vecmem::data::vector_view<T> with ptr(),
// capacity(), size_ptr(); buffer(capacity, mr, buffer_type::resizable); copy.setup(); copy.get_size().
#pragma once

#include <b200seed.h>

#include <memory>
#include <stdexcept>
#include <string>

#include "traccc/cuda/utils/stream_wrapper.hpp"
#include "traccc/edm/measurement_collection.hpp"
#include "traccc/edm/seed_collection.hpp"
#include "traccc/edm/spacepoint_collection.hpp"
#include "traccc/edm/track_parameters.hpp"
#include "traccc/seeding/detail/seeding_config.hpp"
#include "traccc/seeding/detail/track_params_estimation_config.hpp"
#include "traccc/utils/algorithm.hpp"
#include "traccc/utils/memory_resource.hpp"

#include <vecmem/containers/data/vector_buffer.hpp>
#include <vecmem/memory/unique_ptr.hpp>
#include <vecmem/utils/copy.hpp>

namespace traccc::b200 {

namespace detail {
struct handle_deleter {
    void operator()(b200seed_handle* h) const { b200seed_destroy(h); }
};
using handle_ptr = std::unique_ptr<b200seed_handle, handle_deleter>;

// The configuration PODs are byte-for-byte the b200seed_*_cfg structs (size and offsets are
// static_asserted against the reference headers in oracle/ref_probe.cpp).
inline handle_ptr make_handle(const seedfinder_config& f, const spacepoint_grid_config& g,
                              const seedfilter_config& fl, const track_params_estimation_config* t,
                              int device) {
    static_assert(sizeof(seedfinder_config) == sizeof(b200seed_finder_cfg));
    static_assert(sizeof(spacepoint_grid_config) == sizeof(b200seed_grid_cfg));
    static_assert(sizeof(seedfilter_config) == sizeof(b200seed_filter_cfg));
    static_assert(sizeof(track_params_estimation_config) == sizeof(b200seed_tpe_cfg));
    b200seed_handle* h = nullptr;
    const int rc = b200seed_create(reinterpret_cast<const b200seed_finder_cfg*>(&f),
                                   reinterpret_cast<const b200seed_grid_cfg*>(&g),
                                   reinterpret_cast<const b200seed_filter_cfg*>(&fl),
                                   reinterpret_cast<const b200seed_tpe_cfg*>(t), device, &h);
    // get_axes throws std::domain_error for impossible grids (spacepoint_binning_helper.hpp:35-62)
    if (rc == B200SEED_EINVAL) throw std::domain_error(b200seed_last_error(nullptr));
    if (rc != B200SEED_OK) throw std::runtime_error(b200seed_last_error(nullptr));
    return handle_ptr(h);
}
inline void check(int rc, const b200seed_handle* h) {
    if (rc != B200SEED_OK) throw std::runtime_error(b200seed_last_error(h));
}

#if defined(__CUDACC__)
// bound_track_parameters<> is a third-party type (detray): it is filled through its public
// setters, so nothing here depends on its member layout.
__global__ void fill_bound_params(const unsigned int* n_seeds, unsigned int capacity,
                                  const b200seed_bound_params_diag* in,
                                  vecmem::data::vector_view<bound_track_parameters<>> out_view) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n = *n_seeds < capacity ? *n_seeds : capacity;
    if (i >= n) return;
    bound_track_parameters<>* out = out_view.ptr();
    bound_track_parameters<>& p = out[i];
    new (&p) bound_track_parameters<>();  // zero covariance, like estimate_track_params.ipp:54
    const b200seed_bound_params_diag r = in[i];
    p.set_surface_link(detray::geometry::identifier{r.surface_link});
    p.set_bound_local({r.vec[0], r.vec[1]});
    p.set_phi(r.vec[2]);
    p.set_theta(r.vec[3]);
    p.set_qop(r.vec[4]);
    p.set_time(r.vec[5]);
    for (unsigned int j = 0; j < e_bound_size; ++j) getter::element(p.covariance(), j, j) = r.cov_diag[j];
}
#endif
}  // namespace detail

/// Drop-in for traccc::cuda::triplet_seeding_algorithm (same constructor arguments minus the logger).
class triplet_seeding_algorithm
    : public algorithm<edm::seed_collection::buffer(const edm::spacepoint_collection::const_view&)> {
    public:
    triplet_seeding_algorithm(const seedfinder_config& finder_config,
                              const spacepoint_grid_config& grid_config,
                              const seedfilter_config& filter_config, const traccc::memory_resource& mr,
                              const vecmem::copy& copy, const cuda::stream_wrapper& str)
        : m_handle(detail::make_handle(finder_config, grid_config, filter_config, nullptr, str.device())),
          m_k(finder_config.maxSeedsPerSpM ? finder_config.maxSeedsPerSpM : 1u),
          m_mr(mr),
          m_copy(copy),
          m_stream(str) {}

    /// Spacepoints in, seeds out; the returned buffer is not necessarily filled yet
    /// (triplet_seeding_algorithm.hpp:31-35).
    output_type operator()(const edm::spacepoint_collection::const_view& spacepoints) const override {
        // the capacity is known on the host; the filled size of a resizable buffer stays on the
        // device (the reference reads it back here, triplet_seeding_algorithm.cpp:64-72)
        const unsigned int n = spacepoints.capacity();
        if (n == 0) return {};  // "If there are no spacepoints, return right away" (:75-77)
        const auto& xyz = spacepoints.template get<2>();   // std::array<float, 3> per spacepoint
        const auto& var_z = spacepoints.template get<3>();
        const auto& var_r = spacepoints.template get<4>();
        edm::seed_collection::buffer seeds(n * m_k, m_mr.main, vecmem::data::buffer_type::resizable);
        m_copy.get().setup(seeds)->ignore();
        const std::size_t need = b200seed_workspace_bytes(m_handle.get(), n);
        if (need > m_ws_bytes) {
            // kernels of the previous event may still read the old scratch
            if (m_ws_bytes) m_stream.synchronize();
            m_ws = vecmem::make_unique_alloc<char[]>(m_mr.main, need + need / 4 + 256);
            m_ws_bytes = need + need / 4;
        }
        char* ws = m_ws.get();
        ws += (256 - reinterpret_cast<std::uintptr_t>(ws) % 256) % 256;
        const unsigned int* size_word = xyz.size_ptr();  // nullptr for a fixed-size buffer
        const float* p_xyz = reinterpret_cast<const float*>(xyz.ptr());
        auto& b = seeds.template get<0>();
        auto& m = seeds.template get<1>();
        auto& t = seeds.template get<2>();
        auto& q = seeds.template get<3>();
        const int rc =
            size_word ? b200seed_run_n_on_device(m_handle.get(), m_stream.cudaStream(), n, size_word, p_xyz,
                                                 var_z.ptr(), var_r.ptr(), ws, m_ws_bytes, n * m_k, b.ptr(),
                                                 m.ptr(), t.ptr(), q.ptr(), b.size_ptr(), nullptr)
                      : b200seed_run(m_handle.get(), m_stream.cudaStream(), n, p_xyz, var_z.ptr(), var_r.ptr(),
                                     ws, m_ws_bytes, n * m_k, b.ptr(), m.ptr(), t.ptr(), q.ptr(), b.size_ptr(),
                                     nullptr);
        detail::check(rc, m_handle.get());
        return seeds;
    }

    /// After the caller synchronised the stream: throws if an event was truncated by a capacity
    /// bound (B200SEED_EOVERFLOW); the reference never truncates.
    void check_complete() const {
        detail::check(b200seed_check_overflow(m_handle.get(), nullptr), m_handle.get());
    }

    private:
    detail::handle_ptr m_handle;
    unsigned int m_k;
    traccc::memory_resource m_mr;
    std::reference_wrapper<const vecmem::copy> m_copy;
    cuda::stream_wrapper m_stream;
    mutable vecmem::unique_alloc_ptr<char[]> m_ws;
    mutable std::size_t m_ws_bytes = 0;
};

#if defined(__CUDACC__)
/// Drop-in for traccc::cuda::seed_parameter_estimation_algorithm. The reference takes a covfie
/// `magnetic_field` (absent here); this class takes the field VECTOR of a homogeneous field —
/// `field.at(x, y, z)` of covfie's constant backend — and the inhomogeneous case goes through
/// b200seed_estimate_params_inhom with the grid's affine map and data.
class seed_parameter_estimation_algorithm
    : public algorithm<bound_track_parameters_collection_types::buffer(
          const vector3&, const edm::measurement_collection::const_view&,
          const edm::spacepoint_collection::const_view&, const edm::seed_collection::const_view&)> {
    public:
    seed_parameter_estimation_algorithm(const track_params_estimation_config& config,
                                        const traccc::memory_resource& mr, const vecmem::copy& copy,
                                        const cuda::stream_wrapper& str)
        : m_mr(mr), m_copy(copy), m_stream(str) {
        const seedfinder_config f;
        m_handle = detail::make_handle(f, spacepoint_grid_config(f), seedfilter_config(), &config, str.device());
    }

    output_type operator()(const vector3& bfield, const edm::measurement_collection::const_view& measurements,
                           const edm::spacepoint_collection::const_view& spacepoints,
                           const edm::seed_collection::const_view& seeds) const override {
        const unsigned int cap = seeds.capacity();
        if (cap == 0 || spacepoints.capacity() == 0) return {};  // (:49-51)
        bound_track_parameters_collection_types::buffer result(cap, m_mr.main);
        m_copy.get().setup(result)->ignore();
        static_assert(sizeof(detray::geometry::identifier) == sizeof(std::uint64_t));
        const auto& sb = seeds.template get<0>();
        cudaStream_t s = static_cast<cudaStream_t>(m_stream.cudaStream());
        // scratch of this algorithm object (grown on demand, reused per event): the diagonal records
        // the library writes, and a size word for fixed-size seed buffers
        if (cap > m_scratch_cap) {
            if (m_scratch_cap) m_stream.synchronize();  // the previous event may still read it
            m_scratch = vecmem::make_unique_alloc<b200seed_bound_params_diag[]>(m_mr.main, cap + cap / 4);
            m_scratch_cap = cap + cap / 4;
        }
        const float bf[3] = {bfield[0], bfield[1], bfield[2]};
        const unsigned int* n_dev = sb.size_ptr();
        if (!n_dev) {  // a fixed-size seed buffer has no size word: its capacity is its size
            if (!m_fixed_n) m_fixed_n = vecmem::make_unique_alloc<unsigned int>(m_mr.main);
            cudaMemcpyAsync(m_fixed_n.get(), &cap, sizeof(cap), cudaMemcpyHostToDevice, s);
            n_dev = m_fixed_n.get();
        }
        detail::check(
            b200seed_estimate_params_diag(
                m_handle.get(), s, n_dev, cap, sb.ptr(), seeds.template get<1>().ptr(),
                seeds.template get<2>().ptr(),
                reinterpret_cast<const float*>(spacepoints.template get<2>().ptr()),
                spacepoints.template get<0>().ptr(),
                reinterpret_cast<const float*>(measurements.template get<0>().ptr()),
                reinterpret_cast<const std::uint64_t*>(measurements.template get<6>().ptr()), bf,
                m_scratch.get()),
            m_handle.get());
        detail::fill_bound_params<<<(cap + 127) / 128, 128, 0, s>>>(n_dev, cap, m_scratch.get(), result);
        return result;
    }

    private:
    detail::handle_ptr m_handle;
    traccc::memory_resource m_mr;
    std::reference_wrapper<const vecmem::copy> m_copy;
    cuda::stream_wrapper m_stream;
    mutable vecmem::unique_alloc_ptr<b200seed_bound_params_diag[]> m_scratch;
    mutable std::size_t m_scratch_cap = 0;
    mutable vecmem::unique_alloc_ptr<unsigned int> m_fixed_n;
};
#endif

}  // namespace traccc::b200
