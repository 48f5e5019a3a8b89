// traccc_b200/seeding.hpp — header-only C++ host side above the C-ABI (b200seed.h).
//
// Mirrors the reference's algorithm classes for this path, same names / argument meaning /
// error behaviour:
//
//   traccc::cuda::triplet_seeding_algorithm
//       device/cuda/include/traccc/cuda/seeding/triplet_seeding_algorithm.hpp:23-111
//       = algorithm<edm::seed_collection::buffer(const edm::spacepoint_collection::const_view&)>
//   traccc::cuda::seed_parameter_estimation_algorithm
//       device/cuda/include/traccc/cuda/seeding/seed_parameter_estimation_algorithm.hpp:19-58
//       = algorithm<bound_track_parameters_collection_types::buffer(const magnetic_field&,
//             const measurement_collection::const_view&, const spacepoint_collection::const_view&,
//             const seed_collection::const_view&)>
//
// vecmem / detray are not available in this repository's build environment, so the view and
// buffer types below are minimal stand-ins with the *same shape* as the vecmem SoA views
// (capacity, optional size pointer, one contiguous column pointer per variable). INTEGRATION.md
// shows the few lines that map the real vecmem views onto them inside the reference tree.
//
// Like the reference: operator() is const, enqueues everything on the wrapped stream, and
// returns a buffer "which is not necessarily filled yet" — synchronise the stream before
// reading or destroying it. Errors are C++ exceptions (std::domain_error for the
// configurations get_axes rejects, std::runtime_error for CUDA failures).
#pragma once

#include <cuda_runtime_api.h>

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>

#include "../b200seed.h"

namespace traccc::b200 {

// ---- configuration structs: the reference PODs (byte-identical, see b200seed.h) ----------
struct seedfinder_config : b200seed_finder_cfg {
    seedfinder_config() { b200seed_finder_cfg_defaults(this); }
    void setup() { b200seed_finder_cfg_setup(this); }
};
struct spacepoint_grid_config : b200seed_grid_cfg {
    spacepoint_grid_config() = delete;
    explicit spacepoint_grid_config(const seedfinder_config& f) {
        b200seed_grid_cfg_from_finder(&f, this);
    }
};
struct seedfilter_config : b200seed_filter_cfg {
    seedfilter_config() { b200seed_filter_cfg_defaults(this); }
};
struct track_params_estimation_config : b200seed_tpe_cfg {
    track_params_estimation_config() { b200seed_tpe_cfg_defaults(this); }
};

// ---- traccc::cuda::stream_wrapper (device/cuda_utils/.../stream_wrapper.hpp:21-50) -------
class stream_wrapper {
    public:
    explicit stream_wrapper(void* stream) : m_stream(stream) {}
    void* cudaStream() const { return m_stream; }
    void synchronize() const {
        if (cudaStreamSynchronize(static_cast<cudaStream_t>(m_stream)) != cudaSuccess)
            throw std::runtime_error("cudaStreamSynchronize failed");
    }

    private:
    void* m_stream;
};

// ---- memory resource stand-in (vecmem::memory_resource: allocate / deallocate) ----------
struct memory_resource {
    virtual ~memory_resource() = default;
    virtual void* allocate(std::size_t bytes) = 0;
    virtual void deallocate(void* p) = 0;
};
struct cuda_device_memory_resource : memory_resource {
    void* allocate(std::size_t bytes) override {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) throw std::bad_alloc();
        return p;
    }
    void deallocate(void* p) override { cudaFree(p); }
};
class device_allocation {
    public:
    device_allocation() = default;
    device_allocation(memory_resource& mr, std::size_t bytes)
        : m_mr(&mr), m_ptr(mr.allocate(bytes)), m_bytes(bytes) {}
    device_allocation(device_allocation&& o) noexcept { *this = std::move(o); }
    device_allocation& operator=(device_allocation&& o) noexcept {
        reset();
        m_mr = o.m_mr, m_ptr = o.m_ptr, m_bytes = o.m_bytes;
        o.m_ptr = nullptr;
        return *this;
    }
    ~device_allocation() { reset(); }
    void* get() const { return m_ptr; }
    std::size_t bytes() const { return m_bytes; }

    private:
    void reset() {
        if (m_ptr) m_mr->deallocate(m_ptr);
        m_ptr = nullptr;
    }
    memory_resource* m_mr = nullptr;
    void* m_ptr = nullptr;
    std::size_t m_bytes = 0;
};

// ---- views / buffers (shape of the vecmem::edm views of the reference collections) -------
// edm::spacepoint_collection::const_view — core/include/traccc/edm/spacepoint_collection.hpp:223-234
struct spacepoint_const_view {
    std::uint32_t size = 0;
    // resizable buffers (made on the device by the spacepoint formation): the number of filled
    // entries is this device word and `size` is the capacity; nullptr for fixed-size views
    const std::uint32_t* size_ptr = nullptr;
    const std::uint32_t* measurement_index_1 = nullptr;
    const float* global = nullptr;  // std::array<float,3> per spacepoint
    const float* z_variance = nullptr;
    const float* radius_variance = nullptr;
};
// edm::measurement_collection::const_view — the two columns this path reads (:241-260)
struct measurement_const_view {
    std::uint32_t size = 0;
    const float* local_position = nullptr;        // std::array<float,2>
    const std::uint64_t* surface_link = nullptr;  // detray::geometry::identifier
    const std::uint32_t* dimensions = nullptr;    // nullptr: all measurements are 2D
    // index of the measurement's surface in the detector's surface table (the index field of
    // surface_link); only the spacepoint formation reads it
    const std::uint32_t* surface_index = nullptr;
};
// the part of a detray detector view the spacepoint formation needs: the placed surfaces
struct detector_view {
    std::uint32_t n_surfaces = 0;
    const b200seed_surface* surfaces = nullptr;  // device memory
};
// edm::spacepoint_collection::buffer (resizable) returned by the spacepoint formation
struct spacepoint_buffer {
    std::uint32_t capacity = 0;
    std::uint32_t* size = nullptr;  // device: the buffer's size word
    float* global = nullptr;
    float* z_variance = nullptr;
    float* radius_variance = nullptr;
    std::uint32_t* measurement_index_1 = nullptr;
    std::uint32_t* measurement_index_2 = nullptr;
    device_allocation memory;
    /// vecmem::get_data(buffer): the view the seeding algorithms take
    operator spacepoint_const_view() const {
        spacepoint_const_view v;
        v.size = capacity, v.size_ptr = size, v.measurement_index_1 = measurement_index_1;
        v.global = global, v.z_variance = z_variance, v.radius_variance = radius_variance;
        return v;
    }
};
// edm::seed_collection::buffer (resizable) — core/include/traccc/edm/seed_collection.hpp:142-146
struct seed_buffer {
    std::uint32_t capacity = 0;
    std::uint32_t* size = nullptr;  // device: the buffer's size word
    std::uint32_t* bottom_index = nullptr;
    std::uint32_t* middle_index = nullptr;
    std::uint32_t* top_index = nullptr;
    float* quality = nullptr;
    b200seed_counters* counters = nullptr;  // device: per-event counters (logging / parity)
    device_allocation memory;
};
using seed_const_view = seed_buffer;
// bound_track_parameters_collection_types::buffer
struct bound_track_parameters_buffer {
    std::uint32_t capacity = 0;
    b200seed_bound_params* params = nullptr;
    device_allocation memory;
};

namespace detail {
struct handle_deleter {
    void operator()(b200seed_handle* h) const { b200seed_destroy(h); }
};
inline std::unique_ptr<b200seed_handle, handle_deleter> make_handle(
    const b200seed_finder_cfg& f, const b200seed_grid_cfg& g, const b200seed_filter_cfg& fl,
    const b200seed_tpe_cfg* t) {
    int dev = 0;
    cudaGetDevice(&dev);
    b200seed_handle* h = nullptr;
    const int rc = b200seed_create(&f, &g, &fl, t, dev, &h);
    if (rc == B200SEED_EINVAL) throw std::domain_error(b200seed_last_error(nullptr));
    if (rc != B200SEED_OK) throw std::runtime_error(b200seed_last_error(nullptr));
    return std::unique_ptr<b200seed_handle, handle_deleter>(h);
}
inline void check(int rc, const b200seed_handle* h) {
    if (rc == B200SEED_OK) return;
    const std::string msg = b200seed_last_error(h);
    if (rc == B200SEED_EINVAL) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
inline std::size_t up256(std::size_t v) { return (v + 255) / 256 * 256; }
}  // namespace detail

/// Drop-in for traccc::cuda::triplet_seeding_algorithm.
class triplet_seeding_algorithm {
    public:
    using output_type = seed_buffer;

    triplet_seeding_algorithm(const seedfinder_config& finder_config,
                              const spacepoint_grid_config& grid_config,
                              const seedfilter_config& filter_config, memory_resource& mr,
                              const stream_wrapper& str)
        : m_handle(detail::make_handle(finder_config, grid_config, filter_config, nullptr)),
          m_max_seeds_per_spm(finder_config.maxSeedsPerSpM ? finder_config.maxSeedsPerSpM : 1),
          m_mr(&mr),
          m_stream(str) {}

    /// spacepoints in, seeds out (device memory, asynchronous).
    output_type operator()(const spacepoint_const_view& spacepoints) const {
        seed_buffer out;
        const std::uint32_t n = spacepoints.size;
        // capacity bound: at most maxSeedsPerSpM seeds per (valid) middle spacepoint
        out.capacity = n * m_max_seeds_per_spm;
        const std::size_t cap = out.capacity ? out.capacity : 1;
        const std::size_t o_cols = detail::up256(sizeof(b200seed_counters) + 256);
        out.memory = device_allocation(*m_mr, o_cols + 4 * detail::up256(cap * 4));
        auto* base = static_cast<unsigned char*>(out.memory.get());
        out.size = reinterpret_cast<std::uint32_t*>(base);
        out.counters = reinterpret_cast<b200seed_counters*>(base + 256);
        out.bottom_index = reinterpret_cast<std::uint32_t*>(base + o_cols);
        out.middle_index = reinterpret_cast<std::uint32_t*>(base + o_cols + detail::up256(cap * 4));
        out.top_index = reinterpret_cast<std::uint32_t*>(base + o_cols + 2 * detail::up256(cap * 4));
        out.quality = reinterpret_cast<float*>(base + o_cols + 3 * detail::up256(cap * 4));
        // scratch lives as long as this algorithm object (grown on demand, reused per event)
        const std::size_t need = b200seed_workspace_bytes(m_handle.get(), n);
        if (need > m_workspace.bytes()) {
            // kernels of the previous event may still use the old scratch on this stream: a
            // caching / pool memory_resource would hand it out again right away
            if (m_workspace.bytes()) cudaStreamSynchronize(static_cast<cudaStream_t>(m_stream.cudaStream()));
            m_workspace = device_allocation(*m_mr, need + need / 4);
        }
        if (spacepoints.size_ptr && n) {
            // resizable input: its size stays on the device, no D->H read
            detail::check(
                b200seed_run_n_on_device(m_handle.get(), m_stream.cudaStream(), n,
                                         spacepoints.size_ptr, spacepoints.global,
                                         spacepoints.z_variance, spacepoints.radius_variance,
                                         m_workspace.get(), m_workspace.bytes(), out.capacity,
                                         out.bottom_index, out.middle_index, out.top_index,
                                         out.quality, out.size, out.counters),
                m_handle.get());
            return out;
        }
        detail::check(b200seed_run(m_handle.get(), m_stream.cudaStream(), n, spacepoints.global,
                                   spacepoints.z_variance, spacepoints.radius_variance,
                                   m_workspace.get(), m_workspace.bytes(), out.capacity,
                                   out.bottom_index, out.middle_index, out.top_index, out.quality,
                                   out.size, out.counters),
                      m_handle.get());
        return out;
    }

    /// After the caller has synchronised the stream: throws std::runtime_error if an event run
    /// through this object since the last call was truncated by a capacity-bounded buffer (the
    /// reference never truncates; see B200SEED_EOVERFLOW).
    void check_complete() const { detail::check(b200seed_check_overflow(m_handle.get(), nullptr), m_handle.get()); }

    private:
    std::unique_ptr<b200seed_handle, detail::handle_deleter> m_handle;
    std::uint32_t m_max_seeds_per_spm;
    memory_resource* m_mr;
    stream_wrapper m_stream;
    mutable device_allocation m_workspace;
};

/// Drop-in for traccc::cuda::silicon_pixel_spacepoint_formation_algorithm
/// (device/cuda/include/traccc/cuda/seeding/silicon_pixel_spacepoint_formation_algorithm.hpp;
/// common part device/common/src/seeding/silicon_pixel_spacepoint_formation_algorithm.cpp:20-52)
/// = algorithm<edm::spacepoint_collection::buffer(const detector_buffer&,
///                                               const edm::measurement_collection::const_view&)>.
/// Spacepoints come out in measurement order (the reference's host algorithm's order).
class silicon_pixel_spacepoint_formation_algorithm {
    public:
    using output_type = spacepoint_buffer;

    silicon_pixel_spacepoint_formation_algorithm(memory_resource& mr, const stream_wrapper& str)
        : m_mr(&mr), m_stream(str) {
        const seedfinder_config f;
        m_handle = detail::make_handle(f, spacepoint_grid_config(f), seedfilter_config(), nullptr);
    }

    output_type operator()(const detector_view& det,
                           const measurement_const_view& measurements) const {
        spacepoint_buffer out;
        const std::uint32_t n = measurements.size;
        if (n == 0) return out;  // "If there are no measurements, return right away" (:37-40)
        out.capacity = n;
        const std::size_t col = detail::up256(std::size_t(n) * 4);
        out.memory = device_allocation(*m_mr, 256 + detail::up256(std::size_t(n) * 12) + 4 * col);
        auto* base = static_cast<unsigned char*>(out.memory.get());
        out.size = reinterpret_cast<std::uint32_t*>(base);
        out.global = reinterpret_cast<float*>(base + 256);
        unsigned char* c = base + 256 + detail::up256(std::size_t(n) * 12);
        out.z_variance = reinterpret_cast<float*>(c);
        out.radius_variance = reinterpret_cast<float*>(c + col);
        out.measurement_index_1 = reinterpret_cast<std::uint32_t*>(c + 2 * col);
        out.measurement_index_2 = reinterpret_cast<std::uint32_t*>(c + 3 * col);
        detail::check(b200seed_form_spacepoints(
                          m_handle.get(), m_stream.cudaStream(), n, measurements.local_position,
                          measurements.dimensions, measurements.surface_index, det.surfaces,
                          det.n_surfaces, out.global, out.z_variance, out.radius_variance,
                          out.measurement_index_1, out.measurement_index_2, out.size),
                      m_handle.get());
        return out;
    }

    private:
    std::unique_ptr<b200seed_handle, detail::handle_deleter> m_handle;
    memory_resource* m_mr;
    stream_wrapper m_stream;
};

/// Drop-in for traccc::cuda::seed_parameter_estimation_algorithm (homogeneous field).
class seed_parameter_estimation_algorithm {
    public:
    using output_type = bound_track_parameters_buffer;

    seed_parameter_estimation_algorithm(const track_params_estimation_config& config,
                                        memory_resource& mr, const stream_wrapper& str)
        : m_mr(&mr), m_stream(str) {
        const seedfinder_config f;
        m_handle = detail::make_handle(f, spacepoint_grid_config(f), seedfilter_config(), &config);
    }

    /// bfield: the homogeneous field vector (what covfie's constant field returns everywhere).
    output_type operator()(const float bfield[3], const measurement_const_view& measurements,
                           const spacepoint_const_view& spacepoints,
                           const seed_const_view& seeds) const {
        bound_track_parameters_buffer out;
        out.capacity = seeds.capacity;
        out.memory = device_allocation(
            *m_mr, (out.capacity ? out.capacity : 1) * sizeof(b200seed_bound_params));
        out.params = static_cast<b200seed_bound_params*>(out.memory.get());
        detail::check(
            b200seed_estimate_params(m_handle.get(), m_stream.cudaStream(), seeds.size,
                                     seeds.capacity, seeds.bottom_index, seeds.middle_index,
                                     seeds.top_index, spacepoints.global,
                                     spacepoints.measurement_index_1, measurements.local_position,
                                     measurements.surface_link, bfield, out.params),
            m_handle.get());
        return out;
    }

    private:
    std::unique_ptr<b200seed_handle, detail::handle_deleter> m_handle;
    memory_resource* m_mr;
    stream_wrapper m_stream;
};

}  // namespace traccc::b200
