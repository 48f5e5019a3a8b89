// ref_cuda_seeding.cu — the REFERENCE's own CUDA seeding algorithm
// (traccc::cuda::triplet_seeding_algorithm: the nine kernels of
// device/cuda/src/seeding/triplet_seeding_algorithm.cu around the device functions in
// device/common/include/traccc/seeding/device/impl/*.ipp, driven by the host logic of
// device/common/src/seeding/triplet_seeding_algorithm.cpp), compiled verbatim with nvcc from
// the sources under /root/reference (never copied) against the stand-in third-party headers
// in oracle/shim_cuda (CUDA-capable vecmem) and oracle/shim (detray algebra/units, Acts
// logger), with the reference's own CUDA flags (--use_fast_math --expt-relaxed-constexpr,
// C++20). TEST / BASELINE INFRASTRUCTURE: built into oracle/_ref/libtraccc_ref_cuda.so by
// `make -C oracle ref_cuda`; used by tests/test_ref_cuda.py (the CUDA path, the CPU reference
// and the reference's CUDA code must agree) and by bench.py's `reference_cuda` leg (the
// "reference CUDA seeding throughput" the north star compares with).
#include <cuda_runtime_api.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

// reference sources, verbatim --------------------------------------------------------------
#include "traccc/cuda/seeding/triplet_seeding_algorithm.hpp"
// device/common
#include "device/common/src/device/algorithm_base.cpp"
#include "device/common/src/seeding/triplet_seeding_algorithm.cpp"
// device/cuda_utils, device/cuda
#include "device/cuda_utils/src/cuda_error_handling.cpp"
#include "device/cuda_utils/src/stream_wrapper.cpp"
#undef TRACCC_CUDA_ERROR_CHECK  // device/cuda has its own copy of the macro (other namespace)
#include "device/cuda/src/utils/cuda_error_handling.cpp"
#include "device/cuda/src/utils/utils.cpp"
#include "device/cuda/src/utils/algorithm_base.cpp"
#include "device/cuda/src/seeding/triplet_seeding_algorithm.cu"

#include "../include/b200seed.h"

namespace traccc {
const Logger& getDummyLogger() {
    static const Logger l;
    return l;
}
}  // namespace traccc

namespace {

// Caching memory resources (the role vecmem::binary_page_memory_resource over
// vecmem::cuda::device_memory_resource / host_memory_resource plays in the reference's
// throughput applications, examples/run/cuda/src/full_chain_algorithm.cpp): blocks are
// rounded up to a power of two and recycled, so steady-state events make no cudaMalloc calls.
class caching_resource : public std::pmr::memory_resource {
    public:
    explicit caching_resource(bool pinned_host, bool caching) : m_host(pinned_host), m_caching(caching) {}
    ~caching_resource() override {
        for (auto& [sz, v] : m_free)
            for (void* p : v) release(p);
    }
    std::size_t upstream_calls() const { return m_upstream_calls; }

    private:
    static std::size_t round_up(std::size_t n) {
        std::size_t b = 256;
        while (b < n) b <<= 1;
        return b;
    }
    void release(void* p) {
        if (m_host)
            cudaFreeHost(p);
        else
            cudaFree(p);
    }
    void* do_allocate(std::size_t n, std::size_t) override {
        const std::size_t b = round_up(n);
        if (m_caching) {
            std::lock_guard<std::mutex> g(m_mutex);
            auto it = m_free.find(b);
            if (it != m_free.end() && !it->second.empty()) {
                void* p = it->second.back();
                it->second.pop_back();
                return p;
            }
        }
        void* p = nullptr;
        ++m_upstream_calls;
        SHIM_CUDA_CHECK(m_host ? cudaMallocHost(&p, b) : cudaMalloc(&p, b));
        return p;
    }
    void do_deallocate(void* p, std::size_t n, std::size_t) override {
        if (!m_caching) {
            release(p);
            return;
        }
        std::lock_guard<std::mutex> g(m_mutex);
        m_free[round_up(n)].push_back(p);
    }
    bool do_is_equal(const std::pmr::memory_resource& o) const noexcept override { return this == &o; }
    bool m_host, m_caching;
    std::mutex m_mutex;
    std::map<std::size_t, std::vector<void*>> m_free;
    std::size_t m_upstream_calls = 0;
};

template <typename R, typename C>
R cfg_cast(const C* c) {
    static_assert(sizeof(R) == sizeof(C));
    R r;
    std::memcpy(static_cast<void*>(&r), c, sizeof(r));
    return r;
}

struct ref_cuda {
    cudaStream_t stream = nullptr;
    caching_resource device_mr, host_mr;
    vecmem::copy copy;
    std::unique_ptr<traccc::cuda::triplet_seeding_algorithm> alg;
    std::unique_ptr<traccc::edm::spacepoint_collection::buffer> sps;
    traccc::edm::seed_collection::buffer seeds;
    ref_cuda(bool caching) : device_mr(false, caching), host_mr(true, true) {}
};

}  // namespace

extern "C" {

/// Creates a traccc::cuda::triplet_seeding_algorithm on its own stream. caching = 0 mimics
/// seeding_example_cuda (plain cudaMalloc per buffer), 1 the throughput applications.
void* refcuda_create(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                     const b200seed_filter_cfg* filter, int caching) {
    try {
        auto h = std::make_unique<ref_cuda>(caching != 0);
        SHIM_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->copy = vecmem::copy(h->stream);
        const auto f = cfg_cast<traccc::seedfinder_config>(finder);
        traccc::spacepoint_grid_config g(f);
        static_assert(sizeof(g) == sizeof(*grid));
        std::memcpy(static_cast<void*>(&g), grid, sizeof(g));
        const auto fl = cfg_cast<traccc::seedfilter_config>(filter);
        h->alg = std::make_unique<traccc::cuda::triplet_seeding_algorithm>(
            f, g, fl, traccc::memory_resource{h->device_mr, &h->host_mr}, h->copy,
            traccc::cuda::stream_wrapper(h->stream));
        return h.release();
    } catch (const std::exception&) {
        return nullptr;
    }
}

void refcuda_destroy(void* hv) {
    auto* h = static_cast<ref_cuda*>(hv);
    if (!h) return;
    cudaStreamSynchronize(h->stream);
    h->seeds = {};
    h->sps.reset();
    h->alg.reset();
    cudaStreamDestroy(h->stream);
    delete h;
}

/// Host spacepoints -> the device-resident edm::spacepoint_collection::buffer the algorithm reads.
int refcuda_upload(void* hv, uint32_t n, const float* xyz, const float* var_z, const float* var_r) {
    auto* h = static_cast<ref_cuda*>(hv);
    try {
        h->sps = std::make_unique<traccc::edm::spacepoint_collection::buffer>(n, h->device_mr);
        std::vector<unsigned int> m1(n), m2(n, 0xFFFFFFFFu);
        std::vector<float> zero(n, 0.f);
        for (uint32_t i = 0; i < n; ++i) m1[i] = i;
        auto& c = h->sps->m_cols;
        auto up = [&](void* d, const void* s, std::size_t b) {
            if (b) SHIM_CUDA_CHECK(cudaMemcpyAsync(d, s, b, cudaMemcpyHostToDevice, h->stream));
        };
        up(::cuda::std::get<0>(c).m_ptr, m1.data(), n * 4ul);
        up(::cuda::std::get<1>(c).m_ptr, m2.data(), n * 4ul);
        up(::cuda::std::get<2>(c).m_ptr, xyz, n * 12ul);
        up(::cuda::std::get<3>(c).m_ptr, var_z ? var_z : zero.data(), n * 4ul);
        up(::cuda::std::get<4>(c).m_ptr, var_r ? var_r : zero.data(), n * 4ul);
        SHIM_CUDA_CHECK(cudaStreamSynchronize(h->stream));
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

/// reps x { seeds = alg(spacepoints); stream.synchronize(); } on the resident spacepoints —
/// the timed region of seeding_example_cuda.cpp:281-291. Returns the mean wall-clock
/// milliseconds per event (the algorithm blocks on the host several times per event, so only
/// a host clock sees its whole cost), or a negative value on error.
double refcuda_run(void* hv, int reps) {
    auto* h = static_cast<ref_cuda*>(hv);
    if (!h->sps) return -1.;
    try {
        const traccc::edm::spacepoint_collection::const_view view(*h->sps);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) {
            h->seeds = (*h->alg)(view);
            SHIM_CUDA_CHECK(cudaStreamSynchronize(h->stream));
        }
        const auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double, std::milli>(t1 - t0).count() / (reps > 0 ? reps : 1);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "refcuda_run: %s\n", e.what());
        return -2.;
    }
}

/// Seeds of the last run (unordered: the reference's kernels append with atomics).
long refcuda_seeds(void* hv, uint32_t cap, uint32_t* bottom, uint32_t* middle, uint32_t* top,
                   float* quality) {
    auto* h = static_cast<ref_cuda*>(hv);
    try {
        const unsigned int n = h->copy.get_size(h->seeds);
        const unsigned int m = n < cap ? n : cap;
        auto& c = h->seeds.m_cols;
        auto down = [&](void* d, const void* s) {
            if (m) SHIM_CUDA_CHECK(cudaMemcpyAsync(d, s, m * 4ul, cudaMemcpyDeviceToHost, h->stream));
        };
        down(bottom, ::cuda::std::get<0>(c).m_ptr);
        down(middle, ::cuda::std::get<1>(c).m_ptr);
        down(top, ::cuda::std::get<2>(c).m_ptr);
        down(quality, ::cuda::std::get<3>(c).m_ptr);
        SHIM_CUDA_CHECK(cudaStreamSynchronize(h->stream));
        return static_cast<long>(n);
    } catch (const std::exception&) {
        return -1;
    }
}

/// cudaMalloc calls issued so far by the device memory resource (allocation behaviour evidence).
unsigned long refcuda_device_allocations(void* hv) {
    return static_cast<ref_cuda*>(hv)->device_mr.upstream_calls();
}

}  // extern "C"
