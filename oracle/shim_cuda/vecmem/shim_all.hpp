#pragma once
// oracle/shim_cuda — CUDA-capable stand-in for vecmem 1.25.0 (absent offline).
// TEST / BASELINE INFRASTRUCTURE. Just enough of vecmem's public surface — views, buffers
// (fixed-size and resizable), device vectors with atomic push_back, jagged buffers, the SoA
// edm::container, an asynchronous copy object bound to a CUDA stream, unique_alloc_ptr,
// device_atomic_ref — for the reference's *unmodified* CUDA seeding sources
//     device/common/src/seeding/triplet_seeding_algorithm.cpp
//     device/cuda/src/seeding/triplet_seeding_algorithm.cu   (+ the device/*.ipp kernels)
// to compile with nvcc where they lie under /root/reference (oracle/ref_cuda_seeding.cu).
// Nothing here is the reference's code and nothing here carries seeding arithmetic; where the
// stand-in is cheaper than the real library (no CUDA event per copy, no bounds checks) the
// difference favours the reference's timing.
#include <cuda_runtime_api.h>

#include <array>
#include <cassert>
#include <cstddef>
#include <cstring>
#include <memory>
#include <memory_resource>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include <cuda/std/tuple>

#if defined(__CUDACC__)
#define VECMEM_HOST_AND_DEVICE __host__ __device__
#define VECMEM_HOST __host__
#define VECMEM_DEVICE __device__
#else
#define VECMEM_HOST_AND_DEVICE
#define VECMEM_HOST
#define VECMEM_DEVICE
#endif

#define SHIM_CUDA_CHECK(EXP)                                                                  \
    do {                                                                                      \
        cudaError_t e_ = (EXP);                                                               \
        if (e_ != cudaSuccess)                                                                \
            throw std::runtime_error(std::string(#EXP) + ": " + cudaGetErrorString(e_));      \
    } while (0)

namespace vecmem {

using memory_resource = std::pmr::memory_resource;
template <typename T>
using vector = std::vector<T, std::pmr::polymorphic_allocator<T>>;
template <typename T>
using jagged_vector = vector<vector<T>>;

// ---------------------------------------------------------------------------------------
// unique_alloc_ptr / make_unique_alloc (vecmem/memory/unique_ptr.hpp)
// ---------------------------------------------------------------------------------------
namespace details {
template <typename T>
struct alloc_deleter {
    memory_resource* mr = nullptr;
    std::size_t bytes = 0;
    void operator()(std::remove_extent_t<T>* p) const {
        if (p && mr) mr->deallocate(p, bytes, 256);
    }
};
}  // namespace details
template <typename T>
using unique_alloc_ptr = std::unique_ptr<T, details::alloc_deleter<T>>;
template <typename T, std::enable_if_t<!std::is_array_v<T>, bool> = true>
unique_alloc_ptr<T> make_unique_alloc(memory_resource& mr) {
    return unique_alloc_ptr<T>(static_cast<T*>(mr.allocate(sizeof(T), 256)),
                               details::alloc_deleter<T>{&mr, sizeof(T)});
}
template <typename T, std::enable_if_t<std::is_array_v<T>, bool> = true>
unique_alloc_ptr<T> make_unique_alloc(memory_resource& mr, std::size_t n) {
    using E = std::remove_extent_t<T>;
    const std::size_t bytes = (n ? n : 1) * sizeof(E);
    return unique_alloc_ptr<T>(static_cast<E*>(mr.allocate(bytes, 256)),
                               details::alloc_deleter<T>{&mr, bytes});
}

// ---------------------------------------------------------------------------------------
// device_atomic_ref (vecmem/memory/device_atomic_ref.hpp)
// ---------------------------------------------------------------------------------------
template <typename T>
class device_atomic_ref {
    public:
    VECMEM_HOST_AND_DEVICE explicit device_atomic_ref(T& r) : m_ptr(&r) {}
    VECMEM_HOST_AND_DEVICE T fetch_add(T v) const {
#ifdef __CUDA_ARCH__
        return atomicAdd(m_ptr, v);
#else
        T o = *m_ptr;
        *m_ptr += v;
        return o;
#endif
    }
    VECMEM_HOST_AND_DEVICE T fetch_sub(T v) const {
#ifdef __CUDA_ARCH__
        return atomicSub(m_ptr, v);
#else
        T o = *m_ptr;
        *m_ptr -= v;
        return o;
#endif
    }
    VECMEM_HOST_AND_DEVICE T load() const { return *static_cast<volatile T*>(m_ptr); }
    VECMEM_HOST_AND_DEVICE void store(T v) const { *static_cast<volatile T*>(m_ptr) = v; }

    private:
    T* m_ptr;
};

namespace data {
enum class buffer_type { fixed_size = 0, resizable = 1 };

// ---------------------------------------------------------------------------------------
// vector_view / vector_buffer
// ---------------------------------------------------------------------------------------
template <typename T>
struct vector_view {
    using size_type = unsigned int;
    using size_pointer = std::conditional_t<std::is_const_v<T>, const size_type*, size_type*>;
    size_type m_capacity = 0;
    size_pointer m_size = nullptr;
    T* m_ptr = nullptr;
    vector_view() = default;
    VECMEM_HOST_AND_DEVICE vector_view(size_type n, T* p) : m_capacity(n), m_ptr(p) {}
    VECMEM_HOST_AND_DEVICE vector_view(size_type n, size_pointer s, T* p)
        : m_capacity(n), m_size(s), m_ptr(p) {}
    template <typename U,
              std::enable_if_t<std::is_same_v<std::remove_cv_t<T>, std::remove_cv_t<U>> &&
                                   !std::is_same_v<T, U>,
                               bool> = true>
    VECMEM_HOST_AND_DEVICE vector_view(const vector_view<U>& o)
        : m_capacity(o.m_capacity), m_size(o.m_size), m_ptr(o.m_ptr) {}
    VECMEM_HOST_AND_DEVICE size_type size() const { return m_size ? *m_size : m_capacity; }
    VECMEM_HOST_AND_DEVICE size_type capacity() const { return m_capacity; }
    VECMEM_HOST_AND_DEVICE size_pointer size_ptr() const { return m_size; }
    VECMEM_HOST_AND_DEVICE T* ptr() const { return m_ptr; }
};

template <typename T>
struct vector_buffer : vector_view<T> {
    using size_type = unsigned int;
    vector_buffer() = default;
    vector_buffer(size_type capacity, memory_resource& mr,
                  buffer_type type = buffer_type::fixed_size) {
        // one allocation: [size word, padded to 256 B][payload]
        const std::size_t head = (type == buffer_type::resizable) ? 256 : 0;
        const std::size_t bytes = head + std::size_t(capacity ? capacity : 1) * sizeof(T);
        m_mem = make_unique_alloc<char[]>(mr, bytes);
        this->m_capacity = capacity;
        this->m_size = head ? reinterpret_cast<size_type*>(m_mem.get()) : nullptr;
        this->m_ptr = reinterpret_cast<T*>(m_mem.get() + head);
    }
    vector_buffer(vector_buffer&&) = default;
    vector_buffer& operator=(vector_buffer&&) = default;
    unique_alloc_ptr<char[]> m_mem;
};

// ---------------------------------------------------------------------------------------
// jagged views / buffers
// ---------------------------------------------------------------------------------------
template <typename T>
struct jagged_vector_view {
    using size_type = std::size_t;
    using value_type = vector_view<T>;
    size_type m_size = 0;
    value_type* m_ptr = nullptr;       // inner views, accessible where the payload lives
    value_type* m_host_ptr = nullptr;  // host-accessible copy of the inner views (or nullptr)
    jagged_vector_view() = default;
    VECMEM_HOST_AND_DEVICE jagged_vector_view(size_type n, value_type* p, value_type* hp = nullptr)
        : m_size(n), m_ptr(p), m_host_ptr(hp) {}
    template <typename U,
              std::enable_if_t<std::is_same_v<std::remove_cv_t<T>, std::remove_cv_t<U>> &&
                                   !std::is_same_v<T, U>,
                               bool> = true>
    VECMEM_HOST_AND_DEVICE jagged_vector_view(const jagged_vector_view<U>& o)
        : m_size(o.m_size),
          m_ptr(reinterpret_cast<value_type*>(o.m_ptr)),
          m_host_ptr(reinterpret_cast<value_type*>(o.m_host_ptr)) {}
    VECMEM_HOST_AND_DEVICE size_type size() const { return m_size; }
    VECMEM_HOST_AND_DEVICE value_type* ptr() const { return m_ptr; }
    VECMEM_HOST_AND_DEVICE value_type* host_ptr() const { return m_host_ptr; }
};

template <typename T>
struct jagged_vector_data : jagged_vector_view<T> {
    jagged_vector_data() = default;
    explicit jagged_vector_data(std::size_t n)
        : m_rows(std::make_shared<std::vector<vector_view<T>>>(n)) {
        this->m_size = n;
        this->m_ptr = m_rows->data();
        this->m_host_ptr = m_rows->data();
    }
    std::shared_ptr<std::vector<vector_view<T>>> m_rows;
};

template <typename T>
struct jagged_vector_buffer : jagged_vector_view<T> {
    using size_type = std::size_t;
    using value_type = vector_view<T>;
    jagged_vector_buffer() = default;
    template <typename SIZE>
    jagged_vector_buffer(const std::vector<SIZE>& capacities, memory_resource& mr,
                         memory_resource* host_mr = nullptr,
                         buffer_type type = buffer_type::fixed_size) {
        const std::size_t n = capacities.size();
        std::size_t total = 0;
        for (SIZE c : capacities) total += static_cast<std::size_t>(c);
        m_inner = make_unique_alloc<value_type[]>(mr, n);
        m_payload = make_unique_alloc<char[]>(mr, (total ? total : 1) * sizeof(T));
        m_resizable = (type == buffer_type::resizable);
        if (m_resizable) m_sizes = make_unique_alloc<unsigned int[]>(mr, n);
        if (host_mr) {
            m_host_inner = make_unique_alloc<value_type[]>(*host_mr, n);
            this->m_host_ptr = m_host_inner.get();
        } else {
            m_heap_inner.reset(new value_type[n ? n : 1]);
            this->m_host_ptr = m_heap_inner.get();
        }
        T* p = reinterpret_cast<T*>(m_payload.get());
        for (std::size_t i = 0; i < n; ++i) {
            const unsigned int c = static_cast<unsigned int>(capacities[i]);
            this->m_host_ptr[i] = value_type(c, m_resizable ? m_sizes.get() + i : nullptr, p);
            p += c;
        }
        this->m_size = n;
        this->m_ptr = m_inner.get();
    }
    jagged_vector_buffer(jagged_vector_buffer&&) = default;
    jagged_vector_buffer& operator=(jagged_vector_buffer&&) = default;
    bool m_resizable = false;
    unique_alloc_ptr<value_type[]> m_inner, m_host_inner;
    std::unique_ptr<value_type[]> m_heap_inner;
    unique_alloc_ptr<char[]> m_payload;
    unique_alloc_ptr<unsigned int[]> m_sizes;
};
}  // namespace data

// ---------------------------------------------------------------------------------------
// device_vector / jagged_device_vector
// ---------------------------------------------------------------------------------------
template <typename T>
class device_vector {
    public:
    using size_type = unsigned int;
    using value_type = T;
    using reference = T&;
    using const_reference = const T&;
    using pointer = T*;
    using iterator = T*;
    using const_iterator = const T*;
    using size_pointer = typename data::vector_view<T>::size_pointer;
    VECMEM_HOST_AND_DEVICE device_vector(const data::vector_view<T>& v)
        : m_capacity(v.m_capacity), m_size(v.m_size), m_ptr(v.m_ptr) {}
    template <typename U,
              std::enable_if_t<std::is_same_v<std::remove_cv_t<T>, std::remove_cv_t<U>> &&
                                   !std::is_same_v<T, U>,
                               bool> = true>
    VECMEM_HOST_AND_DEVICE device_vector(const data::vector_view<U>& v)
        : m_capacity(v.m_capacity), m_size(v.m_size), m_ptr(v.m_ptr) {}
    VECMEM_HOST_AND_DEVICE size_type size() const { return m_size ? *m_size : m_capacity; }
    VECMEM_HOST_AND_DEVICE size_type capacity() const { return m_capacity; }
    VECMEM_HOST_AND_DEVICE bool empty() const { return size() == 0; }
    VECMEM_HOST_AND_DEVICE T& at(size_type i) const { return m_ptr[i]; }
    VECMEM_HOST_AND_DEVICE T& operator[](size_type i) const { return m_ptr[i]; }
    VECMEM_HOST_AND_DEVICE T& front() const { return m_ptr[0]; }
    VECMEM_HOST_AND_DEVICE T& back() const { return m_ptr[size() - 1]; }
    VECMEM_HOST_AND_DEVICE T* data() const { return m_ptr; }
    VECMEM_HOST_AND_DEVICE T* begin() const { return m_ptr; }
    VECMEM_HOST_AND_DEVICE T* end() const { return m_ptr + size(); }
    VECMEM_HOST_AND_DEVICE const T* cbegin() const { return m_ptr; }
    VECMEM_HOST_AND_DEVICE const T* cend() const { return m_ptr + size(); }
    /// Thread-safe append on a resizable vector; returns the element's index.
    template <typename U = T, std::enable_if_t<!std::is_const_v<U>, bool> = true>
    VECMEM_HOST_AND_DEVICE size_type push_back(const std::remove_const_t<T>& v) const {
        device_atomic_ref<size_type> asize(*m_size);
        const size_type i = asize.fetch_add(1u);
        assert(i < m_capacity);
        m_ptr[i] = v;
        return i;
    }

    private:
    size_type m_capacity;
    size_pointer m_size;
    T* m_ptr;
};

template <typename T>
class jagged_device_vector {
    public:
    using size_type = unsigned int;
    using value_type = device_vector<T>;
    using reference = device_vector<T>;
    using const_reference = device_vector<T>;
    VECMEM_HOST_AND_DEVICE jagged_device_vector(const data::jagged_vector_view<T>& v)
        : m_size(static_cast<size_type>(v.m_size)), m_ptr(v.m_ptr) {}
    VECMEM_HOST_AND_DEVICE size_type size() const { return m_size; }
    VECMEM_HOST_AND_DEVICE bool empty() const { return m_size == 0; }
    VECMEM_HOST_AND_DEVICE device_vector<T> at(size_type i) const { return device_vector<T>(m_ptr[i]); }
    VECMEM_HOST_AND_DEVICE device_vector<T> operator[](size_type i) const {
        return device_vector<T>(m_ptr[i]);
    }

    private:
    size_type m_size;
    data::vector_view<T>* m_ptr;
};

// get_data of host vectors
template <typename T, typename A>
data::vector_view<T> get_data(std::vector<T, A>& v) {
    return {static_cast<unsigned int>(v.size()), v.data()};
}
template <typename T, typename A>
data::vector_view<const T> get_data(const std::vector<T, A>& v) {
    return {static_cast<unsigned int>(v.size()), v.data()};
}
template <typename T, typename A1, typename A2>
data::jagged_vector_data<T> get_data(std::vector<std::vector<T, A1>, A2>& v,
                                     memory_resource* = nullptr) {
    data::jagged_vector_data<T> d(v.size());
    for (std::size_t i = 0; i < v.size(); ++i)
        d.m_ptr[i] = data::vector_view<T>(static_cast<unsigned int>(v[i].size()), v[i].data());
    return d;
}
template <typename T, typename A1, typename A2>
data::jagged_vector_data<const T> get_data(const std::vector<std::vector<T, A1>, A2>& v,
                                           memory_resource* = nullptr) {
    data::jagged_vector_data<const T> d(v.size());
    for (std::size_t i = 0; i < v.size(); ++i)
        d.m_ptr[i] =
            data::vector_view<const T>(static_cast<unsigned int>(v[i].size()), v[i].data());
    return d;
}

// ---------------------------------------------------------------------------------------
// edm::container — SoA container of vector variables
// ---------------------------------------------------------------------------------------
namespace edm {
namespace type {
template <typename T>
struct vector {
    using value = T;
};
template <typename T>
struct scalar {
    using value = T;
};
template <typename T>
struct jagged_vector {
    using value = T;
};
}  // namespace type

namespace details {
template <bool CONST, typename... T>
struct ref_proxy {
    ::cuda::std::tuple<std::conditional_t<CONST, const T&, T&>...> m_refs;
    VECMEM_HOST_AND_DEVICE ref_proxy(std::conditional_t<CONST, const T&, T&>... r) : m_refs(r...) {}
    template <std::size_t I>
    VECMEM_HOST_AND_DEVICE auto& get() const {
        return ::cuda::std::get<I>(m_refs);
    }
};
template <typename... T>
struct value_proxy {
    ::cuda::std::tuple<T...> m_vals;
    value_proxy() = default;
    VECMEM_HOST_AND_DEVICE value_proxy(const T&... v) : m_vals(v...) {}
    template <std::size_t I>
    VECMEM_HOST_AND_DEVICE auto& get() {
        return ::cuda::std::get<I>(m_vals);
    }
    template <std::size_t I>
    VECMEM_HOST_AND_DEVICE const auto& get() const {
        return ::cuda::std::get<I>(m_vals);
    }
};
}  // namespace details

template <bool CONST, typename... T>
struct view_data {
    using size_type = unsigned int;
    using size_pointer = std::conditional_t<CONST, const size_type*, size_type*>;
    size_type m_capacity = 0;
    size_pointer m_size = nullptr;
    ::cuda::std::tuple<vecmem::data::vector_view<std::conditional_t<CONST, const T, T>>...> m_cols;
    VECMEM_HOST_AND_DEVICE size_type capacity() const { return m_capacity; }
    template <std::size_t I>
    VECMEM_HOST_AND_DEVICE auto& get() {
        return ::cuda::std::get<I>(m_cols);
    }
    template <std::size_t I>
    VECMEM_HOST_AND_DEVICE const auto& get() const {
        return ::cuda::std::get<I>(m_cols);
    }
};

template <template <typename> class INTERFACE, typename... VARTYPES>
struct container {
    template <typename B>
    using interface_type = INTERFACE<B>;
    using idx = std::index_sequence_for<VARTYPES...>;

    struct view : view_data<false, typename VARTYPES::value...> {};
    struct const_view : view_data<true, typename VARTYPES::value...> {
        const_view() = default;
        VECMEM_HOST_AND_DEVICE const_view(const view& v) {
            this->m_capacity = v.m_capacity;
            this->m_size = v.m_size;
            copy_cols(v, idx{});
        }

        private:
        template <std::size_t... I>
        VECMEM_HOST_AND_DEVICE void copy_cols(const view& v, std::index_sequence<I...>) {
            ((::cuda::std::get<I>(this->m_cols) = ::cuda::std::get<I>(v.m_cols)), ...);
        }
    };

    /// Owning device buffer: one allocation per column (+ the size word when resizable)
    struct buffer : view {
        using size_type = unsigned int;
        buffer() = default;
        buffer(size_type capacity, memory_resource& mr,
               vecmem::data::buffer_type type = vecmem::data::buffer_type::fixed_size) {
            this->m_capacity = capacity;
            if (type == vecmem::data::buffer_type::resizable) {
                m_size_mem = make_unique_alloc<char[]>(mr, 256);
                this->m_size = reinterpret_cast<size_type*>(m_size_mem.get());
            }
            alloc(capacity, mr, idx{});
        }
        buffer(buffer&&) = default;
        buffer& operator=(buffer&&) = default;
        VECMEM_HOST size_type capacity() const { return this->m_capacity; }

        private:
        template <std::size_t... I>
        void alloc(size_type capacity, memory_resource& mr, std::index_sequence<I...>) {
            ((std::get<I>(m_mem) = make_unique_alloc<char[]>(
                  mr, std::size_t(capacity ? capacity : 1) * sizeof(typename VARTYPES::value)),
              ::cuda::std::get<I>(this->m_cols) = vecmem::data::vector_view<typename VARTYPES::value>(
                  capacity, this->m_size,
                  reinterpret_cast<typename VARTYPES::value*>(std::get<I>(m_mem).get()))),
             ...);
        }
        unique_alloc_ptr<char[]> m_size_mem;
        std::array<unique_alloc_ptr<char[]>, sizeof...(VARTYPES)> m_mem;
    };

    // ---- device containers -------------------------------------------------------------
    template <bool CONST>
    struct device_base {
        using size_type = unsigned int;
        using vw = std::conditional_t<CONST, const_view, view>;
        vw m_view;
        VECMEM_HOST_AND_DEVICE device_base(const vw& v) : m_view(v) {}
        VECMEM_HOST_AND_DEVICE size_type size() const {
            return m_view.m_size ? *m_view.m_size : m_view.m_capacity;
        }
        VECMEM_HOST_AND_DEVICE size_type capacity() const { return m_view.m_capacity; }
        template <std::size_t I>
        VECMEM_HOST_AND_DEVICE auto get() const {
            using T = std::tuple_element_t<I, std::tuple<typename VARTYPES::value...>>;
            return vecmem::device_vector<std::conditional_t<CONST, const T, T>>(
                ::cuda::std::get<I>(m_view.m_cols));
        }
    };
    template <bool CONST>
    struct device_impl : INTERFACE<device_base<CONST>> {
        using base = INTERFACE<device_base<CONST>>;
        using size_type = unsigned int;
        using proxy_type = INTERFACE<details::ref_proxy<CONST, typename VARTYPES::value...>>;
        using const_proxy_type = INTERFACE<details::ref_proxy<true, typename VARTYPES::value...>>;
        using object_type = INTERFACE<details::value_proxy<typename VARTYPES::value...>>;
        VECMEM_HOST_AND_DEVICE device_impl(const typename device_base<CONST>::vw& v) : base(v) {}
        VECMEM_HOST_AND_DEVICE const_proxy_type at(size_type i) const { return cproxy(i, idx{}); }
        VECMEM_HOST_AND_DEVICE const_proxy_type operator[](size_type i) const { return at(i); }
        VECMEM_HOST_AND_DEVICE proxy_type at(size_type i) { return mproxy(i, idx{}); }
        VECMEM_HOST_AND_DEVICE proxy_type operator[](size_type i) { return at(i); }
        /// Thread-safe append on a resizable container; returns the element's index.
        VECMEM_HOST_AND_DEVICE size_type push_back(const object_type& o) {
            device_atomic_ref<size_type> asize(*const_cast<size_type*>(this->m_view.m_size));
            const size_type i = asize.fetch_add(1u);
            assert(i < this->m_view.m_capacity);
            store(i, o, idx{});
            return i;
        }

        private:
        template <std::size_t... I>
        VECMEM_HOST_AND_DEVICE const_proxy_type cproxy(size_type i, std::index_sequence<I...>) const {
            return const_proxy_type(::cuda::std::get<I>(this->m_view.m_cols).m_ptr[i]...);
        }
        template <std::size_t... I>
        VECMEM_HOST_AND_DEVICE proxy_type mproxy(size_type i, std::index_sequence<I...>) {
            return proxy_type(::cuda::std::get<I>(this->m_view.m_cols).m_ptr[i]...);
        }
        template <std::size_t... I>
        VECMEM_HOST_AND_DEVICE void store(size_type i, const object_type& o, std::index_sequence<I...>) {
            ((::cuda::std::get<I>(this->m_view.m_cols).m_ptr[i] = o.template get<I>()), ...);
        }
    };
    using device = device_impl<false>;
    using const_device = device_impl<true>;

    // ---- host container ------------------------------------------------------------------
    struct host_base {
        using size_type = std::size_t;
        std::tuple<vecmem::vector<typename VARTYPES::value>...> m_cols;
        explicit host_base(vecmem::memory_resource& mr)
            : m_cols(vecmem::vector<typename VARTYPES::value>(&mr)...) {}
        size_type size() const { return std::get<0>(m_cols).size(); }
        void resize(size_type n) {
            std::apply([n](auto&... c) { (c.resize(n), ...); }, m_cols);
        }
        void reserve(size_type n) {
            std::apply([n](auto&... c) { (c.reserve(n), ...); }, m_cols);
        }
        template <std::size_t I>
        auto& get() {
            return std::get<I>(m_cols);
        }
        template <std::size_t I>
        const auto& get() const {
            return std::get<I>(m_cols);
        }
    };
    struct host : INTERFACE<host_base> {
        using base = INTERFACE<host_base>;
        using size_type = std::size_t;
        explicit host(vecmem::memory_resource& mr) : base(mr) {}
    };
};
}  // namespace edm

// ---------------------------------------------------------------------------------------
// copy (vecmem/utils/copy.hpp + vecmem::cuda::async_copy): everything is enqueued on one
// CUDA stream; wait() synchronises that stream.
// ---------------------------------------------------------------------------------------
class async_size {
    public:
    async_size(unique_alloc_ptr<unsigned int> p, cudaStream_t s) : m_p(std::move(p)), m_stream(s) {}
    unsigned int get() const {
        SHIM_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        return *m_p;
    }

    private:
    unique_alloc_ptr<unsigned int> m_p;
    cudaStream_t m_stream;
};

class copy {
    public:
    struct event {
        cudaStream_t m_stream;
        void wait() const { SHIM_CUDA_CHECK(cudaStreamSynchronize(m_stream)); }
        void ignore() const {}
        const event* operator->() const { return this; }
    };
    using event_type = event;

    copy() = default;
    explicit copy(cudaStream_t stream) : m_stream(stream) {}
    cudaStream_t stream() const { return m_stream; }

    // ---- setup -----------------------------------------------------------------------------
    template <typename T>
    event setup(data::vector_view<T>& v) const {
        if (v.m_size) SHIM_CUDA_CHECK(cudaMemsetAsync(v.m_size, 0, sizeof(unsigned int), m_stream));
        return {m_stream};
    }
    template <typename T>
    event setup(data::jagged_vector_buffer<T>& b) const {
        if (b.m_size == 0) return {m_stream};
        SHIM_CUDA_CHECK(cudaMemcpyAsync(b.m_ptr, b.m_host_ptr,
                                        b.m_size * sizeof(data::vector_view<T>),
                                        cudaMemcpyHostToDevice, m_stream));
        if (b.m_resizable)
            SHIM_CUDA_CHECK(
                cudaMemsetAsync(b.m_sizes.get(), 0, b.m_size * sizeof(unsigned int), m_stream));
        return {m_stream};
    }
    template <typename B, typename = decltype(std::declval<B&>().m_cols)>
    event setup(B& edm_buffer) const {
        if (edm_buffer.m_size)
            SHIM_CUDA_CHECK(cudaMemsetAsync(edm_buffer.m_size, 0, sizeof(unsigned int), m_stream));
        return {m_stream};
    }

    // ---- memset ----------------------------------------------------------------------------
    template <typename T>
    event memset(const data::vector_view<T>& v, int value) const {
        if (v.m_capacity)
            SHIM_CUDA_CHECK(cudaMemsetAsync(v.m_ptr, value, std::size_t(v.m_capacity) * sizeof(T), m_stream));
        return {m_stream};
    }

    // ---- sizes -----------------------------------------------------------------------------
    template <typename T>
    unsigned int get_size(const data::vector_view<T>& v) const {
        return read_size(v.m_capacity, v.m_size);
    }
    template <typename T>
    async_size get_size(const data::vector_view<T>& v, memory_resource& host_mr) const {
        return read_size_async(v.m_capacity, v.m_size, host_mr);
    }
    template <typename V, typename = decltype(std::declval<const V&>().m_cols)>
    unsigned int get_size(const V& edm_view) const {
        return read_size(edm_view.m_capacity, edm_view.m_size);
    }
    template <typename V, typename = decltype(std::declval<const V&>().m_cols)>
    async_size get_size(const V& edm_view, memory_resource& host_mr) const {
        return read_size_async(edm_view.m_capacity, edm_view.m_size, host_mr);
    }

    // ---- copies ----------------------------------------------------------------------------
    template <typename T1, typename T2, typename A>
    event operator()(const data::vector_view<T1>& from, std::vector<T2, A>& to) const {
        const unsigned int n = get_size(from);
        to.resize(n);
        if (n)
            SHIM_CUDA_CHECK(cudaMemcpyAsync(to.data(), from.m_ptr, std::size_t(n) * sizeof(T2),
                                            cudaMemcpyDefault, m_stream));
        return {m_stream};
    }
    template <typename T1, typename T2>
    event operator()(const data::vector_view<T1>& from, const data::vector_view<T2>& to) const {
        const unsigned int n = from.m_size ? get_size(from) : from.m_capacity;
        if (n)
            SHIM_CUDA_CHECK(cudaMemcpyAsync(const_cast<std::remove_const_t<T2>*>(to.m_ptr), from.m_ptr,
                                            std::size_t(n) * sizeof(T2), cudaMemcpyDefault, m_stream));
        return {m_stream};
    }

    private:
    unsigned int read_size(unsigned int capacity, const unsigned int* p) const {
        if (!p) return capacity;
        unsigned int n = 0;
        SHIM_CUDA_CHECK(cudaMemcpyAsync(&n, p, sizeof(n), cudaMemcpyDefault, m_stream));
        SHIM_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        return n;
    }
    async_size read_size_async(unsigned int capacity, const unsigned int* p,
                               memory_resource& host_mr) const {
        unique_alloc_ptr<unsigned int> h = make_unique_alloc<unsigned int>(host_mr);
        if (!p) {
            *h = capacity;
        } else {
            SHIM_CUDA_CHECK(cudaMemcpyAsync(h.get(), p, sizeof(unsigned int), cudaMemcpyDefault, m_stream));
        }
        return async_size(std::move(h), m_stream);
    }
    cudaStream_t m_stream = nullptr;
};

}  // namespace vecmem
