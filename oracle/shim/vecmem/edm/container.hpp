#pragma once
// Stand-in for vecmem::edm::container (vecmem 1.25.0 is not available offline): a minimal but
// FUNCTIONAL SoA container — host (one pmr vector per variable), (const_)view, (const_)device
// and element proxies — with the member names the reference's code uses
// (size/resize/reserve/push_back/at/operator[]/get<I>). Only vector variables are supported,
// which is all the seeding path's collections have. TEST INFRASTRUCTURE (oracle/_ref build).
#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

#include "vecmem/containers/data/vector_view.hpp"
#include "vecmem/containers/device_vector.hpp"
#include "vecmem/containers/vector.hpp"
#include "vecmem/memory/memory_resource.hpp"

namespace vecmem::edm {
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
// proxy over one element: references into the columns (or owned values for push_back)
template <bool CONST, typename... T>
struct ref_proxy {
    std::tuple<std::conditional_t<CONST, const T&, T&>...> m_refs;
    ref_proxy(std::conditional_t<CONST, const T&, T&>... r) : m_refs(r...) {}
    template <std::size_t I>
    auto& get() const {
        return std::get<I>(m_refs);
    }
};
template <typename... T>
struct value_proxy {
    std::tuple<T...> m_vals;
    value_proxy() = default;
    value_proxy(const T&... v) : m_vals(v...) {}
    template <std::size_t I>
    auto& get() {
        return std::get<I>(m_vals);
    }
    template <std::size_t I>
    const auto& get() const {
        return std::get<I>(m_vals);
    }
};
}  // namespace details

template <typename... T>
struct view_data {
    std::tuple<vecmem::data::vector_view<T>...> m_cols;
    unsigned int m_size = 0;
    unsigned int capacity() const { return m_size; }
    template <std::size_t I>
    auto& get() {
        return std::get<I>(m_cols);
    }
    template <std::size_t I>
    const auto& get() const {
        return std::get<I>(m_cols);
    }
};

template <template <typename> class INTERFACE, typename... VARTYPES>
struct container {
    template <typename B>
    using interface_type = INTERFACE<B>;

    struct view : view_data<typename VARTYPES::value...> {};
    struct const_view : view_data<const typename VARTYPES::value...> {
        const_view() = default;
        const_view(const view& v) {
            copy(v, std::index_sequence_for<VARTYPES...>{});
            this->m_size = v.m_size;
        }
        private:
        template <std::size_t... I>
        void copy(const view& v, std::index_sequence<I...>) {
            ((std::get<I>(this->m_cols) = std::get<I>(v.m_cols)), ...);
        }
    };
    struct buffer : view {};

    // ---- device containers -------------------------------------------------------
    template <bool CONST>
    struct device_base {
        using size_type = unsigned int;
        using vw = std::conditional_t<CONST, const_view, view>;
        std::tuple<vecmem::device_vector<std::conditional_t<CONST, const typename VARTYPES::value,
                                                            typename VARTYPES::value>>...>
            m_cols;
        size_type m_size;
        device_base(const vw& v) : m_cols(make(v, std::index_sequence_for<VARTYPES...>{})), m_size(v.m_size) {}
        size_type size() const { return m_size; }
        size_type capacity() const { return m_size; }
        template <std::size_t I>
        auto& get() {
            return std::get<I>(m_cols);
        }
        template <std::size_t I>
        const auto& get() const {
            return std::get<I>(m_cols);
        }
        private:
        template <std::size_t... I>
        static auto make(const vw& v, std::index_sequence<I...>) {
            return std::make_tuple(
                vecmem::device_vector<std::conditional_t<CONST, const typename VARTYPES::value,
                                                         typename VARTYPES::value>>(std::get<I>(v.m_cols))...);
        }
    };
    template <bool CONST>
    struct device_impl : INTERFACE<device_base<CONST>> {
        using base = INTERFACE<device_base<CONST>>;
        using size_type = unsigned int;
        using proxy_type = INTERFACE<details::ref_proxy<CONST, typename VARTYPES::value...>>;
        using const_proxy_type = INTERFACE<details::ref_proxy<true, typename VARTYPES::value...>>;
        device_impl(const typename device_base<CONST>::vw& v) : base(v) {}
        const_proxy_type at(size_type i) const { return cproxy(i, std::index_sequence_for<VARTYPES...>{}); }
        const_proxy_type operator[](size_type i) const { return at(i); }
        proxy_type at(size_type i) { return mproxy(i, std::index_sequence_for<VARTYPES...>{}); }
        proxy_type operator[](size_type i) { return at(i); }
        private:
        template <std::size_t... I>
        const_proxy_type cproxy(size_type i, std::index_sequence<I...>) const {
            return const_proxy_type(std::get<I>(this->m_cols)[i]...);
        }
        template <std::size_t... I>
        proxy_type mproxy(size_type i, std::index_sequence<I...>) {
            return proxy_type(std::get<I>(this->m_cols)[i]...);
        }
    };
    using device = device_impl<false>;
    using const_device = device_impl<true>;

    // ---- host container -------------------------------------------------------------
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
        using edm_view_type = view;
        using size_type = std::size_t;
        using object_type = INTERFACE<details::value_proxy<typename VARTYPES::value...>>;
        using proxy_type = INTERFACE<details::ref_proxy<false, typename VARTYPES::value...>>;
        using const_proxy_type = INTERFACE<details::ref_proxy<true, typename VARTYPES::value...>>;
        explicit host(vecmem::memory_resource& mr) : base(mr) {}
        void push_back(const object_type& o) { push(o, std::index_sequence_for<VARTYPES...>{}); }
        proxy_type at(size_type i) { return mproxy(i, std::index_sequence_for<VARTYPES...>{}); }
        proxy_type operator[](size_type i) { return at(i); }
        const_proxy_type at(size_type i) const { return cproxy(i, std::index_sequence_for<VARTYPES...>{}); }
        const_proxy_type operator[](size_type i) const { return at(i); }
        private:
        template <std::size_t... I>
        void push(const object_type& o, std::index_sequence<I...>) {
            (std::get<I>(this->m_cols).push_back(o.template get<I>()), ...);
        }
        template <std::size_t... I>
        proxy_type mproxy(size_type i, std::index_sequence<I...>) {
            return proxy_type(std::get<I>(this->m_cols)[i]...);
        }
        template <std::size_t... I>
        const_proxy_type cproxy(size_type i, std::index_sequence<I...>) const {
            return const_proxy_type(std::get<I>(this->m_cols)[i]...);
        }
    };
};

}  // namespace vecmem::edm

namespace vecmem {
// get_data(host container) -> view of its columns
template <typename HOST>
auto get_data(HOST& h) -> decltype(h.m_cols, typename HOST::edm_view_type()) {
    typename HOST::edm_view_type v;
    v.m_size = static_cast<unsigned int>(h.size());
    std::apply([&](auto&... vc) { std::apply([&](auto&... hc) { ((vc = vecmem::get_data(hc)), ...); }, h.m_cols); },
               v.m_cols);
    return v;
}
}  // namespace vecmem
