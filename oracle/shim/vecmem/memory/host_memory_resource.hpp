#pragma once
#include <cstdlib>
#include "vecmem/memory/memory_resource.hpp"
namespace vecmem {
class host_memory_resource : public memory_resource {
    void* do_allocate(std::size_t n, std::size_t a) override { return ::operator new(n, std::align_val_t(a)); }
    void do_deallocate(void* p, std::size_t, std::size_t a) override { ::operator delete(p, std::align_val_t(a)); }
    bool do_is_equal(const memory_resource& o) const noexcept override { return this == &o; }
};
}
