#pragma once
// Stand-in for vecmem::edm::container: only the *names* the reference's collection aliases
// mention. The probe instantiates the reference's interface templates (edm::spacepoint<BASE>)
// over its own BASE, so no container machinery is needed.
#include <cstddef>
namespace vecmem::edm {
namespace type {
template <typename T>
struct vector {};
template <typename T>
struct scalar {};
template <typename T>
struct jagged_vector {};
}  // namespace type
template <template <typename> class INTERFACE, typename... VARTYPES>
struct container {
    struct host;
    struct device;
    struct const_device;
    struct view;
    struct const_view;
    struct buffer;
};
}  // namespace vecmem::edm
