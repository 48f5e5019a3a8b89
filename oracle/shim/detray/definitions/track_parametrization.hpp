#pragma once
// detray/definitions/track_parametrization.hpp (published values): indices of the bound and
// free track-parameter vectors.
namespace detray {
enum bound_indices : unsigned int {
    e_bound_loc0 = 0u,
    e_bound_loc1 = 1u,
    e_bound_phi = 2u,
    e_bound_theta = 3u,
    e_bound_qoverp = 4u,
    e_bound_time = 5u,
    e_bound_size = 6u
};
enum free_indices : unsigned int {
    e_free_pos0 = 0u,
    e_free_pos1 = 1u,
    e_free_pos2 = 2u,
    e_free_time = 3u,
    e_free_dir0 = 4u,
    e_free_dir1 = 5u,
    e_free_dir2 = 6u,
    e_free_qoverp = 7u,
    e_free_size = 8u
};
}  // namespace detray
