"""traccc_b200 — B200-native (sm_100a) triplet track seeding + seed parameter estimation,
a drop-in for traccc::cuda::triplet_seeding_algorithm and
traccc::cuda::seed_parameter_estimation_algorithm behind the C-ABI of include/b200seed.h.

`from traccc_b200 import seeding` needs torch + the built libb200seed.so; the config
mirrors in `traccc_b200._lib` only need the library.
"""
from ._lib import (B200SeedError, seedfilter_config, seedfinder_config, spacepoint_grid_config,
                   track_params_estimation_config)

__all__ = ["B200SeedError", "seedfinder_config", "spacepoint_grid_config", "seedfilter_config",
           "track_params_estimation_config"]
