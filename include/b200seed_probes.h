/* b200seed_probes.h — TEST-ONLY entry points of libb200seed.so.
 *
 * The cut arithmetic of the CUDA kernels (traccc_b200/csrc/seed_math.cuh) is written as
 * __host__ __device__ functions; these probes run the HOST compilation of the same functions so that
 * the CPU test-suite (tests/test_host_math.py, tests/test_abi.py) can compare them with the oracle
 * and with the reference's own headers bit for bit without a GPU. No product path calls them and
 * they are not part of the drop-in interface (include/b200seed.h).
 *
 * `devcfg` is an opaque blob filled by b200seed_host_probe_devcfg (the flattened device copy of
 * the three reference configs, see seed_math.cuh: DevCfg). */
#ifndef B200SEED_PROBES_H
#define B200SEED_PROBES_H

#include "b200seed.h"

#ifdef __cplusplus
extern "C" {
#endif

/* get_axes for a grid config without creating a handle (no GPU needed);
 * B200SEED_EINVAL where the reference throws std::domain_error. */
int b200seed_axes_for(const b200seed_grid_cfg* grid, uint32_t* n_phi, uint32_t* n_z);

/* Fills `out` (>= the returned size) with the device configuration; returns its size in bytes,
 * or B200SEED_EINVAL. */
int b200seed_host_probe_devcfg(const b200seed_finder_cfg* f, const b200seed_grid_cfg* g,
                               const b200seed_filter_cfg* fl, void* out, size_t out_bytes);
/* the device's atan2f (fdlibm algorithm) */
float b200seed_host_probe_atan2f(float y, float x);
/* is_valid_sp + bin index of n spacepoints (0xFFFFFFFF = rejected) */
void b200seed_host_probe_bins(const void* devcfg, uint32_t n, const float* xyz, uint32_t* bins);
/* doublet decision of n (middle, other) pairs {x,y,z,varZ,varR}: 0 none, 1 bottom, 2 top;
 * lin_circle {Zo,cotTheta,iDeltaR,Er,U,V} of the accepted ones */
void b200seed_host_probe_doublets(const void* devcfg, uint32_t n, const float* m, const float* o,
                                  int32_t* kind, float* lc);
/* helix-radius cut of n pairs {x1,y1,x2,y2}: the reference chain and the division-free
 * pre-decision (0 fail, 1 pass, 2 undecided); bounded: the variant without magnitude guards */
void b200seed_host_probe_stage2(const void* devcfg, uint32_t n, const float* xy, int32_t* exact,
                                int32_t* fast);
void b200seed_host_probe_stage2_bounded(const void* devcfg, uint32_t n, const float* xy, int32_t* fast);
/* whether the other spacepoint's pruning cell lies inside the cell window the doublet kernels
 * visit for the middle (must be 1 wherever the exact cut passes) */
void b200seed_host_probe_cell_window(const void* devcfg, const b200seed_finder_cfg* finder,
                                     uint32_t n_sp, uint32_t n, const float* m, const float* o,
                                     const uint32_t* o_bin, int32_t* visited, uint32_t* grid_out);
/* triplet decision of n (middle, lb, lt) combinations; out = {curvature, impact} */
void b200seed_host_probe_triplets(const void* devcfg, uint32_t n, const float* m, const float* lb,
                                  const float* lt, int32_t* ok, int32_t* cut1, float* out);
/* division-free pre-filter of the triplet cuts: rej[i] = 1 => the exact cuts reject as well */
void b200seed_host_probe_triplet_prefilter(const void* devcfg, uint32_t n, const float* m,
                                           const float* lb, const float* lt, int32_t* rej);

#ifdef __cplusplus
}
#endif
#endif /* B200SEED_PROBES_H */
