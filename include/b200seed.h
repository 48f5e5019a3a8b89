/*
 * b200seed.h — C-ABI of the B200-native triplet track seeding + seed
 * parameter estimation library (libb200seed.so).
 *
 * This is the drop-in boundary for the traccc hot path
 *   traccc::cuda::triplet_seeding_algorithm
 *       (reference: device/cuda/include/traccc/cuda/seeding/triplet_seeding_algorithm.hpp:23-111,
 *        orchestration device/common/src/seeding/triplet_seeding_algorithm.cpp:56-262)
 *   traccc::cuda::seed_parameter_estimation_algorithm
 *       (reference: device/cuda/include/traccc/cuda/seeding/seed_parameter_estimation_algorithm.hpp:19-58,
 *        orchestration device/common/src/seeding/seed_parameter_estimation_algorithm.cpp:31-63)
 *
 * Plain C, plain pointers and sizes. All `d_*` pointers are DEVICE pointers to
 * the contiguous column arrays that live inside the reference's vecmem SoA
 * views (edm::spacepoint_collection, edm::seed_collection,
 * edm::measurement_collection). All `h_*` pointers are HOST pointers.
 * Every function returns 0 on success, a negative B200SEED_E* code otherwise;
 * b200seed_last_error() gives the message (the C++ adapter turns it into the
 * exception the reference would throw).
 */
#ifndef B200SEED_H
#define B200SEED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------ */
/* Configuration structs: byte-for-byte restatements of the reference PODs. */
/* ------------------------------------------------------------------------ */

/* traccc::seedfinder_config — core/include/traccc/seeding/detail/seeding_config.hpp:17-139
 * (33 four-byte words, 132 bytes). The six "derived" members are filled by
 * b200seed_finder_cfg_setup(), which restates seedfinder_config::setup() (:123-138). */
typedef struct b200seed_finder_cfg {
    float zMin, zMax, rMax, rMin;
    float collisionRegionMin, collisionRegionMax;
    float phiMin, phiMax;
    float minPt;
    float cotThetaMax;
    float deltaRMin, deltaRMax;
    float deltaZMax;
    float impactMax;
    float sigmaScattering;
    float maxPtScattering;
    uint32_t maxSeedsPerSpM;
    float bFieldInZ;
    float beamPos[2];
    float radLengthPerSeed;
    float zAlign, rAlign;
    float sigmaError;
    /* derived */
    float highland;
    float maxScatteringAngle2;
    float pTPerHelixRadius;
    float minHelixDiameter2;
    float minHelixRadius;
    float pT2perRadius;
    int32_t phiBinDeflectionCoverage;
    uint32_t neighbor_scope[2];
} b200seed_finder_cfg;

/* traccc::spacepoint_grid_config — seeding_config.hpp:142-189 (44 bytes). It is a
 * COPY of eleven finder fields taken at construction (:145-156): later edits of the
 * finder config do not propagate (tests/cpu/test_seeding.cpp:38-44 relies on that). */
typedef struct b200seed_grid_cfg {
    float bFieldInZ;
    float minPt;
    float rMax;
    float zMax;
    float zMin;
    float deltaRMax;
    float cotThetaMax;
    float impactMax;
    float phiMin;
    float phiMax;
    int32_t phiBinDeflectionCoverage;
} b200seed_grid_cfg;

/* traccc::seedfilter_config — seeding_config.hpp:191-219 (56 bytes with padding). */
typedef struct b200seed_filter_cfg {
    float deltaInvHelixDiameter;
    float impactWeightFactor;
    float compatSeedWeight;
    float deltaRMin;
    size_t compatSeedLimit;
    float good_spB_min_radius;
    float good_spB_weight_increase;
    float good_spT_max_radius;
    float good_spT_weight_increase;
    float good_spB_min_weight;
    float seed_min_weight;
    float spB_min_radius;
} b200seed_filter_cfg;

/* traccc::track_params_estimation_config —
 * core/include/traccc/seeding/detail/track_params_estimation_config.hpp:18-33 (56 bytes). */
typedef struct b200seed_tpe_cfg {
    float initial_sigma[6];
    float initial_sigma_qopt;
    float initial_sigma_pt_rel;
    float initial_inflation[6];
} b200seed_tpe_cfg;

/* Fill the structs with the reference's in-class defaults (units: mm, GeV, T =
 * 2.99792458e-4 GeV/(e mm) as in detray::unit<float>). finder_defaults also calls setup. */
void b200seed_finder_cfg_defaults(b200seed_finder_cfg* cfg);
void b200seed_finder_cfg_setup(b200seed_finder_cfg* cfg);
void b200seed_grid_cfg_from_finder(const b200seed_finder_cfg* finder, b200seed_grid_cfg* grid);
void b200seed_filter_cfg_defaults(b200seed_filter_cfg* cfg);
void b200seed_tpe_cfg_defaults(b200seed_tpe_cfg* cfg);

/* ------------------------------------------------------------------------ */
/* Output record of the parameter estimation                                 */
/* ------------------------------------------------------------------------ */

/* One detray::bound_track_parameters<> worth of data (reference:
 * core/include/traccc/edm/track_parameters.hpp:28-49). vec = (loc0, loc1, phi, theta,
 * q/p, time); cov is the 6x6 covariance (only the diagonal is non-zero on this path,
 * so row/column-major is immaterial). 176 bytes. */
typedef struct b200seed_bound_params {
    uint64_t surface_link;
    float vec[6];
    float cov[36];
} b200seed_bound_params;

/* Compact form of b200seed_bound_params. On this path the covariance is diagonal
 * (core/src/seeding/track_params_estimation.cpp:64-86 and estimate_track_params.ipp:60-87 only set
 * the (j, j) elements of a zero-initialised matrix), so 30 of the 36 floats of every record are
 * structural zeros: 56 bytes instead of 176 to move over PCIe. b200seed_expand_params() restores
 * the full records on the host. */
typedef struct b200seed_bound_params_diag {
    uint64_t surface_link;
    float vec[6];
    float cov_diag[6];
} b200seed_bound_params_diag;

/* The part of a parameter record that has to be computed from the seed: phi, theta, q/p and the
 * variance of q/p (which depends on theta and q/p). Everything else in b200seed_bound_params is
 * either copied from the measurement of the bottom spacepoint (surface_link, loc0, loc1), zero
 * (time) or a constant of the b200seed_tpe_cfg (the other five variances):
 * b200seed_expand_seed_params() rebuilds the records on the host, bit for bit. 16 bytes per seed
 * instead of 56 / 176, and the measurement columns need not go to the device at all. With
 * B200SEED_PCIE_PARAMS=compact in the environment of b200seed_create / b200seed_pool_create the
 * host-buffer entry points (b200seed_run_host, b200seed_pool_process) move the parameters over PCIe
 * in this form and complete the records on the host (for hosts whose D->H rate is the limit; the
 * host then pays ~1 ms of gathers per 10k-particle event, which is why it is not the default). */
typedef struct b200seed_seed_params {
    float phi, theta, qop, var_qop;
} b200seed_seed_params;

/* A parameter record without the five variances that are constants of the b200seed_tpe_cfg and
 * without the time (zero): 32 bytes instead of 56 / 176. b200seed_expand_packed_params() restores
 * the records on the host — a sequential copy, no look-ups. With B200SEED_PCIE_PARAMS=packed the
 * host-buffer entry points move the parameters over PCIe in this form (measured slower than the
 * records themselves on the hosts tried: not the default). */
typedef struct b200seed_bound_params_packed {
    uint64_t surface_link;
    float loc0, loc1, phi, theta, qop, var_qop;
} b200seed_bound_params_packed;

/* Device-side counters of one event. Written by b200seed_run when d_counters != NULL;
 * they replace the reference's D->H size reads (triplet_seeding_algorithm.cpp:64-224)
 * for logging and for the parity tests. */
typedef struct b200seed_counters {
    uint32_t n_spacepoints;      /* input size */
    uint32_t n_valid;            /* spacepoints passing is_valid_sp */
    uint32_t n_active_middles;   /* middles with >=1 bottom and >=1 top doublet */
    uint32_t n_mid_bot;          /* mid-bottom doublets of active middles */
    uint32_t n_mid_top;          /* mid-top doublets of active middles */
    uint32_t n_triplets;         /* triplets passing triplet_finding_helper::isCompatible */
    uint32_t n_seeds;            /* seeds written (== *d_n_seeds) */
    uint32_t overflow;           /* bit mask of B200SEED_OVF_* ; 0 == results complete */
    uint64_t pair_tests;         /* sum over valid middles of the spacepoints in their neighbour
                                    bins == candidate pairs the reference tests */
    uint64_t triplet_tests;      /* sum over active middles of nMidBot * nMidTop */
    uint64_t pair_visited;       /* candidate pairs this library actually evaluated (after the
                                    conservative (r, z) cell pruning); a performance counter: it
                                    depends on how middles are grouped, not only on the event */
    uint32_t n_fallback_middles; /* middles the group kernel handed to the warp-per-middle kernel */
    uint32_t reserved_;
    uint64_t triplet_visited;    /* (mid-bottom, mid-top) pairs inside the conservative cotTheta
                                    windows == combinations the triplet kernel actually evaluated
                                    (triplet_tests is what the reference evaluates) */
} b200seed_counters;

#define B200SEED_OVF_DOUBLETS 1u /* doublet arena too small: raise max_doublets */
#define B200SEED_OVF_SEEDS 2u    /* seed_capacity too small                     */
#define B200SEED_OVF_DUMP 4u     /* debug triplet dump buffer too small         */
#define B200SEED_OVF_TRIPLETS 8u /* one mid-bottom doublet has more triplets than the
                                    shared-memory list holds AND the doublet arena has no
                                    room left for them: raise max_doublets        */

/* Error codes */
#define B200SEED_OK 0
#define B200SEED_EINVAL -1   /* bad argument / unsupported configuration (std::domain_error upstream) */
#define B200SEED_ECUDA -2    /* CUDA runtime error (TRACCC_CUDA_ERROR_CHECK upstream) */
#define B200SEED_ENOMEM -3   /* workspace too small */
#define B200SEED_EOVERFLOW -4 /* a capacity-bounded buffer was too small: the event's results are
                                 truncated (counters.overflow says which); the reference never
                                 truncates, so callers must treat this as a failed event */

typedef struct b200seed_handle b200seed_handle;

/* ------------------------------------------------------------------------ */
/* Life cycle                                                                */
/* ------------------------------------------------------------------------ */

/* Replaces the constructors of cuda::triplet_seeding_algorithm
 * (cuda/seeding/triplet_seeding_algorithm.hpp:35-40) and
 * cuda::seed_parameter_estimation_algorithm (…/seed_parameter_estimation_algorithm.hpp:33-37).
 * Computes the phi/z axes like get_axes (spacepoint_binning_helper.hpp:22-110); a
 * configuration get_axes would reject with std::domain_error gives B200SEED_EINVAL.
 * `tpe` may be NULL (defaults). One handle per host thread / stream, like the reference. */
int b200seed_create(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                    const b200seed_filter_cfg* filter, const b200seed_tpe_cfg* tpe,
                    int device, b200seed_handle** out);
void b200seed_destroy(b200seed_handle* h);
/* Message of the last failure on this handle (or of the last failed create when h==NULL). */
const char* b200seed_last_error(const b200seed_handle* h);

/* Axes chosen at construction: phi is circular, z is regular/closed. */
int b200seed_get_axes(const b200seed_handle* h, uint32_t* n_phi, float* phi_min, float* phi_max,
                      uint32_t* n_z, float* z_min, float* z_max);

/* Doublet arena capacity (entries per direction). 0 selects the default policy
 * max(2^20, 5e-3 * max_spacepoints^2). */
int b200seed_set_max_doublets(b200seed_handle* h, uint64_t max_doublets);

/* Tuning knob: mid-bottom doublets of one middle staged in shared memory by the doublet
 * kernel (mid-tops: half of it) before its list is allocated in the arena; longer lists
 * take a second scan that writes straight to the arena. 0 = automatic (384 up to 55k
 * spacepoints, 512 above). Values whose shared-memory footprint (80 bytes * cap per CTA) exceeds
 * the device's opt-in limit are rejected with B200SEED_EINVAL. Results do not depend on it. */
int b200seed_set_stage_cap(b200seed_handle* h, uint32_t cap);

/* Tuning knob: accepted triplets of one middle kept in shared memory before they are merged into
 * its top-N (0 = automatic: 96 up to 80k spacepoints, 128 above). A single mid-bottom doublet with
 * more accepted triplets than this takes a slow path through global memory (the unused tail of
 * the doublet arena). Results do not depend on it. */
int b200seed_set_triplet_list_cap(b200seed_handle* h, uint32_t cap);

/* Truncation check for the asynchronous entry points (b200seed_run, b200seed_run_n_on_device):
 * after the caller has synchronised the stream, returns B200SEED_EOVERFLOW (and the B200SEED_OVF_*
 * mask in *mask_out, may be NULL) if any event run on this handle since the last call overflowed a
 * capacity-bounded buffer, B200SEED_OK otherwise; the record is cleared. Works whether or not
 * d_counters was passed. The host-buffer entry points (b200seed_run_host, b200seed_pool_process)
 * return B200SEED_EOVERFLOW themselves. */
int b200seed_check_overflow(b200seed_handle* h, uint32_t* mask_out);

/* Bytes of device scratch b200seed_run needs for events of up to max_spacepoints. */
size_t b200seed_workspace_bytes(const b200seed_handle* h, uint32_t max_spacepoints);

/* ------------------------------------------------------------------------ */
/* The hot path                                                              */
/* ------------------------------------------------------------------------ */

/* Replaces device::triplet_seeding_algorithm::operator()
 * (device/common/src/seeding/triplet_seeding_algorithm.cpp:56-262).
 * In : spacepoint columns `global` (std::array<float,3>, stride 3 floats), z_variance,
 *      radius_variance (edm/spacepoint_collection.hpp:223-234).
 * Out: seed columns bottom/middle/top index + quality (edm/seed_collection.hpp:142-146),
 *      the resizable buffer's size word *d_n_seeds, in the reference CPU's seed order.
 * Everything is enqueued on `stream` (a cudaStream_t); no host synchronisation happens,
 * like the reference "returns a buffer which is not necessarily filled yet".
 * n_sp == 0 writes *d_n_seeds = 0 and returns. */
int b200seed_run(b200seed_handle* h, void* stream, uint32_t n_sp, const float* d_xyz,
                 const float* d_var_z, const float* d_var_r, void* d_workspace,
                 size_t workspace_bytes, uint32_t seed_capacity, uint32_t* d_bottom,
                 uint32_t* d_middle, uint32_t* d_top, float* d_quality, uint32_t* d_n_seeds,
                 b200seed_counters* d_counters);

/* Replaces device::seed_parameter_estimation_algorithm::operator()
 * (device/common/src/seeding/seed_parameter_estimation_algorithm.cpp:31-63) with a
 * homogeneous field `bfield` (3 floats, host memory, copied by value).
 * d_sp_meas_index_1 : spacepoint column measurement_index_1
 * d_meas_local      : measurement column local_position (stride 2 floats)
 * d_meas_surface    : measurement column surface_link (64-bit identifier)
 * The number of seeds is read on the device from *d_n_seeds (<= seed_capacity). */
int b200seed_estimate_params(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                             uint32_t seed_capacity, const uint32_t* d_bottom,
                             const uint32_t* d_middle, const uint32_t* d_top,
                             const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                             const float* d_meas_local, const uint64_t* d_meas_surface,
                             const float bfield[3], b200seed_bound_params* d_params);

/* An inhomogeneous magnetic field sampled on a regular grid: the reference's
 * cuda::inhom_global_bfield_backend_t = covfie affine<linear<clamp<strided<array<float3>>>>>
 * (device/cuda/src/utils/magnetic_field_types.hpp:27-32). */
typedef struct b200seed_field_grid {
    float affine[12];   /* row-major 3x4: grid coordinate = A * (x, y, z, 1)              */
    uint32_t size[3];   /* grid points per axis                                            */
    const float* data;  /* DEVICE: size[0]*size[1]*size[2] float3 field vectors, row-major:
                           point (i, j, k) at ((i * size[1] + j) * size[2] + k) * 3        */
} b200seed_field_grid;

/* b200seed_estimate_params with the field looked up at every seed's bottom spacepoint
 * (device/common/.../impl/estimate_track_params.ipp:45-50): trilinear interpolation, indices
 * clamped to the grid. `field` is a host struct whose data pointer is device memory. */
int b200seed_estimate_params_inhom(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                                   uint32_t seed_capacity, const uint32_t* d_bottom,
                                   const uint32_t* d_middle, const uint32_t* d_top,
                                   const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                                   const float* d_meas_local, const uint64_t* d_meas_surface,
                                   const b200seed_field_grid* field,
                                   b200seed_bound_params* d_params);

/* b200seed_estimate_params with the 56-byte diagonal output records (homogeneous field). */
int b200seed_estimate_params_diag(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                                  uint32_t seed_capacity, const uint32_t* d_bottom,
                                  const uint32_t* d_middle, const uint32_t* d_top,
                                  const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                                  const float* d_meas_local, const uint64_t* d_meas_surface,
                                  const float bfield[3], b200seed_bound_params_diag* d_params);
/* b200seed_estimate_params with the 16-byte b200seed_seed_params output (homogeneous field); the
 * measurement columns are not needed on the device. */
int b200seed_estimate_params_compact(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                                     uint32_t seed_capacity, const uint32_t* d_bottom,
                                     const uint32_t* d_middle, const uint32_t* d_top,
                                     const float* d_xyz, const float bfield[3],
                                     b200seed_seed_params* d_params);
/* HOST helper: the records of n seeds from their b200seed_seed_params, the seeds' bottom
 * spacepoints and the HOST copies of the columns b200seed_estimate_params reads on the device
 * (sp_meas_index_1 / meas_local / meas_surface: NULL means what it means there). out_full and / or
 * out_diag (either may be NULL) receive exactly what b200seed_estimate_params /
 * b200seed_estimate_params_diag would have written. */
void b200seed_expand_seed_params(const b200seed_handle* h, uint32_t n, const uint32_t* bottom,
                                 const b200seed_seed_params* in, const uint32_t* sp_meas_index_1,
                                 const float* meas_local, const uint64_t* meas_surface,
                                 b200seed_bound_params* out_full, b200seed_bound_params_diag* out_diag);
/* b200seed_estimate_params with the 32-byte packed output records (homogeneous field). */
int b200seed_estimate_params_packed(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                                    uint32_t seed_capacity, const uint32_t* d_bottom,
                                    const uint32_t* d_middle, const uint32_t* d_top,
                                    const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                                    const float* d_meas_local, const uint64_t* d_meas_surface,
                                    const float bfield[3], b200seed_bound_params_packed* d_params);
/* HOST helper: n packed records -> full and / or diagonal records (either may be NULL), exactly
 * what b200seed_estimate_params / _diag would have written. */
void b200seed_expand_packed_params(const b200seed_handle* h, uint32_t n,
                                   const b200seed_bound_params_packed* in,
                                   b200seed_bound_params* out_full, b200seed_bound_params_diag* out_diag);
/* HOST helper: n diagonal records -> full 176-byte records (off-diagonal elements zero). */
void b200seed_expand_params(const b200seed_bound_params_diag* in, uint32_t n,
                            b200seed_bound_params* out);

/* ------------------------------------------------------------------------ */
/* The step before the path: spacepoint formation       */
/* ------------------------------------------------------------------------ */

/* A placed planar surface: what detray::tracking_surface::local_to_global uses for a 2D
 * measurement (core/include/traccc/seeding/impl/spacepoint_formation.ipp:34-40) — the
 * translation and the columns of the rotation of the surface's transform3. The caller
 * flattens the detector into a table of these, indexed by the surface index that the
 * measurement's surface_link (detray::geometry::identifier) carries. */
typedef struct b200seed_surface {
    float translation[3];
    float x_axis[3];
    float y_axis[3];
    float z_axis[3];
} b200seed_surface;

/* Replaces device::silicon_pixel_spacepoint_formation_algorithm::operator()
 * (device/common/src/seeding/silicon_pixel_spacepoint_formation_algorithm.cpp:20-52, kernel
 * device/common/include/traccc/seeding/device/impl/form_spacepoints.ipp:19-51).
 * In : measurement columns local_position (stride 2 floats), dimensions (NULL = all 2D) and,
 *      per measurement, the index of its surface in d_surfaces.
 * Out: one spacepoint per 2D measurement — global (stride 3 floats), zero variances,
 *      measurement_index_1 = measurement index, measurement_index_2 = 0xFFFFFFFF — and the
 *      resizable buffer's size word *d_n_sp. The output columns need room for n_meas entries;
 *      d_var_z / d_var_r / d_meas_index_1 / d_meas_index_2 may be NULL.
 * Order: measurement order, i.e. the order of the reference's HOST algorithm
 * (core/src/seeding/silicon_pixel_spacepoint_formation.hpp:47-59); the reference's device kernel
 * appends atomically in arbitrary order. Measurements whose surface index is >= n_surfaces are
 * skipped. Asynchronous on `stream`; n_meas == 0 writes *d_n_sp = 0. */
int b200seed_form_spacepoints(b200seed_handle* h, void* stream, uint32_t n_meas,
                              const float* d_meas_local, const uint32_t* d_meas_dim,
                              const uint32_t* d_meas_surface_index,
                              const b200seed_surface* d_surfaces, uint32_t n_surfaces, float* d_xyz,
                              float* d_var_z, float* d_var_r, uint32_t* d_meas_index_1,
                              uint32_t* d_meas_index_2, uint32_t* d_n_sp);

/* b200seed_run for spacepoints whose number only exists on the device (the size word written
 * by b200seed_form_spacepoints): max_sp is an upper bound (the capacity of the columns, the
 * value the workspace is sized for), the kernels read min(*d_n_sp, max_sp). Chains formation
 * and seeding on one stream without the D->H size read of the reference
 * (triplet_seeding_algorithm.cpp:64-73). */
int b200seed_run_n_on_device(b200seed_handle* h, void* stream, uint32_t max_sp,
                             const uint32_t* d_n_sp, const float* d_xyz, const float* d_var_z,
                             const float* d_var_r, void* d_workspace, size_t workspace_bytes,
                             uint32_t seed_capacity, uint32_t* d_bottom, uint32_t* d_middle,
                             uint32_t* d_top, float* d_quality, uint32_t* d_n_seeds,
                             b200seed_counters* d_counters);

/* End-to-end convenience with HOST buffers: H->D of the event, seeding, parameter
 * estimation, D->H of seeds + parameters, stream synchronised before returning. This is
 * what seeding_example_cuda.cpp:264-356 does around the two algorithms. Device staging
 * buffers are owned by the handle and grow on demand. h_params may be NULL. */
int b200seed_run_host(b200seed_handle* h, void* stream, uint32_t n_sp, const float* h_xyz,
                      const float* h_var_z, const float* h_var_r,
                      const uint32_t* h_sp_meas_index_1, uint32_t n_meas,
                      const float* h_meas_local, const uint64_t* h_meas_surface,
                      const float bfield[3], uint32_t seed_capacity, uint32_t* h_bottom,
                      uint32_t* h_middle, uint32_t* h_top, float* h_quality,
                      b200seed_bound_params* h_params, uint32_t* h_n_seeds,
                      b200seed_counters* h_counters);

/* ------------------------------------------------------------------------ */
/* Throughput: many events through one device                               */
/* ------------------------------------------------------------------------ */

/* One event of b200seed_pool_process: the arguments of b200seed_run_host as a record.
 * All pointers are HOST pointers (pinned memory for asynchronous copies); n_seeds,
 * counters and status are written by the pool. */
typedef struct b200seed_event_io {
    uint32_t n_spacepoints;
    uint32_t n_measurements;
    const float* xyz;
    const float* var_z;
    const float* var_r;
    const uint32_t* sp_meas_index_1;
    const float* meas_local;
    const uint64_t* meas_surface;
    float bfield[3];
    uint32_t seed_capacity;
    uint32_t* bottom;
    uint32_t* middle;
    uint32_t* top;
    float* quality;
    b200seed_bound_params* params; /* may be NULL: seeding only */
    uint32_t n_seeds;
    int32_t status;
    b200seed_counters counters;
    /* the parameters as 56-byte diagonal records (may be NULL). When set, the parameters cross
     * PCIe in this form (a third of the bytes); `params` may be set as well (both are filled). */
    b200seed_bound_params_diag* params_diag;
    /* the parameters as 32-byte packed records (may be NULL): delivered as they come off the
     * device, no host work; b200seed_expand_packed_params() restores the other forms on demand.
     * When set, the parameters cross PCIe in this form; `params` / `params_diag` are then expanded
     * from it on the host if they are set as well. */
    b200seed_bound_params_packed* params_packed;
} b200seed_event_io;

typedef struct b200seed_pool b200seed_pool;

/* The host side of a throughput job on one device, replacing what the reference's
 * multi-threaded throughput application does around the two algorithms
 * (examples/run/common/include/traccc/examples/impl/throughput_mt.ipp:170-298: one
 * algorithm instance + stream per host thread, events handed out dynamically).
 * n_workers host threads are created; each owns two algorithm instances / CUDA streams, so
 * it always has one event in flight on the device while it collects the previous one.
 * For several GPUs create one pool per device (events are independent: no exchange). */
int b200seed_pool_create(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                         const b200seed_filter_cfg* filter, const b200seed_tpe_cfg* tpe,
                         int device, int n_workers, b200seed_pool** out);
/* Process events[0..n_events): H->D, seeding, parameter estimation, D->H for each; returns
 * when all are done (0, or the first failing event's status; see events[i].status). */
int b200seed_pool_process(b200seed_pool* pool, b200seed_event_io* events, uint32_t n_events);
const char* b200seed_pool_last_error(const b200seed_pool* pool);
void b200seed_pool_destroy(b200seed_pool* pool);

/* ------------------------------------------------------------------------ */
/* Introspection for the parity tests and the bench                          */
/* ------------------------------------------------------------------------ */

/* Byte offsets of the intermediate arrays inside the workspace for a given
 * max_spacepoints (valid after b200seed_run on that workspace):
 *   bin_offsets : uint32[n_bins + 1]   start of each (phi + n_phi * z) bin in sorted order
 *   sorted_index: uint32[n_valid]      original spacepoint index per sorted position
 *                                      (ascending inside each bin == CPU grid order)
 *   sp_xyzr     : float4[n_valid]      {x, y, z, radius} per sorted position
 *   mid_counts  : uint32[2][max_sp]    nMidBot / nMidTop per sorted position (0 if inactive)
 *   mid_offsets : uint32[2][max_sp]    start of each middle's list in the doublet arena
 *                                      (bump-allocated: the placement of the lists relative
 *                                      to each other is arbitrary)
 *   doublets    : 32-byte records [2][max_doublets]: {cotTheta, iDeltaR, Er, U, V, Zo,
 *                 radius of the other spacepoint, sorted position of the other spacepoint};
 *                 mid-bottom lists are stored in order of discovery; mid-top lists are sorted
 *                 by cotTheta and carry, in the Zo slot (u32 bits), the key that orders the
 *                 partners of a middle like the reference's loops do:
 *                 (position of the partner's phi bin in the neighbour walk) * n_valid +
 *                 sorted position
 *   triplet_dump: 32-byte records {sorted pos bottom, middle, top (u32), index of the
 *                 mid-bottom doublet in its stored list, key of the mid-top doublet (u32),
 *                 curvature, weight after the compatible-seed bonus, z_vertex (f32)} in no
 *                 particular order; filled only when dumping is enabled. */
typedef struct b200seed_ws_layout {
    size_t bin_offsets;
    size_t sorted_index;
    size_t sp_xyzr;
    size_t mid_counts;
    size_t mid_offsets;
    size_t doublets;
    size_t triplet_dump;
    size_t triplet_dump_count; /* uint32 */
    uint64_t max_doublets;
    uint64_t max_triplet_dump;
    uint32_t n_bins;
    uint32_t max_spacepoints;
} b200seed_ws_layout;
int b200seed_workspace_layout(const b200seed_handle* h, uint32_t max_spacepoints,
                              b200seed_ws_layout* out);

/* Enable (capacity > 0) or disable (0) the debug dump of every triplet. Changes
 * b200seed_workspace_bytes. Off by default: the production path never materialises
 * triplets in HBM. */
int b200seed_set_triplet_dump(b200seed_handle* h, uint64_t max_triplets);

/* Per-kernel device timing. When enabled, b200seed_run brackets every kernel with CUDA
 * events on `stream`; b200seed_get_timings synchronises those events and returns the
 * milliseconds of the last run. names/ms hold up to `cap` entries; returns the count. */
int b200seed_set_timing(b200seed_handle* h, int enabled);
int b200seed_get_timings(b200seed_handle* h, const char** names, float* ms, int cap);
/* Number of kernels one b200seed_run (+ estimate_params) launches. */
int b200seed_launches_per_event(const b200seed_handle* h, int with_params);

/* Measured non-fused FP32 rate of `device` in ops/s (FMUL/FADD issue rate; the library is
 * built without FMA contraction): the FP32 roofline denominator of bench.py. */
int b200seed_measure_fp32_peak(int device, double* ops_per_s);

const char* b200seed_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SEED_H */
