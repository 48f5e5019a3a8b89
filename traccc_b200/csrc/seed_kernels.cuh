// seed_kernels.cuh — hand-written sm_100a kernels of the triplet seeding path.
//
// Pipeline of one event (all on one stream, no host synchronisation, every intermediate
// addressed through the workspace):
//
//   k_form_spacepoints  (optional, the step before the path) 2D measurements -> spacepoints in
//                    measurement order; single-pass compaction with decoupled look-back
//   k_bin_count      is_valid_sp + bin index per spacepoint, per-block bin histogram, bin totals,
//                    population of the fine (r, z) cells inside every bin
//   k_cell_scan      CTA per bin: start of the bin (sum of the totals before it), scan of its row
//                    of per-block counts, start of every (r row, z cell) inside the bin; cost
//                    class of every (bin, r row) for the ticket order of k_doublets
//   k_bin_scatter    stable scatter into bin-sorted float4 {x,y,z,r} / float2 {varZ,varR}
//                    (the reference's grid order) + the cell-sorted copy used for pruning + the
//                    ticket order (mid_order: longest middles first)
//   k_doublets<0>    warp per middle (ticket queue): cell windows of the neighbour bins ->
//                    flattened candidate list -> doublet cuts at full lane occupancy (helix cut
//                    decided by a division-free polynomial, exact chain only near its boundary),
//                    ballot/popc compaction, lin_circle, arena write; work lists for the next
//                    kernels (active middles in eight classes, longest first; middles whose lists
//                    outgrow shared memory)
//   k_doublets<3>    the middles whose row populations leave one side (almost) empty: 32 per warp
//                    pre-screened for a partner on that side, one per LANE; survivors scanned
//                    warp-wide, that side first. Programmatic dependent of k_doublets<0>: no
//                    shared data, fills its tail
//   k_doublets<2>    survivors beyond four per batch (fallback list; normally empty)
//   k_doublets<1>    the middles whose lists outgrew the staging area: records written straight
//                    to the arena, bucket sort (normally empty)
//   k_triplets<DENSE> warp per active middle, longest jobs first: lane-owns-mid-bottom windows in
//                    the cotTheta-sorted mid-tops -> flattened pairs -> exact cuts, compatible-seed
//                    bonus, per-middle top-N in shared memory (DENSE: candidate compaction, warp
//                    selection and a division-free pre-filter for the busiest events)
//   k_seed_gather    seed offsets (single-pass look-back scan) + seeds in the reference CPU's
//                    order, original spacepoint indices, counters
//   k_estimate_params one thread per seed, records staged per warp and written coalesced
//
// Replaces device/cuda/src/seeding/triplet_seeding_algorithm.cu:30-160 (9 kernels, 7
// blocking D->H reads between them), seed_parameter_estimation_algorithm.cu:22-33 and
// silicon_pixel_spacepoint_formation_algorithm.cu.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/b200seed.h"
#include "seed_math.cuh"

namespace b200seed {

constexpr uint32_t INVALID_BIN = 0xFFFFFFFFu;
constexpr int BIN_THREADS = 256;    // spacepoints per binning block
#ifndef B200_WARPS_PER_CTA
#define B200_WARPS_PER_CTA 8
#endif
constexpr int WARPS_PER_CTA = B200_WARPS_PER_CTA;
#ifdef B200_TAIL_PROBE
// Debug build only (tools/tail_probe.py, tools/middle_cost.py): time (globaltimer, ns) at which
// every warp of the two search kernels drew its first ticket and ran out of tickets, for
// k_triplets the start and the sizes of its last middle, and the SM cycles k_doublets spent on
// every middle; read back by b200seed_debug_tail_probe / b200seed_debug_middle_cycles.
__device__ unsigned long long g_tail_probe[2][4][16384];
__device__ uint32_t g_middle_cycles[1 << 19];
__device__ __forceinline__ unsigned long long probe_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif
constexpr int MAX_TOPK = 16;        // upper bound on maxSeedsPerSpM
constexpr int MAX_COMPAT = 8;       // upper bound on compatSeedLimit
// k_triplets hands its middles out longest jobs first: "heavy" = nMidBot * nMidTop at or above
// this (the median of an ordinary event is ~1.3k, the mean ~5k, the maximum ~30k). Measured on
// the 10k-particle event: 8192 -> 131 us, 4096 -> 123, 2048 -> 119, 1024 -> 120, four classes
// 115 us but more atomic traffic in k_doublets and no better throughput.
#ifndef B200_TRIPLET_HEAVY_WORK
#define B200_TRIPLET_HEAVY_WORK 2048ull
#endif
constexpr unsigned long long TRIPLET_HEAVY_WORK = B200_TRIPLET_HEAVY_WORK;
// Work split between k_triplets (heavy middles, a warp each) and k_triplets_pool (light middles,
// several per warp; seed_pool.cuh): decided by the doublet kernels when they append a middle to
// the work list (heavy from the front, light from the back).
constexpr uint32_t POOL_NT = 64;    // a light middle has at most this many mid-tops
constexpr uint32_t POOL_NB = 2047;  // ... and mid-bottoms
__host__ __device__ inline bool pool_is_light(uint32_t nB, uint32_t nT) {
    return nT <= POOL_NT && nB <= POOL_NB && (unsigned long long)nB * nT < TRIPLET_HEAVY_WORK;
}

// Work list of the triplet kernels: WORK_CLASSES lists of n_sp entries each, filled by the doublet
// kernels and drawn class by class — longest jobs first, so that a launch does not end on a few
// warps that drew a long middle last, and its final tickets are the shortest jobs (the time a warp
// spends on a middle follows its number of mid-bottom rows). Classes 0-3: the heavy middles
// (see pool_is_light), 4-7: the light ones.
constexpr uint32_t WORK_CLASSES = 8, WORK_HEAVY_CLASSES = 4;
constexpr uint32_t LANES_FIRST_CLASS = 5;  // light middles with fewer than 32 mid-bottom rows
__host__ __device__ inline uint32_t work_class(uint32_t nB, uint32_t nT) {
    if (!pool_is_light(nB, nT)) return (nB < 192u ? 1u : 0u) + (nB < 96u ? 1u : 0u) + (nB < 48u ? 1u : 0u);
    return 4u + (nB < 32u ? 1u : 0u) + (nB < 16u ? 1u : 0u) + (nB < 8u ? 1u : 0u);
}

// Flag in the entries of mid_order for the classes with a scarce side: it is the lower one.
constexpr uint32_t MID_LOWER_SCARCE = 0x80000000u;
#ifndef B200_SIDED_INLINE
#define B200_SIDED_INLINE 4u
#endif
constexpr uint32_t SIDED_INLINE = B200_SIDED_INLINE;  // survivors of a pre-screened batch scanned by the same warp

// Small control block at the start of the workspace, zeroed at the start of each event.
struct Control {
    uint32_t cursor[2];       // doublet arena bump pointers (bottom / top)
    uint32_t ticket;          // k_triplets work queue
    uint32_t n_active;
    uint32_t n_mid_bot;
    uint32_t n_mid_top;
    uint32_t n_triplets;
    uint32_t overflow;
    unsigned long long pair_tests;
    unsigned long long triplet_tests;
    uint32_t dump_cursor;
    uint32_t n_valid;         // written by the first scan
    uint32_t n_seeds_total;   // written by the second scan
    uint32_t ticket_d;        // k_doublets work queue
    uint32_t n_spill;         // middles handed to k_doublets<true> (lists longer than the staging area)
    uint32_t ticket_s;        // its work queue
    uint32_t pad0_[2];
    uint32_t has_variance;    // set by k_bin_scatter if any z / radius variance is non-zero
    uint32_t pad_;
    unsigned long long pair_visited;  // candidates actually loaded by k_doublets
    // k_doublets_tile (seed_tile.cuh): groups of neighbouring middles
    uint32_t n_group_big;     // groups listed from the front of group_list (many middles: drawn first)
    uint32_t n_group_small;   // groups listed from the back
    uint32_t ticket_g;        // its work queue
    uint32_t n_fallback;      // middles handed back to the warp-per-middle kernel
    uint32_t ticket_f;        // work queue of that pass
    uint32_t ticket_p;        // work queue of k_triplets_pool (light middles)
    unsigned long long triplet_visited;  // (mid-bottom, mid-top) pairs inside the cotTheta windows
    uint32_t n_slow;          // middles handed to the slow path (a row outgrew the triplet list)
    uint32_t ticket_q;        // its work queue
    uint32_t slow_done;       // set by k_seed_gather's tile 0 when they are finished
    uint32_t pad3_;
    uint32_t n_cls[WORK_CLASSES];  // entries in each class of the work list
    uint32_t n_dcls[WORK_CLASSES];  // ... of k_doublets' own ticket order (mid_order)
    uint32_t ticket_e;        // work queue of k_doublets<3> (middles with a scarce side)
    uint32_t n_main;          // tickets of k_doublets<0>: n_valid minus those of k_doublets<3>
    uint32_t pad4_[2];
};

// Start of every class in ticket order (s_pre[WORK_CLASSES] = all): once per CTA, before a
// __syncthreads().
__device__ __forceinline__ void work_prefix(const uint32_t* n_cls, uint32_t* s_pre) {
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t c = 0; c < WORK_CLASSES; ++c) {
            s_pre[c] = acc;
            acc += n_cls[c];
        }
        s_pre[WORK_CLASSES] = acc;
    }
}
// The middle behind ticket t (t < s_pre[WORK_CLASSES])
__device__ __forceinline__ uint32_t work_item(const uint32_t* list, uint32_t n_sp, const uint32_t* s_pre,
                                              uint32_t t) {
    uint32_t c = 0;
#pragma unroll
    for (uint32_t k = 1; k < WORK_CLASSES; ++k) c += (t >= s_pre[k]) ? 1u : 0u;
    return __ldg(list + size_t(c) * n_sp + (t - s_pre[c]));
}
// Append middle m (lane-level; nB != 0)
__device__ __forceinline__ void work_push(uint32_t* list, uint32_t n_sp, Control* ctrl, uint32_t m,
                                          uint32_t nB, uint32_t nT) {
    const uint32_t c = work_class(nB, nT);
    list[size_t(c) * n_sp + atomicAdd(&ctrl->n_cls[c], 1u)] = m;
}

// One doublet record in the arena: two float4.
//   a = {cotTheta, iDeltaR, Er, U}     b = {V, Zo, radius(other), bits(sorted pos of other)}
struct __align__(16) DoubletRec {
    float4 a, b;
};

// Debug dump record (32 bytes), see b200seed.h
struct __align__(16) TripletDumpRec {
    uint32_t pos_b, pos_m, pos_t, mb_idx;
    uint32_t mt_idx;
    float curvature, weight, z_vertex;
};

// ---------------------------------------------------------------------------
// (0) spacepoint formation from 2D measurements — the step before seeding
//     (core/include/traccc/seeding/impl/spacepoint_formation.ipp:21-47, host loop
//     core/src/seeding/silicon_pixel_spacepoint_formation.hpp:33-62, device kernel
//     device/common/.../impl/form_spacepoints.ipp:19-51).
// One thread per measurement; global = translation + l0 * x_axis + l1 * y_axis of the placed
// surface. The reference's device kernel appends with an atomic (random order); here the valid
// measurements are compacted in measurement order — the HOST algorithm's order — by a
// single-pass scan with decoupled look-back: every CTA publishes its count, then the
// inclusive prefix, in one 64-bit status word {epoch:30, state:2, value:32}. The epoch and the
// ticket base advance on the host with every call, so nothing has to be cleared between events.
// ---------------------------------------------------------------------------
constexpr int FORM_THREADS = 256;
constexpr int FORM_ITEMS = 1;  // measurements per thread (4 was measured slower: 49 CTAs do not cover the table gathers)
constexpr int FORM_TILE = FORM_THREADS * FORM_ITEMS;
constexpr unsigned long long FORM_AGGREGATE = 1ull, FORM_PREFIX = 2ull;

__device__ __forceinline__ unsigned long long form_pack(uint32_t epoch, unsigned long long state,
                                                        uint32_t value) {
    return ((unsigned long long)(epoch & 0x3FFFFFFFu) << 34) | (state << 32) | value;
}

__global__ void __launch_bounds__(FORM_THREADS)
k_form_spacepoints(const uint32_t n_meas, const float* __restrict__ meas_local,
                   const uint32_t* __restrict__ meas_dim, const uint32_t* __restrict__ meas_surface,
                   const b200seed_surface* __restrict__ surfaces, const uint32_t n_surfaces,
                   float* __restrict__ xyz, float* __restrict__ var_z, float* __restrict__ var_r,
                   uint32_t* __restrict__ mi1, uint32_t* __restrict__ mi2,
                   uint32_t* __restrict__ n_sp_out, unsigned long long* __restrict__ status,
                   unsigned long long* __restrict__ ticket_ctr, const unsigned long long ticket_base,
                   const uint32_t epoch) {
    __shared__ uint32_t s_tile, s_warp[FORM_ITEMS][FORM_THREADS / 32], s_prefix;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // tiles are handed out in the order the CTAs start: a CTA only ever waits for tiles
    // whose CTAs are already running
    if (threadIdx.x == 0) s_tile = uint32_t(atomicAdd(ticket_ctr, 1ull) - ticket_base);
    __syncthreads();
    const uint32_t tile = s_tile;
    // item k of thread t is measurement tile * FORM_TILE + k * FORM_THREADS + t (coalesced)
    uint32_t sf[FORM_ITEMS], bal[FORM_ITEMS];
    bool valid[FORM_ITEMS];
#pragma unroll
    for (int k = 0; k < FORM_ITEMS; ++k) {
        const uint32_t i = tile * FORM_TILE + k * FORM_THREADS + threadIdx.x;
        valid[k] = false;
        sf[k] = 0;
        if (i < n_meas) {
            sf[k] = __ldg(meas_surface + i);
            // "We use 2D (pixel) measurements only" (spacepoint_formation.ipp:19-23)
            valid[k] = (!meas_dim || __ldg(meas_dim + i) == 2u) && sf[k] < n_surfaces;
        }
        bal[k] = __ballot_sync(0xffffffffu, valid[k]);
        if (lane == 0) s_warp[k][warp] = __popc(bal[k]);
    }
    __syncthreads();
    uint32_t before[FORM_ITEMS], total = 0;
#pragma unroll
    for (int k = 0; k < FORM_ITEMS; ++k) {
        before[k] = total;
#pragma unroll
        for (int w = 0; w < FORM_THREADS / 32; ++w) {
            const uint32_t c = s_warp[k][w];
            if (w < int(warp)) before[k] += c;
            total += c;
        }
    }
    if (warp == 0) {
        volatile unsigned long long* st = status;
        uint32_t prefix = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = form_pack(epoch, FORM_PREFIX, total);
        } else {
            if (lane == 0) st[tile] = form_pack(epoch, FORM_AGGREGATE, total);
            // look back over the predecessors, 32 at a time, closest first
            int base = int(tile) - 1;
            while (true) {
                const int j = base - int(lane);
                unsigned long long w = form_pack(epoch, FORM_PREFIX, 0);  // before tile 0: prefix 0
                if (j >= 0) {
                    do {
                        w = st[j];
                    } while ((w >> 34) != (epoch & 0x3FFFFFFFu) || ((w >> 32) & 3ull) == 0ull);
                }
                const uint32_t is_pref = __ballot_sync(0xffffffffu, ((w >> 32) & 3ull) == FORM_PREFIX);
                // lanes up to (and including) the first inclusive prefix contribute
                const uint32_t first = is_pref ? uint32_t(__ffs(int(is_pref)) - 1) : 32u;
                uint32_t v = (lane <= first) ? uint32_t(w) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                prefix += v;
                if (is_pref) break;
                base -= 32;
            }
            if (lane == 0) st[tile] = form_pack(epoch, FORM_PREFIX, prefix + total);
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == (n_meas + FORM_TILE - 1) / FORM_TILE - 1) *n_sp_out = prefix + total;
        }
    }
    __syncthreads();
    const uint32_t prefix = s_prefix;
#pragma unroll
    for (int k = 0; k < FORM_ITEMS; ++k) {
        if (!valid[k]) continue;
        const uint32_t i = tile * FORM_TILE + k * FORM_THREADS + threadIdx.x;
        const uint32_t p = prefix + before[k] + __popc(bal[k] & ((1u << lane) - 1u));
        const float l0 = __ldg(meas_local + 2 * size_t(i)), l1 = __ldg(meas_local + 2 * size_t(i) + 1);
        const b200seed_surface S = surfaces[sf[k]];
        xyz[3 * size_t(p)] = (S.x_axis[0] * l0 + S.y_axis[0] * l1) + S.translation[0];
        xyz[3 * size_t(p) + 1] = (S.x_axis[1] * l0 + S.y_axis[1] * l1) + S.translation[1];
        xyz[3 * size_t(p) + 2] = (S.x_axis[2] * l0 + S.y_axis[2] * l1) + S.translation[2];
        if (var_z) var_z[p] = 0.f;
        if (var_r) var_r[p] = 0.f;
        if (mi1) mi1[p] = i;
        if (mi2) mi2[p] = 0xFFFFFFFFu;  // INVALID_MEASUREMENT_INDEX (spacepoint_collection.hpp:47-48)
    }
}

// Number of spacepoints of the event: the host's value, or — when the spacepoints were made on
// the device by k_form_spacepoints and their number never went to the host — the device word
// clamped to the host's upper bound (the grid is sized for the bound).
__device__ __forceinline__ uint32_t dev_count(uint32_t n_max, const uint32_t* n_dev) {
    if (!n_dev) return n_max;
    const uint32_t n = *n_dev;
    return n < n_max ? n : n_max;
}

// ---------------------------------------------------------------------------
// (1) binning: count
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BIN_THREADS)
k_bin_count(const DevCfg cfg, const CellGrid g, const uint32_t n_sp_max,
            const float* __restrict__ xyz,
            uint32_t* __restrict__ bin_of, uint32_t* __restrict__ blk_hist,
            uint32_t* __restrict__ cell_cnt, const uint32_t nbins, const uint32_t nblk,
            const uint32_t* __restrict__ n_sp_dev, uint32_t* __restrict__ bin_tot) {
    extern __shared__ uint32_t s_hist[];
    for (uint32_t b = threadIdx.x; b < nbins; b += BIN_THREADS) s_hist[b] = 0;
    __syncthreads();
    const uint32_t n_sp = dev_count(n_sp_max, n_sp_dev);
    const uint32_t i = blockIdx.x * BIN_THREADS + threadIdx.x;
    if (i < n_sp) {
        const float x = __ldg(xyz + 3 * size_t(i));
        const float y = __ldg(xyz + 3 * size_t(i) + 1);
        const float z = __ldg(xyz + 3 * size_t(i) + 2);
        const uint32_t bin = sp_bin(cfg, x, y, z);
        bin_of[i] = bin;
        if (bin != INVALID_BIN) {
            atomicAdd(&s_hist[bin], 1u);
            atomicAdd(&cell_cnt[sp_cell(cfg, g, bin, sp_radius(x, y), z)], 1u);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += BIN_THREADS) {
        const uint32_t c = s_hist[b];
        blk_hist[size_t(b) * nblk + blockIdx.x] = c;
        if (c) atomicAdd(&bin_tot[b], c);  // population of the bin (k_cell_scan sums them into bin_off)
    }
}

// ---------------------------------------------------------------------------
// (1) binning: stable scatter into the bin-sorted SoA
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BIN_THREADS)
k_bin_scatter(const DevCfg cfg, const CellGrid g, const uint32_t n_sp_max,
              const float* __restrict__ xyz, const float* __restrict__ var_z,
              const float* __restrict__ var_r, const uint32_t* __restrict__ bin_of,
              const uint32_t* __restrict__ blk_scan, const uint32_t nblk,
              float4* __restrict__ sp4, float2* __restrict__ var2,
              uint32_t* __restrict__ sorted_index, uint32_t* __restrict__ sorted_bin,
              const uint32_t* __restrict__ cell_off, uint32_t* __restrict__ cell_cur,
              float4* __restrict__ csp4, uint32_t* __restrict__ ccanon,
              const uint32_t* __restrict__ n_sp_dev, Control* __restrict__ ctrl, const uint32_t nbins,
              const uint32_t* __restrict__ seg_info, uint32_t* __restrict__ mid_order) {
    // rank among the earlier spacepoints of this block that fall into the same bin: ascending
    // original index inside every bin, like the CPU's push_back loop
    // (core/src/seeding/spacepoint_binning.cpp:40-50). Inside a warp: __match_any_sync + popc of
    // the lower lanes; across the warps of the block: a shared-memory histogram (dynamic, nbins
    // words) that the warps update one after the other, in order.
    extern __shared__ uint32_t s_hist[];
    __shared__ uint32_t s_dpre[WORK_CLASSES];  // start of every cost class in mid_order
    // (loaded now, summed after the ranking loop: the round trip overlaps it)
    const uint32_t dcl = (seg_info && threadIdx.x < WORK_CLASSES) ? ctrl->n_dcls[threadIdx.x] : 0u;
    const uint32_t n_sp = dev_count(n_sp_max, n_sp_dev);
    const uint32_t i = blockIdx.x * BIN_THREADS + threadIdx.x;
    const uint32_t bin = (i < n_sp) ? bin_of[i] : INVALID_BIN;
    const uint32_t lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
    for (uint32_t b = threadIdx.x; b < nbins; b += BIN_THREADS) s_hist[b] = 0u;
    const uint32_t peers = __match_any_sync(0xffffffffu, bin);
    const uint32_t in_warp = __popc(peers & ((1u << lane_) - 1u));
    const uint32_t leader = __ffs(int(peers)) - 1u;
    uint32_t before = 0;
    __syncthreads();
    for (uint32_t w = 0; w < BIN_THREADS / 32; ++w) {
        if (warp_ == w && bin != INVALID_BIN && lane_ == leader) {
            before = s_hist[bin];  // one leader per distinct bin: no conflict inside the warp
            s_hist[bin] = before + __popc(peers);
        }
        __syncthreads();
    }
    before = __shfl_sync(0xffffffffu, before, leader);
    if (seg_info) {
        if (warp_ == 0) {
            uint32_t incl = dcl;
#pragma unroll
            for (int o = 1; o < int(WORK_CLASSES); o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane_ >= uint32_t(o)) incl += t;
            }
            if (lane_ < WORK_CLASSES) s_dpre[lane_] = incl - dcl;
        }
        __syncthreads();
    }
    if (bin == INVALID_BIN) return;
    const uint32_t rank = before + in_warp;
    const uint32_t pos = blk_scan[size_t(bin) * nblk + blockIdx.x] + rank;
    const float x = __ldg(xyz + 3 * size_t(i));
    const float y = __ldg(xyz + 3 * size_t(i) + 1);
    const float z = __ldg(xyz + 3 * size_t(i) + 2);
    const float r = sp_radius(x, y);
    const float4 P = make_float4(x, y, z, r);
    sp4[pos] = P;
    const float2 VV = make_float2(var_z ? __ldg(var_z + i) : 0.f, var_r ? __ldg(var_r + i) : 0.f);
    var2[pos] = VV;
    // the reference's standard input has zero variances (read_spacepoints.cpp:72-77): the search
    // kernels then skip the per-partner variance loads (same arithmetic with the zeros in place)
    if (VV.x != 0.f || VV.y != 0.f) ctrl->has_variance = 1u;
    sorted_index[pos] = i;
    sorted_bin[pos] = bin;
    // cell-sorted copy (order inside a cell is arbitrary: nothing downstream depends on it)
    const uint32_t cell = sp_cell(cfg, g, bin, r, z);
    // ticket order of k_doublets: the segment of this (bin, r row) inside its cost class (k_cell_scan)
    uint32_t info = 0, row_start = 0;
    if (seg_info) {
        const uint32_t row = cell_row(g, r);
        info = __ldg(seg_info + bin * g.NR + row);
        row_start = __ldg(cell_off + size_t(bin) * g.CPB + row * g.NZc);
    }
    const uint32_t cpos = cell_off[cell] + atomicAdd(&cell_cur[cell], 1u);
    csp4[cpos] = P;
    ccanon[cpos] = pos;
    if (seg_info)
        mid_order[s_dpre[info >> 28] + (info & 0x07FFFFFFu) + (cpos - row_start)] =
            pos | ((info & (1u << 27)) ? MID_LOWER_SCARCE : 0u);
}

// CTA per reference bin: cell_off[bin * CPB + c] = bin_off[bin] + exclusive scan of the
// cell populations of this bin; the population array is zeroed again (k_bin_scatter uses
// it as its cursor). cell_off[nbins * CPB] = n_valid.
// With group_list != null the CTA also cuts the cell-ordered content of its bin into the work
// items of k_doublets_tile: a group = up to gmax consecutive (in cell order) spacepoints of one
// r row inside one block of zspan z cells; descriptor = (first cell-ordered position << 5) |
// (size - 1). Groups of at least `big` middles are listed from the front (drawn first), the
// others from the back of group_list[n_sp].
__global__ void __launch_bounds__(256)
k_cell_scan(uint32_t* __restrict__ cell_cnt, uint32_t* __restrict__ cell_off,
            uint32_t* __restrict__ bin_off, const uint32_t* __restrict__ bin_tot,
            uint32_t* __restrict__ blk_hist, const uint32_t nblk, const uint32_t CPB, const uint32_t nbins,
            uint32_t* __restrict__ group_list, Control* __restrict__ ctrl, const uint32_t NZc,
            const uint32_t gmax, const uint32_t zspan, const uint32_t big, const uint32_t n_sp,
            uint32_t* __restrict__ seg_info, const uint32_t row_reach, const bool sided) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_total;
    const uint32_t bin = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // exclusive scan of len words at p (in place) continuing from `from`; returns the end value
    auto scan_in_place = [&](uint32_t* p, const uint32_t len, const uint32_t from, uint32_t* copy_to,
                             const bool zero_src) -> uint32_t {
        uint32_t run0 = from;
        for (uint32_t base = 0; base < len; base += 256 * 4) {
            const uint32_t idx = base + threadIdx.x * 4;
            uint32_t v[4];
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = (idx + k < len) ? p[idx + k] : 0u;
                sum += v[k];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= uint32_t(off)) incl += t;
            }
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const uint32_t ws = (lane < 8) ? s_warp[lane] : 0u;
                uint32_t wincl = ws;
#pragma unroll
                for (int off = 1; off < 8; off <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, off);
                    if (lane >= uint32_t(off)) wincl += t;
                }
                if (lane < 8) s_warp[lane] = wincl - ws;
                if (lane == 7) s_total = wincl;
            }
            __syncthreads();
            uint32_t run = run0 + s_warp[warp] + incl - sum;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (idx + k < len) {
                    (copy_to ? copy_to : p)[idx + k] = run;
                    if (zero_src) p[idx + k] = 0u;
                }
                run += v[k];
            }
            run0 += s_total;
            __syncthreads();
        }
        return run0;
    };
    // Start of this bin in the grid order = populations of the bins before it (k_bin_count's
    // totals; this used to be a single-CTA scan kernel of its own between the two).
    uint32_t part = 0;
    for (uint32_t b = threadIdx.x; b < bin; b += 256) part += __ldg(bin_tot + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_warp[warp] = part;
    __syncthreads();
    uint32_t bin_base = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) bin_base += s_warp[w];
    __syncthreads();
    if (threadIdx.x == 0) {
        bin_off[bin] = bin_base;
        if (bin == nbins - 1) {
            const uint32_t n_valid = bin_base + __ldg(bin_tot + bin);
            bin_off[nbins] = n_valid;
            ctrl->n_valid = n_valid;
            if (!sided) ctrl->n_main = n_valid;
        }
    }
    // position of the first spacepoint every block of k_bin_count / k_bin_scatter has in this bin
    scan_in_place(blk_hist + size_t(bin) * nblk, nblk, bin_base, nullptr, false);
    // cell offsets of this bin; the populations are zeroed again (k_bin_scatter's cursors)
    const uint32_t carry = scan_in_place(cell_cnt + size_t(bin) * CPB, CPB, bin_base,
                                         cell_off + size_t(bin) * CPB, true);
    if (bin == nbins - 1 && threadIdx.x == 0) cell_off[size_t(nbins) * CPB] = carry;
    if (seg_info) {
        // Ticket order of k_doublets: the spacepoints of one (bin, r row) form a segment of the
        // cell order and share a cost class, estimated from what can be known here — the
        // populations of the rows below and above within deltaRMax (own row left out: it holds
        // partners of both kinds). A middle with an empty side cannot seed and costs one scan;
        // the more partners on the scarcer side, the more doublets, lin_circles and sorting.
        // Longest first, so that the launch ends on its shortest jobs (see work_class).
        __syncthreads();  // the offsets of this bin, written above by other threads of the CTA
        if (warp == 0) {
            const uint32_t NR = CPB / NZc;
            const uint32_t b0 = __ldg(bin_off + bin);
            const uint32_t start = (lane < NR) ? cell_off[size_t(bin) * CPB + lane * NZc] : carry;
            const uint32_t next = __shfl_down_sync(0xffffffffu, start, 1);
            const uint32_t end = (lane + 1u < NR) ? next : carry;
            const uint32_t pop = (lane < NR) ? end - start : 0u;
            uint32_t below = 0, above = 0;
            for (uint32_t d = 1; d <= row_reach; ++d) {
                const uint32_t lo = __shfl_sync(0xffffffffu, pop, (lane - d) & 31u);
                const uint32_t hi = __shfl_sync(0xffffffffu, pop, (lane + d) & 31u);
                if (lane >= d) below += lo;
                if (lane + d < NR) above += hi;
            }
            if (lane < NR && pop != 0u) {
                const uint32_t nb = carry - b0;  // population of the bin
                const uint32_t sc = below < above ? below : above;
                uint32_t c;
                if (sc == 0u) c = ((below + above) * 4u >= nb) ? 6u : 7u;
                else c = (sc * 8u >= nb) ? 0u : (sc * 16u >= nb) ? 1u : (sc * 32u >= nb) ? 2u
                       : (sc * 64u >= nb) ? 3u : (sc * 256u >= nb) ? 4u : 5u;
                const uint32_t base = atomicAdd(&ctrl->n_dcls[c], pop);
                if (sided && c < WORK_CLASSES / 2) atomicAdd(&ctrl->n_main, pop);
                // bit 27: the scarcer side is the lower one (only read for the classes >= 4)
                seg_info[bin * NR + lane] = (c << 28) | ((sided && c >= 4u && below < above) ? (1u << 27) : 0u) | base;
            }
        }
    }
    if (!group_list) return;
    // the offsets of this bin in shared memory (dynamic, CPB + 1 words)
    extern __shared__ uint32_t s_off[];
    __syncthreads();  // the offsets of this bin, written above by other threads of the CTA
    for (uint32_t c = threadIdx.x; c < CPB; c += 256) s_off[c] = cell_off[size_t(bin) * CPB + c];
    if (threadIdx.x == 0) s_off[CPB] = carry;  // end of the bin
    __syncthreads();
    const uint32_t NR = CPB / NZc;
    // groups of this CTA are collected in shared memory and appended with two atomics
    __shared__ uint32_t s_nbig, s_nsmall, s_bbig, s_bsmall;
    if (threadIdx.x == 0) s_nbig = s_nsmall = 0u;
    __syncthreads();
    // one thread per (row, block of zspan z cells): the block's spacepoints, in near-equal
    // pieces of at most gmax. pass 0 counts, pass 1 writes: the CTA's groups land in two
    // contiguous runs of group_list (two global atomics per CTA).
    const uint32_t zblocks = (NZc + zspan - 1u) / zspan;
    for (int pass = 0; pass < 2; ++pass) {
        for (uint32_t bi = threadIdx.x; bi < NR * zblocks; bi += 256) {
            const uint32_t row = bi / zblocks, zb0 = (bi - row * zblocks) * zspan;
            const uint32_t zb1 = (zb0 + zspan < NZc) ? (zb0 + zspan) : NZc;
            const uint32_t start = s_off[row * NZc + zb0];
            const uint32_t n = s_off[row * NZc + zb1] - start;
            if (n == 0u) continue;
            const uint32_t pieces = (n + gmax - 1u) / gmax;
            if (pass == 0) {
                // pieces differ by at most one in size: all >= big or all < big unless n / pieces
                // sits on the threshold; count them one by one (pieces is small)
                uint32_t nb = 0;
                for (uint32_t k = 0; k < pieces; ++k)
                    nb += (uint32_t((unsigned long long)n * (k + 1u) / pieces) -
                               uint32_t((unsigned long long)n * k / pieces) >= big) ? 1u : 0u;
                if (nb) atomicAdd(&s_nbig, nb);
                if (pieces - nb) atomicAdd(&s_nsmall, pieces - nb);
            } else {
                for (uint32_t k = 0; k < pieces; ++k) {
                    const uint32_t a0 = uint32_t((unsigned long long)n * k / pieces);
                    const uint32_t a1 = uint32_t((unsigned long long)n * (k + 1u) / pieces);
                    const uint32_t sz = a1 - a0;
                    const uint32_t slot = (sz >= big) ? (s_bbig + atomicAdd(&s_nbig, 1u))
                                                      : n_sp - 1u - (s_bsmall + atomicAdd(&s_nsmall, 1u));
                    group_list[slot] = ((start + a0) << 5) | (sz - 1u);
                }
            }
        }
        __syncthreads();
        if (pass == 0 && threadIdx.x == 0) {
            s_bbig = s_nbig ? atomicAdd(&ctrl->n_group_big, s_nbig) : 0u;
            s_bsmall = s_nsmall ? atomicAdd(&ctrl->n_group_small, s_nsmall) : 0u;
            s_nbig = s_nsmall = 0u;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// (2)+(3) doublets
// ---------------------------------------------------------------------------
struct DoubletArgs {
    const uint32_t* bin_off;      // [nbins+1]
    const float4* sp4;            // sorted {x,y,z,r} (reference grid order)
    const float2* var2;           // sorted {varZ,varR}
    const uint32_t* sorted_bin;   // [n_valid]
    const uint32_t* cell_off;     // [nbins*CPB+1] start of every cell in the cell-sorted arrays
    const float4* csp4;           // cell-sorted {x,y,z,r}
    const uint32_t* ccanon;       // cell-sorted -> sorted position
    uint32_t* cnt_b;              // [n_sp] doublets per middle (0 if inactive)
    uint32_t* cnt_t;
    uint32_t* off_b;              // [n_sp] arena offsets
    uint32_t* off_t;
    DoubletRec* arena_b;
    DoubletRec* arena_t;
    Control* ctrl;
    CellGrid g;
    uint32_t max_doublets;
    uint32_t cap_b, cap_t;        // staged doublets per warp (shared memory)
    uint32_t* spill_list;         // [n_sp] middles whose lists do not fit the staging area
    uint32_t* active_list;        // [n_sp] middles with doublets on both sides = work list of
                                  // k_triplets: heavy ones from the front, light ones from the back
    uint32_t* seed_cnt;           // [n_sp] set to 0 here for the middles without work
    uint32_t n_sp;
    const uint32_t* fallback_list;  // [n_sp] middles handed back by k_doublets_tile (MODE 2)
    const uint32_t* mid_order;      // [n_sp] ticket order of MODE 0 (null: grid order)
    uint32_t split_sides;           // the classes with a scarce side go to the MODE 3 launch
};

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= uint32_t(o)) v += t;
    }
    return v;
}

// Lane that owns flattened element p, given the inclusive prefix sums `incl` of the lanes'
// segment lengths (p < total): the number of lanes whose inclusive prefix is <= p.
__device__ __forceinline__ uint32_t owner_lane(uint32_t incl, uint32_t p) {
    uint32_t r = 0;
#pragma unroll
    for (uint32_t step = 16; step >= 1; step >>= 1) {
        const uint32_t v = __shfl_sync(0xffffffffu, incl, (r + step - 1) & 31u);
        if (v <= p) r += step;
    }
    return r & 31u;
}

// The neighbour bins of a middle in the reference's order: phi bins outer (circular zone,
// axis.hpp:336-362), z bins inner (regular zone, axis.hpp:162-173), as in
// doublet_finding.hpp:67-80. Neighbour q (0 <= q < nq) is bin(q); q / nz is the position of
// its phi bin in the walk.
// a / b for small operands (a < 2^21) with a precomputed 1.0f / b: exact, because
// (a + 0.5) / b is at least 0.5 / b away from the next integer.
__device__ __forceinline__ uint32_t div_small(uint32_t a, float inv_b) {
    return uint32_t((float(a) + 0.5f) * inv_b);
}

struct NeighbourWalk {
    uint32_t r0, n_phi_seq, z0, nz, nq;
    float inv_nz;
    __device__ __forceinline__ void init(const DevCfg& cfg, uint32_t bin, float zM) {
        const uint32_t phi_bin = bin % cfg.nPhi;
        r0 = circular_remap(cfg.nPhi, phi_bin, -int(cfg.scope0));
        const uint32_t r1 = circular_remap(cfg.nPhi, phi_bin, int(cfg.scope1));
        n_phi_seq = (r0 < r1) ? (r1 - r0 + 1u) : (cfg.nPhi - r0 + r1 + 1u);
        const int ibin = axis_ibin(cfg.zAxisMin, cfg.zAxisMax, cfg.nZ, zM);
        const int ibinmin = ibin - int(cfg.scope0);
        const int ibinmax = ibin + int(cfg.scope1);
        z0 = (ibinmin >= 0) ? uint32_t(ibinmin) : 0u;
        const uint32_t z1 = (ibinmax < int(cfg.nZ)) ? uint32_t(ibinmax) : cfg.nZ - 1u;
        nz = (z1 + 1u > z0) ? (z1 + 1u - z0) : 0u;
        nq = n_phi_seq * nz;
        inv_nz = 1.f / float(nz ? nz : 1u);
    }
    // position of neighbour q's phi bin in the walk
    __device__ __forceinline__ uint32_t wphi(uint32_t q) const { return div_small(q, inv_nz); }
    __device__ __forceinline__ uint32_t zbin(uint32_t q) const { return z0 + (q - wphi(q) * nz); }
    __device__ __forceinline__ uint32_t bin(const DevCfg& cfg, uint32_t q) const {
        uint32_t pb = r0 + wphi(q);
        if (pb > cfg.nPhi - 1u) pb -= cfg.nPhi;
        return pb + zbin(q) * cfg.nPhi;
    }
};

// Position of a spacepoint in the reference's candidate order of one middle: neighbour phi
// bins in walk order, then the grid order (z bins ascending, position in bin).
__device__ __forceinline__ uint32_t canon_key(uint32_t walk_phi, uint32_t n_valid, uint32_t pos) {
    return walk_phi * n_valid + pos;
}

// Position of element (ck, kk) of a mid-top list in the (cotTheta, reference order) sort;
// the keys are unique.
template <typename LoadCot, typename LoadKey>
__device__ __forceinline__ uint32_t top_rank(LoadCot cot, LoadKey key, uint32_t n, float ck,
                                             uint32_t kk) {
    uint32_t ks = 0;
    for (uint32_t j = 0; j < n; ++j) {
        const float cj = cot(j);
        ks += ((cj < ck) || (cj == ck && key(j) < kk)) ? 1u : 0u;
    }
    return ks;
}

// Shared memory per warp of k_doublets (32-bit words).
__host__ __device__ inline uint32_t doublet_smem_words(uint32_t cap_b, uint32_t cap_t) {
    return cap_b + 3u * cap_t;
}

// One LANE decides whether middle m has any doublet partner on its scarce side (lower == true:
// a bottom partner in the rows up to its own; else a top partner from its own row upwards): the
// same windows, cells and cuts as the warp-wide scan of k_doublets, one candidate after the
// other. 32 middles per warp at a time instead of one: the middles of these classes almost never
// have one (barrel layers without a layer below / above them), and a warp-wide pass over a
// handful of candidates is all latency. n_pairs: the middle's contribution to pair_tests.
__device__ __forceinline__ bool sided_has_partner(const DevCfg& cfg, const DoubletArgs& a, const CellGrid& g,
                                                  const uint32_t m, const bool lower, const bool bounded,
                                                  uint32_t& n_pairs, uint32_t& n_visited) {
    const float4 M = __ldg(a.sp4 + m);
    NeighbourWalk walk;
    walk.init(cfg, __ldg(a.sorted_bin + m), M.z);
    n_pairs = 0;
    for (uint32_t q = 0; q < walk.nq; ++q) {
        const uint32_t b = walk.bin(cfg, q);
        n_pairs += __ldg(a.bin_off + b + 1) - __ldg(a.bin_off + b);
    }
    const float er = 1e-2f + 1e-5f * (M.w + absf(cfg.deltaRMax));
    const uint32_t row_lo = cell_row(g, M.w - cfg.deltaRMax - er);
    const uint32_t row_hi = cell_row(g, M.w + cfg.deltaRMax + er);
    const uint32_t rowM = cell_row(g, M.w);
    const uint32_t r0 = lower ? row_lo : (rowM > row_lo ? rowM : row_lo);
    const uint32_t r1 = lower ? (rowM < row_hi ? rowM : row_hi) : row_hi;
    const int want = lower ? 1 : 2;
    for (uint32_t row = r0; row <= r1; ++row) {
        float L, U;
        if (!cell_row_window(cfg, g, M.w, M.z, row, L, U)) continue;
        for (uint32_t q = 0; q < walk.nq; ++q) {
            const uint32_t zb = walk.zbin(q);
            const uint32_t base = walk.bin(cfg, q) * g.CPB + row * g.NZc;
            const uint32_t lo = __ldg(a.cell_off + base + cell_z(g, zb, L));
            const uint32_t hi = __ldg(a.cell_off + base + cell_z(g, zb, U) + 1u);
            n_visited += hi - lo;
            for (uint32_t c = lo; c < hi; ++c) {
                const float4 P = __ldg(a.csp4 + c);
                if (doublet_stage1(cfg, M.w, M.z, P.w, P.z) != want) continue;
                int d = bounded ? doublet_stage2_fast_bounded(cfg, M.x, M.y, P.x, P.y)
                                : doublet_stage2_fast(cfg, M.x, M.y, P.x, P.y);
                if (d == 2) d = doublet_stage2(cfg, M.x, M.y, P.x, P.y) ? 1 : 0;
                if (d != 0) return true;
            }
        }
    }
    return false;
}

// Arena records. The order of a mid-bottom list is arbitrary (nothing downstream depends on
// it; the parity tests sort it by canon_key). Mid-top lists are sorted by cotTheta so that
// k_triplets can binary-search the scattering window of each mid-bottom doublet, and carry
// their index in the reference's order:
//   bottom: a = {cotTheta, iDeltaR, Er, U}  b = {V, Zo,               r_other, pos_other}
//   top   : a = {cotTheta, iDeltaR, Er, U}  b = {V, bits(canon_key),  r_other, pos_other}
#ifndef B200_DOUBLET_MIN_CTAS
#define B200_DOUBLET_MIN_CTAS (32 / B200_WARPS_PER_CTA)
#endif
// SPILL == false: every middle; a middle whose lists do not fit the shared-memory staging area
// is only recorded in spill_list. SPILL == true: those (rare) middles, two scans each — count,
// then write straight to the arena and bucket-sort the mid-tops. Two instantiations because the
// second scan and the sort cost the common kernel 76 bytes of register spills when they
// live in the same function (169 -> 153 us for the 10k-particle event).
// MODE 0: every middle (the whole search in this kernel; kept as the A/B baseline of
// k_doublets_tile, B200SEED_DOUBLETS=legacy). MODE 1: the spill pass. MODE 2: the middles
// k_doublets_tile handed back (fallback_list): groups whose doublets outgrow its queues.
template <int MODE>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, B200_DOUBLET_MIN_CTAS)
k_doublets(const DevCfg cfg, const DoubletArgs a) {
    constexpr bool SPILL = (MODE == 1);
    constexpr bool LISTED = (MODE == 2);
    constexpr bool SIDES = (MODE == 3);
    extern __shared__ __align__(16) uint32_t s_mem[];
    __shared__ unsigned long long s_pairs[2];
    __shared__ uint32_t s_acc[3];  // active, nb, nt
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltmask = lanemask_lt();
    uint32_t* stage_b = s_mem + size_t(warp) * doublet_smem_words(a.cap_b, a.cap_t);
    uint32_t* stage_t = stage_b + a.cap_b;
    uint32_t* key_s = stage_t + a.cap_t;
    float* cot_s = reinterpret_cast<float*>(key_s + a.cap_t);
    if (threadIdx.x == 0) {
        s_pairs[0] = s_pairs[1] = 0ull;
        s_acc[0] = s_acc[1] = s_acc[2] = 0u;
    }
    // k_doublets<3> (launched behind this one with programmatic stream serialization) shares no
    // data with this launch: its CTAs may take the slots this launch frees while its last middles
    // are still running
    if (MODE == 0) asm volatile("griddepcontrol.launch_dependents;");
    __syncthreads();
    const uint32_t n_valid = a.ctrl->n_valid;
    const bool has_var = a.ctrl->has_variance != 0u;
    const bool bounded = cfg.fast_bounded != 0u;  // the pre-decision needs no magnitude guards
    const CellGrid g = a.g;
    unsigned long long pairs = 0ull, visited = 0ull;  // per lane
    uint32_t acc_active = 0, acc_nb = 0, acc_nt = 0;

    // With a cost-ordered ticket list (mid_order) the classes whose row populations leave one side
    // (almost) empty — the last WORK_CLASSES / 2 — belong to the MODE 3 launch.
    // (n_main: the tickets of MODE 0, written by k_cell_scan — all valid spacepoints without the split)
    const uint32_t n_work = SPILL ? a.ctrl->n_spill
                                  : (LISTED ? a.ctrl->n_fallback
                                            : (SIDES ? n_valid - a.ctrl->n_main : a.ctrl->n_main));
    if ((SPILL || LISTED || SIDES) && n_work == 0u) {  // the usual case: no ticket traffic at all
        if (SIDES && blockIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");  // (see the end of the kernel)
        return;
    }
    // MODE 3: the middles come in batches of 32, one per lane for the pre-screening of the scarce
    // side; the (rare) survivors are then processed one after the other like any middle
    uint32_t alive_mask = 0, batch_entry = 0;
    while (true) {
        uint32_t m = 0;
        if (SIDES) {
            bool exhausted = false;
            while (alive_mask == 0u) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&a.ctrl->ticket_e, 32u);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n_work) {
                    exhausted = true;
                    break;
                }
                bool alive = false;
                uint32_t seen = 0;
                batch_entry = 0u;
                if (base + lane < n_work) {
                    batch_entry = __ldg(a.mid_order + (n_valid - n_work) + base + lane);
                    const uint32_t mm = batch_entry & ~MID_LOWER_SCARCE;
                    uint32_t np = 0;
                    alive = !(cfg.deltaRMin >= 0.f) ||
                            sided_has_partner(cfg, a, g, mm, (batch_entry & MID_LOWER_SCARCE) != 0u, bounded, np, seen);
                    if (!alive) {  // cannot seed (seed_finding.cpp:85-95): what the full scan would leave
                        pairs += np;
                        a.cnt_b[mm] = 0u, a.cnt_t[mm] = 0u, a.off_b[mm] = 0u, a.off_t[mm] = 0u;
                        a.seed_cnt[mm] = 0u;
                    }
                }
                alive_mask = __ballot_sync(0xffffffffu, alive);
                // This warp scans at most SIDED_INLINE of the survivors itself; more than that (not
                // in the toy detector, where ~1 % survive) go to the fallback list, which
                // k_doublets<2> works off with a warp each — a batch full of survivors must not
                // become 32 warp-wide scans in a row on one warp.
                if (__popc(alive_mask) > int(SIDED_INLINE)) {
                    uint32_t rest = alive_mask, keep = 0;
#pragma unroll
                    for (uint32_t k = 0; k < SIDED_INLINE; ++k) {
                        const uint32_t b = rest & (0u - rest);
                        keep |= b;
                        rest ^= b;
                    }
                    if ((rest >> lane) & 1u)
                        const_cast<uint32_t*>(a.fallback_list)[atomicAdd(&a.ctrl->n_fallback, 1u)] =
                            batch_entry & ~MID_LOWER_SCARCE;
                    alive_mask = keep;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) seen += __shfl_xor_sync(0xffffffffu, seen, o);
                if (lane == 0) visited += seen;
            }
            if (exhausted) break;
            const uint32_t src = __ffs(int(alive_mask)) - 1u;
            alive_mask &= alive_mask - 1u;
            m = __shfl_sync(0xffffffffu, batch_entry, src);
        } else {
        if (lane == 0)
            m = atomicAdd(SPILL ? &a.ctrl->ticket_s : (LISTED ? &a.ctrl->ticket_f : &a.ctrl->ticket_d), 1u);
        m = __shfl_sync(0xffffffffu, m, 0);
        }
#ifdef B200_TAIL_PROBE
        if (MODE == 0 && lane == 0) {
            const uint32_t w = blockIdx.x * WARPS_PER_CTA + warp;
            if (w < 16384u) {
                if (g_tail_probe[0][0][w] == 0ull) g_tail_probe[0][0][w] = probe_now();
                if (m >= n_work) g_tail_probe[0][1][w] = probe_now();
            }
        }
        const long long probe_c0 = clock64();
#endif
        if (!SIDES && m >= n_work) break;
        if (SPILL) m = a.spill_list[m];
        if (LISTED) m = a.fallback_list[m];
        if (MODE == 0 && a.mid_order) m = __ldg(a.mid_order + m);  // longest middles first
        bool lower_scarce = false;
        if (SIDES) {
            lower_scarce = (m & MID_LOWER_SCARCE) != 0u;
            m &= ~MID_LOWER_SCARCE;
        }
#ifdef B200_CELL_ORDER_TICKETS
        // tickets in CELL order: the warps of a CTA then work on middles of the same
        // (bin, r row, z cell) neighbourhood at the same time and share their candidate cells in L1
        if (MODE == 0) m = __ldg(a.ccanon + m);
#endif
        const float4 M = __ldg(a.sp4 + m);
        const float2 VM = __ldg(a.var2 + m);  // {varZ, varR}
        // lin_circle's cosPhiM / sinPhiM (doublet_finding_helper.hpp:225-226) depend on the middle only
        const float cosM = M.x / M.w, sinM = M.y / M.w;
        NeighbourWalk walk;
        walk.init(cfg, __ldg(a.sorted_bin + m), M.z);
        // the reference tests every spacepoint of the neighbour bins: count them
        if (!SPILL) {
            for (uint32_t q = lane; q < walk.nq; q += 32) {
                const uint32_t b = walk.bin(cfg, q);
                pairs += __ldg(a.bin_off + b + 1) - __ldg(a.bin_off + b);
            }
        }
        // rows that can hold a doublet partner: |r - rM| <= deltaRMax
        const float er = 1e-2f + 1e-5f * (M.w + absf(cfg.deltaRMax));
        const uint32_t row_lo = cell_row(g, M.w - cfg.deltaRMax - er);
        const uint32_t row_hi = cell_row(g, M.w + cfg.deltaRMax + er);
        const uint32_t nrows = row_hi - row_lo + 1u;  // <= NR <= 32
        const uint32_t ncombo = walk.nq * nrows;
        const float inv_nrows = 1.f / float(nrows);
        // z window of every row (lane = row); the same for all neighbour bins
        float winL = 0.f, winU = 0.f;
        bool win_ok = false;
        if (lane < nrows) win_ok = cell_row_window(cfg, g, M.w, M.z, row_lo + lane, winL, winU);
        const uint32_t win_mask = __ballot_sync(0xffffffffu, win_ok);

        uint32_t nB = 0, nT = 0;
        uint32_t offB = 0, offT = 0;
        // pass 0: stage the doublets' cell-sorted positions in shared memory. pass 1 (only
        // when a staging list overflowed): same scan, records written straight to the arena.
        bool deferred = false;  // handed to the SPILL instantiation
        // room in the arena for the two lists of this middle (bump allocation)
        auto allocate = [&](const uint32_t allocT) -> bool {
            if (lane == 0) {
                offB = atomicAdd(&a.ctrl->cursor[0], nB);
                offT = atomicAdd(&a.ctrl->cursor[1], allocT);
            }
            offB = __shfl_sync(0xffffffffu, offB, 0);
            offT = __shfl_sync(0xffffffffu, offT, 0);
            if (offB > a.max_doublets || nB > a.max_doublets - offB || offT > a.max_doublets ||
                allocT > a.max_doublets - offT) {
                if (lane == 0) atomicOr(&a.ctrl->overflow, B200SEED_OVF_DOUBLETS);
                nB = nT = 0;
                return false;
            }
            return true;
        };
        bool run = true;
        if (SPILL) {
            // the common kernel counted this middle's doublets already; unsorted mid-tops go to
            // the second half of a 2 * nT allocation (room to sort in place)
            nB = a.cnt_b[m];
            nT = a.cnt_t[m];
            run = allocate(2u * nT);
        }
        for (int pass = SPILL ? 1 : 0; run && pass < (SPILL ? 2 : 1); ++pass) {
            const bool direct = SPILL;
            uint32_t wB = 0, wT = 0;  // doublets found so far in this pass
            // MODE 3: the rows of the scarce side first (bottoms sit in the rows up to the
            // middle's own, tops from its own row upwards: deltaRMin >= 0); if they hold no
            // partner the middle cannot seed (seed_finding.cpp:85-95) and the other, populated
            // side is never read. All other modes: one segment, all rows.
            uint32_t split = 0;  // rows [0, split) form the lower segment
            if (SIDES && cfg.deltaRMin >= 0.f) {
                const uint32_t rowM = cell_row(g, M.w);
                const uint32_t iM = (rowM > row_lo) ? ((rowM - row_lo < nrows) ? rowM - row_lo : nrows) : 0u;
                // lower side scarce: [0, iM] | (iM, nrows); upper side scarce: [iM, nrows) | [0, iM)
                split = lower_scarce ? ((iM + 1u < nrows) ? iM + 1u : nrows) : iM;
            }
            for (uint32_t seg = 0; seg < (SIDES ? 2u : 1u); ++seg) {
            const bool low_seg = SIDES && (lower_scarce ? seg == 0u : seg == 1u);
            const uint32_t rbase = SIDES ? (low_seg ? 0u : split) : 0u;
            const uint32_t nr = SIDES ? (low_seg ? split : nrows - split) : nrows;
            if (SIDES && (nr == 0u || (seg == 1u && split != 0u && split != nrows &&
                                       (lower_scarce ? wB == 0u : wT == 0u))))
                continue;
            const uint32_t ncombo_s = SIDES ? walk.nq * nr : ncombo;
            const float inv_nr = SIDES ? 1.f / float(nr) : inv_nrows;
            for (uint32_t j0 = 0; j0 < ncombo_s; j0 += 32) {
                // lane j: one (neighbour bin, row) -> contiguous run of cells
                const uint32_t j = j0 + lane;
                uint32_t lo = 0, len = 0, wphi = 0;
                const uint32_t jj = (j < ncombo_s) ? j : 0u;
                const uint32_t q = div_small(jj, inv_nr), ri = rbase + (jj - q * nr);
                const float L = __shfl_sync(0xffffffffu, winL, ri);
                const float U = __shfl_sync(0xffffffffu, winU, ri);
                if (j < ncombo_s && ((win_mask >> ri) & 1u)) {
                    const uint32_t zb = walk.zbin(q);
                    const uint32_t base = walk.bin(cfg, q) * g.CPB + (row_lo + ri) * g.NZc;
                    lo = __ldg(a.cell_off + base + cell_z(g, zb, L));
                    len = __ldg(a.cell_off + base + cell_z(g, zb, U) + 1u) - lo;
                    wphi = walk.wphi(q);
                }
                const uint32_t incl = warp_incl_scan(len, lane);
                const uint32_t excl = incl - len;
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                if (!SPILL && lane == 0) visited += total;
                for (uint32_t p0 = 0; p0 < total; p0 += 32) {
                    const uint32_t p = p0 + lane;
                    const uint32_t o = owner_lane(incl, p);
                    const uint32_t c = __shfl_sync(0xffffffffu, lo, o) +
                                       (p - __shfl_sync(0xffffffffu, excl, o));
                    const uint32_t wp = __shfl_sync(0xffffffffu, wphi, o);
                    int st = 0;
                    float4 P = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p < total) {
                        P = __ldg(a.csp4 + c);
                        st = doublet_stage1(cfg, M.w, M.z, P.w, P.z);
                        if (st != 0) {
                            // division-free pre-decision; the exact reference chain only for
                            // the few pairs inside its uncertainty band
                            int d = bounded ? doublet_stage2_fast_bounded(cfg, M.x, M.y, P.x, P.y)
                                            : doublet_stage2_fast(cfg, M.x, M.y, P.x, P.y);
                            if (d == 2) d = doublet_stage2(cfg, M.x, M.y, P.x, P.y) ? 1 : 0;
                            if (d == 0) st = 0;
                        }
                    }
                    const uint32_t mB = __ballot_sync(0xffffffffu, st == 1);
                    const uint32_t mT = __ballot_sync(0xffffffffu, st == 2);
                    if (st != 0) {
                        const bool top = (st == 2);
                        const uint32_t k = top ? (wT + __popc(mT & ltmask))
                                               : (wB + __popc(mB & ltmask));
                        if (!direct) {
                            if (top) {
                                if (k < a.cap_t) {
                                    stage_t[k] = c;
                                    key_s[k] = wp;
                                }
                            } else if (k < a.cap_b) {
                                stage_b[k] = c;
                            }
                        } else {
                            const uint32_t pos = __ldg(a.ccanon + c);
                            const float2 V = has_var ? __ldg(a.var2 + pos) : make_float2(0.f, 0.f);
                            const LinCircle l = transform_coordinates_cs(!top, cosM, sinM, M.x, M.y, M.z, M.w, VM.x, VM.y, P.x, P.y, P.z, V.x, V.y);
                            DoubletRec r;
                            r.a = make_float4(l.cotTheta, l.iDeltaR, l.Er, l.U);
                            r.b = make_float4(l.V,
                                              top ? __uint_as_float(canon_key(wp, n_valid, pos))
                                                  : l.Zo,
                                              P.w, __uint_as_float(pos));
                            // unsorted tops go to the second half of a 2*nT allocation
                            (top ? a.arena_t + offT + nT : a.arena_b + offB)[k] = r;
                        }
                    }
                    wB += __popc(mB);
                    wT += __popc(mT);
                }
            }
            }
            if (direct) {
                // Sort the mid-top records by (cotTheta, reference order): bucket sort between
                // the two halves of the allocation. U = unsorted records (second half);
                // T = first half = final place.
                DoubletRec* U = a.arena_t + offT + nT;
                DoubletRec* T = a.arena_t + offT;
                uint32_t* bcnt = key_s;                                  // [nbk] count -> cursor
                uint32_t* bstart = reinterpret_cast<uint32_t*>(cot_s);   // [nbk]
                __syncwarp();
                float cmin = __uint_as_float(0x7f800000u), cmax = __uint_as_float(0xff800000u);
                for (uint32_t k = lane; k < nT; k += 32) {
                    const float c = U[k].a.x;
                    cmin = fminf(cmin, c);
                    cmax = fmaxf(cmax, c);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
                    cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
                }
                uint32_t nbk = nT / 16u + 1u;  // about 16 records per bucket
                if (nbk > a.cap_t) nbk = a.cap_t;
                float scale = float(nbk) / (cmax - cmin);
                if (!(scale > 0.f) || !(scale < 1e30f)) {  // all equal / non-finite: one bucket
                    nbk = 1;
                    scale = 0.f;
                }
                auto bucket_of = [&](float c) -> uint32_t {
                    const float t = (c - cmin) * scale;  // monotone in c
                    if (!(t > 0.f)) return 0u;
                    return (t >= float(nbk)) ? nbk - 1u : uint32_t(t);
                };
                for (uint32_t b = lane; b < nbk; b += 32) bcnt[b] = 0u;
                __syncwarp();
                for (uint32_t k = lane; k < nT; k += 32) atomicAdd(&bcnt[bucket_of(U[k].a.x)], 1u);
                __syncwarp();
                uint32_t carry = 0;
                for (uint32_t b0 = 0; b0 < nbk; b0 += 32) {
                    const uint32_t b = b0 + lane;
                    const uint32_t v = (b < nbk) ? bcnt[b] : 0u;
                    const uint32_t incl = warp_incl_scan(v, lane);
                    if (b < nbk) {
                        bstart[b] = carry + incl - v;
                        bcnt[b] = carry + incl - v;  // cursor
                    }
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                __syncwarp();
                for (uint32_t k = lane; k < nT; k += 32) {
                    const DoubletRec r = U[k];
                    T[atomicAdd(&bcnt[bucket_of(r.a.x)], 1u)] = r;
                }
                __syncwarp();
                for (uint32_t k = lane; k < nT; k += 32) U[k] = T[k];  // bucketed copy
                __syncwarp();
                // rank inside the bucket (after the scatter bcnt[b] is the end of bucket b)
                for (uint32_t k = lane; k < nT; k += 32) {
                    const DoubletRec r = U[k];
                    const uint32_t b = bucket_of(r.a.x);
                    const uint32_t s0 = bstart[b], s1 = bcnt[b];
                    const uint32_t kk = __float_as_uint(r.b.y);
                    uint32_t rank = 0;
                    for (uint32_t j = s0; j < s1; ++j) {
                        const float cj = U[j].a.x;
                        rank += ((cj < r.a.x) || (cj == r.a.x && __float_as_uint(U[j].b.y) < kk)) ? 1u : 0u;
                    }
                    T[s0 + rank] = r;
                }
                break;
            }
            nB = wB;
            nT = wT;
            // A middle continues only with >= 1 bottom and >= 1 top (seed_finding.cpp:85-95)
            if (nB == 0 || nT == 0) {
                nB = nT = 0;
                break;
            }
            if (nB > a.cap_b || nT > a.cap_t) {
                // does not fit the staging area: leave the counts for k_doublets<true>
                if (lane == 0) {
                    a.spill_list[atomicAdd(&a.ctrl->n_spill, 1u)] = m;
                    a.cnt_b[m] = nB;
                    a.cnt_t[m] = nT;
                }
                nB = nT = 0;
                deferred = true;
                break;
            }
            if (!allocate(nT)) break;
            __syncwarp();
            // common path: lin_circle of every staged doublet, full lanes
            for (uint32_t k = lane; k < nB; k += 32) {
                const uint32_t c = stage_b[k];
                const float4 P = __ldg(a.csp4 + c);
                const uint32_t pos = __ldg(a.ccanon + c);
                const float2 V = has_var ? __ldg(a.var2 + pos) : make_float2(0.f, 0.f);
                const LinCircle l = transform_coordinates_cs(true, cosM, sinM, M.x, M.y, M.z, M.w, VM.x, VM.y, P.x,
                                                          P.y, P.z, V.x, V.y);
                DoubletRec r;
                r.a = make_float4(l.cotTheta, l.iDeltaR, l.Er, l.U);
                r.b = make_float4(l.V, l.Zo, P.w, __uint_as_float(pos));
                a.arena_b[offB + k] = r;
            }
            // mid-tops: lin_circle once, the sort keys go to shared memory, the record waits
            // in registers (the first 32 per lane-slot) for its two ranks
            for (uint32_t k0 = 0; k0 < nT; k0 += 32) {
                const uint32_t k = k0 + lane;
                if (k < nT) {
                    const uint32_t c = stage_t[k];
                    const float4 P = __ldg(a.csp4 + c);
                    const uint32_t pos = __ldg(a.ccanon + c);
                    const float2 V = has_var ? __ldg(a.var2 + pos) : make_float2(0.f, 0.f);
                    cot_s[k] = transform_coordinates_cs(false, cosM, sinM, M.x, M.y, M.z, M.w, VM.x, VM.y, P.x,
                                                     P.y, P.z, V.x, V.y).cotTheta;
                    key_s[k] = canon_key(key_s[k], n_valid, pos);
                }
            }
            // pad to a multiple of four for the vectorised rank loop
            if (lane < 4u && nT + lane < a.cap_t) cot_s[nT + lane] = __uint_as_float(0x7f800000u);
            __syncwarp();
            // sorted position of every mid-top = number of smaller cotTheta values; equal values
            // (which would share a position) are detected through the sum of the positions and
            // resolved by the reference order. stage_b is free again: it holds the positions.
            uint32_t* rank_s = stage_b;
            uint32_t rank_sum = 0;
            const uint32_t n4 = (nT + 3u) & ~3u;
            const bool vec_ok = n4 <= a.cap_t;
            for (uint32_t k0 = 0; k0 < nT; k0 += 32) {
                const uint32_t k = k0 + lane;
                if (k < nT) {
                    const float ck = cot_s[k];
                    uint32_t lt = 0;
                    if (vec_ok) {
                        const float4* c4 = reinterpret_cast<const float4*>(cot_s);
                        for (uint32_t j = 0; j < n4 / 4u; ++j) {
                            const float4 v = c4[j];
                            lt += (v.x < ck) ? 1u : 0u;
                            lt += (v.y < ck) ? 1u : 0u;
                            lt += (v.z < ck) ? 1u : 0u;
                            lt += (v.w < ck) ? 1u : 0u;
                        }
                    } else {
                        for (uint32_t j = 0; j < nT; ++j) lt += (cot_s[j] < ck) ? 1u : 0u;
                    }
                    rank_s[k] = lt;
                    rank_sum += lt;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rank_sum += __shfl_xor_sync(0xffffffffu, rank_sum, o);
            if (rank_sum != (nT * (nT - 1u)) / 2u) {  // equal cotTheta values (or NaN): full order
                for (uint32_t k = lane; k < nT; k += 32)
                    rank_s[k] = top_rank([&](uint32_t j) { return cot_s[j]; },
                                         [&](uint32_t j) { return key_s[j]; }, nT, cot_s[k], key_s[k]);
            }
            __syncwarp();
            for (uint32_t k0 = 0; k0 < nT; k0 += 32) {
                const uint32_t k = k0 + lane;
                if (k < nT) {
                    const uint32_t c = stage_t[k];
                    const float4 P = __ldg(a.csp4 + c);
                    const uint32_t pos = __ldg(a.ccanon + c);
                    const float2 V = has_var ? __ldg(a.var2 + pos) : make_float2(0.f, 0.f);
                    const LinCircle l = transform_coordinates_cs(false, cosM, sinM, M.x, M.y, M.z, M.w, VM.x, VM.y,
                                                              P.x, P.y, P.z, V.x, V.y);
                    const uint32_t ks = rank_s[k];
                    DoubletRec r;
                    r.a = make_float4(cot_s[k], l.iDeltaR, l.Er, l.U);
                    r.b = make_float4(l.V, __uint_as_float(key_s[k]), P.w, __uint_as_float(pos));
                    a.arena_t[offT + ks] = r;
                }
            }
            break;
        }
        __syncwarp();
        if (lane == 0 && !deferred) {
            a.cnt_b[m] = nB;
            a.cnt_t[m] = nT;
            a.off_b[m] = offB;
            a.off_t[m] = offT;
            // Work list of k_triplets, longest jobs first: the middles with many
            // (mid-bottom, mid-top) combinations are handed out before the light ones, so the
            // launch does not end on a few warps that drew a heavy middle last.
            if (nB == 0u)
                a.seed_cnt[m] = 0u;
            else
                work_push(a.active_list, a.n_sp, a.ctrl, m, nB, nT);
        }
#ifdef B200_TAIL_PROBE
        if (MODE == 0 && lane == 0 && m < (1u << 19)) g_middle_cycles[m] = uint32_t(clock64() - probe_c0);
#endif
        if (nB) {
            ++acc_active;
            acc_nb += nB;
            acc_nt += nT;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    if (lane == 0) {
        atomicAdd(&s_pairs[0], pairs);
        atomicAdd(&s_pairs[1], visited);
        atomicAdd(&s_acc[0], acc_active);
        atomicAdd(&s_acc[1], acc_nb);
        atomicAdd(&s_acc[2], acc_nt);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_pairs[0]) atomicAdd(&a.ctrl->pair_tests, s_pairs[0]);
        if (s_pairs[1]) atomicAdd(&a.ctrl->pair_visited, s_pairs[1]);
        if (s_acc[0]) {
            atomicAdd(&a.ctrl->n_active, s_acc[0]);
            atomicAdd(&a.ctrl->n_mid_bot, s_acc[1]);
            atomicAdd(&a.ctrl->n_mid_top, s_acc[2]);
        }
    }
    // This launch started before k_doublets<0> had finished (programmatic stream serialization) and
    // needs nothing from it — but the launches behind it do, and they only wait for this one: it
    // must not complete first (PTX: a grid launched as a dependent has to execute
    // griddepcontrol.wait for the stream order to hold).
    // (one CTA is enough to keep the grid from completing; the others give their slots back)
    if (SIDES && blockIdx.x == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------
// (3)+(4) triplets, compatible-seed bonus, per-middle top-N
// ---------------------------------------------------------------------------
struct TripletArgs {
    const float4* sp4;
    const float2* var2;
    const uint32_t* sorted_bin;
    const uint32_t* cnt_b;
    const uint32_t* cnt_t;
    const uint32_t* off_b;
    const uint32_t* off_t;
    const DoubletRec* arena_b;
    const DoubletRec* arena_t;
    Control* ctrl;
    uint32_t* seed_cnt;     // [n_sp]
    uint32_t* seed_b;       // [n_sp * K] sorted position of the bottom spacepoint
    uint32_t* seed_t;       // [n_sp * K]
    float* seed_w;          // [n_sp * K]
    TripletDumpRec* dump;   // optional
    uint32_t max_dump;
    uint32_t list_cap;      // triplets of one 32-row block kept in shared memory
    const uint32_t* active_list;  // work list written by k_doublets (heavy first)
    uint32_t n_sp;
    uint32_t heavy_only;          // != 0: k_triplets takes the first `heavy_only` work classes only
                                  // (WORK_HEAVY_CLASSES: the rest is k_triplets_pool's; LANES_FIRST_CLASS:
                                  // the rest is k_triplets_lanes')
    DoubletRec* scratch_t;        // the mid-top arena again, writable: its unused tail holds the
                                  // triplets of a row that outgrows the shared-memory list
    uint32_t max_doublets;
    uint32_t* slow_list;          // [2 * n_sp] (middle, first row k_triplets did not finish)
};

// One triplet of the current row block (shared memory, 16 bytes).
struct __align__(16) BlockTriplet {
    uint32_t key;     // canon_key of the top spacepoint, later the position of the bottom one
    float curvature;  // later: radius of the bottom spacepoint
    float weight;     // -impact * impactWeightFactor, later the final weight
    float rT;         // radius of the top spacepoint, later the sorter sum
};

constexpr uint32_t TCOT_CAP = 256;  // cotTheta of the first mid-tops of a middle kept in smem
constexpr uint32_t HANDED_OVER = 0xFFFFFFF0u;  // row cursor of a middle given to the slow path

// cotTheta of mid-top t: shared-memory copy for the first TCOT_CAP, the arena otherwise
struct TopCot {
    const float* sm;
    const DoubletRec* LT;
    __device__ __forceinline__ float operator()(uint32_t t) const {
        return (t < TCOT_CAP) ? sm[t] : __ldg(&LT[t].a.x);
    }
};
// first t in [0, n) with cot(t) >= v  (mid-top records are sorted by cotTheta)
__device__ __forceinline__ uint32_t cot_lower_bound(const TopCot& cot, uint32_t n, float v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cot(mid) < v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
// first t in [from, n) with cot(t) > v; gallops from `from` (the windows are short)
__device__ __forceinline__ uint32_t cot_upper_bound(const TopCot& cot, uint32_t from, uint32_t n,
                                                    float v) {
    uint32_t lo = from, hi = n, step = 2;
    while (lo + step < n) {
        if (cot(lo + step) > v) {
            hi = lo + step;
            break;
        }
        lo += step + 1;
        step <<= 1;
    }
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cot(mid) <= v)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// The same two searches over a list that lies completely in shared memory (n >= 1), without
// data-dependent branches: halve the interval log2(n) times.
__device__ __forceinline__ uint32_t smem_lower_bound(const float* sm, uint32_t n, float v) {
    uint32_t base = 0;
    while (n > 1u) {
        const uint32_t half = n >> 1;
        base += (sm[base + half - 1u] < v) ? half : 0u;
        n -= half;
    }
    return base + ((sm[base] < v) ? 1u : 0u);
}
__device__ __forceinline__ uint32_t smem_upper_bound(const float* sm, uint32_t n, float v) {
    uint32_t base = 0;
    while (n > 1u) {
        const uint32_t half = n >> 1;
        base += (sm[base + half - 1u] <= v) ? half : 0u;
        n -= half;
    }
    return base + ((sm[base] <= v) ? 1u : 0u);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Per-warp shared memory of k_triplets.
__host__ __device__ inline size_t triplet_smem_per_warp(uint32_t list_cap, bool dense) {
    // list (16) + pos of top (4) + mid-bottom index (4) + order (2) + bonus / rank (1) per entry
    // + 64 pending (row, mid-top) pairs of k_triplets<DENSE>'s pre-filter
    return size_t(list_cap) * (16 + 4 + 4 + 2 + 1) + MAX_TOPK * 5 * 4 + TCOT_CAP * 4 +
           (dense ? 64 * 4 : 0);
}

// One triplet of a mid-bottom row that outgrows the shared-memory list: a 32-byte slot in the unused
// tail of the mid-top arena (global memory).
struct __align__(16) BigRowTriplet {
    uint32_t key;     // canon_key of the top spacepoint, later the position of the bottom one
    float curvature;  // later: radius of the bottom spacepoint
    float weight;     // -impact * impactWeightFactor, later the final weight
    float rT;         // radius of the top spacepoint, later the sorter sum
    uint32_t pos_t;
    uint32_t ord;     // slot q holds the list index of the q-th triplet in the reference's order
    uint32_t aux;     // compatible-seed count, later the rank in the top-K (0xFF: none)
    uint32_t pad_;
};
static_assert(sizeof(BigRowTriplet) == sizeof(DoubletRec), "one arena slot per triplet");

// The slow path of the triplet search: a middle one of whose mid-bottom doublets produced more
// accepted triplets than the shared-memory list of k_triplets holds. k_triplets hands such a
// middle over (with the row it stopped at); this kernel redoes it from the first row, row by
// row, with the same steps as k_triplets' flush — exact cuts against ALL mid-tops of the middle (the
// cotTheta window is only a pruning), reference order inside the row, compatible-seed bonus, final
// weight / single-seed cut, merge into the middle's top-K — with the row's triplets in global memory
// (nt slots from the unused tail of the mid-top arena, allocated once per middle), then makes the
// final per-middle selection. It runs inside k_seed_gather (tile 0 does it before any tile reads the
// per-middle seed counts; one load of n_slow per CTA otherwise), because everywhere else it cost
// the common path: as a call inside k_triplets' row loop +36 %, after that loop +13 %, from the last
// CTA of k_triplets +4 % (64 registers instead of 62, spills around the call), as its own empty
// launch -2 % throughput. Warp per handed-over middle.
__device__ __noinline__ void triplets_slow_middles(const DevCfg& cfg, const TripletArgs& a,
                                                   uint32_t* smem /* >= 5 * MAX_TOPK words per warp */) {
    const uint32_t n_slow = *reinterpret_cast<volatile uint32_t*>(&a.ctrl->n_slow);
    const uint32_t warp = threadIdx.x >> 5;
    // the kernel's dynamic shared memory is free by now: the per-warp top-K lives there
    float* top_w = reinterpret_cast<float*>(smem + size_t(warp) * 5 * MAX_TOPK);
    float* top_s = top_w + MAX_TOPK;
    float* top_rb = top_s + MAX_TOPK;
    uint32_t* top_b = reinterpret_cast<uint32_t*>(top_rb + MAX_TOPK);
    uint32_t* top_t = top_b + MAX_TOPK;
    const uint32_t n_valid = a.ctrl->n_valid;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t ltmask = lanemask_lt();
    const uint32_t K = cfg.maxSeedsPerSpM;
    uint32_t found = 0;
  while (true) {
    uint32_t tk = 0;
    if (lane == 0) tk = atomicAdd(&a.ctrl->ticket_q, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= n_slow) break;
    const uint32_t m = a.slow_list[2 * tk], row_begin = a.slow_list[2 * tk + 1];
    const uint32_t nb = a.cnt_b[m], nt = a.cnt_t[m];
    const DoubletRec* LB = a.arena_b + a.off_b[m];
    const DoubletRec* LT = a.arena_t + a.off_t[m];
    const float4 M = __ldg(a.sp4 + m);
    const float2 VM = __ldg(a.var2 + m);
    const float rM = M.w, varZM = VM.x, varRM = VM.y;
    const uint32_t walk_r0 = circular_remap(cfg.nPhi, __ldg(a.sorted_bin + m) % cfg.nPhi, -int(cfg.scope0));
    uint32_t ntop = 0;  // the whole middle is redone: nothing of k_triplets' top-K was kept
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&a.ctrl->cursor[1], nt);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base > a.max_doublets || nt > a.max_doublets - base) {
        // no room left in the arena either: the rest of the middle is lost and the event is flagged
        // (an error for the caller, B200SEED_EOVERFLOW); what was found so far is kept
        if (lane == 0) atomicOr(&a.ctrl->overflow, B200SEED_OVF_TRIPLETS);
    } else {
    BigRowTriplet* S = reinterpret_cast<BigRowTriplet*>(a.scratch_t + base);
    auto tie_key = [&](uint32_t pos) -> unsigned long long {
        const uint32_t pb = __ldg(a.sorted_bin + pos) % cfg.nPhi;
        const uint32_t w = (pb + cfg.nPhi - walk_r0) % cfg.nPhi;
        return (unsigned long long)w * n_valid + pos;
    };
    auto before = [&](float w1, float s1, uint32_t b1, uint32_t t1, float w2, float s2, uint32_t b2,
                      uint32_t t2) -> bool {
        if (w1 != w2 || s1 != s2) return seed_before(w1, s1, w2, s2);
        const unsigned long long k1 = tie_key(b1), k2 = tie_key(b2);
        return (k1 != k2) ? (k1 < k2) : (tie_key(t1) < tie_key(t2));
    };
    for (uint32_t row = 0; row < nb; ++row) {
        // (0) the exact cuts against every mid-top
        const float4 ba = __ldg(&LB[row].a);
        const float4 bb = __ldg(&LB[row].b);
        LinCircle lb;
        lb.cotTheta = ba.x, lb.iDeltaR = ba.y, lb.Er = ba.z, lb.U = ba.w;
        lb.V = bb.x, lb.Zo = bb.y;
        float is2, s2;
        triplet_row_constants(cfg, lb.cotTheta, is2, s2);
        uint32_t n = 0;
        for (uint32_t p0 = 0; p0 < nt; p0 += 32) {
            const uint32_t tt = p0 + lane;
            bool ok = false;
            BigRowTriplet e{};
            if (tt < nt) {
                const float4 ta = __ldg(&LT[tt].a);
                const float4 tb = __ldg(&LT[tt].b);
                LinCircle lt;
                lt.cotTheta = ta.x, lt.iDeltaR = ta.y, lt.Er = ta.z, lt.U = ta.w;
                lt.V = tb.x, lt.Zo = 0.f;
                float curvature = 0.f, impact = 0.f;
                ok = triplet_is_compatible(cfg, rM, varRM, varZM, lb, lt, is2, s2, curvature, impact);
                e.key = __float_as_uint(tb.y);
                e.curvature = curvature;
                e.weight = -impact * cfg.impactWeightFactor;
                e.rT = tb.z;
                e.pos_t = __float_as_uint(tb.w);
            }
            const uint32_t mk = __ballot_sync(0xffffffffu, ok);
            if (ok) S[n + __popc(mk & ltmask)] = e;
            n += __popc(mk);
        }
        __syncwarp();
        if (n == 0) continue;
        // (1) reference order inside the row: slot rank(i) holds i (the keys are unique)
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t key = S[i].key;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n; ++j) rank += (S[j].key < key) ? 1u : 0u;
            S[rank].ord = i;
        }
        __syncwarp();
        // (2) compatible-seed bonus (triplet_finding.hpp:107-179)
        for (uint32_t i = lane; i < n; i += 32) {
            const float c_rT = S[i].rT, c_curv = S[i].curvature;
            const float lower = c_curv - cfg.deltaInvHelixDiameter;
            const float upper = c_curv + cfg.deltaInvHelixDiameter;
            float compat[MAX_COMPAT];
            uint32_t ncompat = 0;
            for (uint32_t q = 0; q < n; ++q) {
                const uint32_t j = S[q].ord;
                if (j == i) continue;
                const float o_rT = S[j].rT, o_curv = S[j].curvature;
                const float deltaR = c_rT - o_rT;
                if (absf(deltaR) < cfg.filterDeltaRMin) continue;
                if (o_curv < lower) continue;
                if (o_curv > upper) continue;
                bool newCompSeed = true;
#pragma unroll
                for (uint32_t c = 0; c < MAX_COMPAT; ++c)
                    if (c < ncompat && absf(compat[c] - o_rT) < cfg.filterDeltaRMin) newCompSeed = false;
                if (newCompSeed) {
#pragma unroll
                    for (uint32_t c = 0; c < MAX_COMPAT; ++c)
                        if (c == ncompat) compat[c] = o_rT;
                    ++ncompat;
                }
                if (ncompat >= cfg.compatSeedLimit) break;
            }
            S[i].aux = ncompat;
        }
        __syncwarp();
        // (3) final weight, single-seed cut, sorter sum; optional dump
        const uint32_t pos_b = __float_as_uint(bb.w);
        const float rB = bb.z;
        const float4 PB = __ldg(a.sp4 + pos_b);
        for (uint32_t i = lane; i < n; i += 32) {
            BigRowTriplet cur = S[i];
            float w = cur.weight;  // the reference adds compatSeedWeight one at a time (:171)
            for (uint32_t q = cur.aux; q > 0; --q) w += cfg.compatSeedWeight;
            if (a.dump && row >= row_begin) {  // earlier rows were dumped by k_triplets
                const uint32_t d = atomicAdd(&a.ctrl->dump_cursor, 1u);
                if (d < a.max_dump) {
                    TripletDumpRec r;
                    r.pos_b = pos_b, r.pos_m = m, r.pos_t = cur.pos_t, r.mb_idx = row;
                    r.mt_idx = cur.key, r.curvature = cur.curvature, r.weight = w;
                    r.z_vertex = bb.y;
                    a.dump[d] = r;
                } else {
                    atomicOr(&a.ctrl->overflow, B200SEED_OVF_DUMP);
                }
            }
            w += seed_weight_increase(cfg, rB, cur.rT);
            const bool keep = single_seed_cut(cfg, rB, w);
            const float4 PT = __ldg(a.sp4 + cur.pos_t);
            cur.weight = w;
            cur.rT = sorter_sum(PB.y, PB.z, PT.y, PT.z);
            cur.curvature = rB;
            cur.key = keep ? pos_b : 0xFFFFFFFFu;
            cur.aux = 0xFFu;
            S[i] = cur;
        }
        __syncwarp();
        // (4) merge into the per-middle top-K (triplet_sorter's order; full ties: order of discovery)
        uint32_t nkept = 0;
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t rank = 0xFFu;
            if (i < n) {
                const BigRowTriplet c = S[i];
                bool in = (c.key != 0xFFFFFFFFu);
                if (in && ntop == K)
                    in = before(c.weight, c.rT, c.key, c.pos_t, top_w[K - 1], top_s[K - 1], top_b[K - 1],
                                top_t[K - 1]);
                if (in) {
                    rank = 0;
                    for (uint32_t q = 0; q < ntop; ++q)
                        rank += before(top_w[q], top_s[q], top_b[q], top_t[q], c.weight, c.rT, c.key, c.pos_t)
                                    ? 1u : 0u;
                    for (uint32_t j = 0; j < n && rank < K; ++j) {
                        if (j == i) continue;
                        const uint32_t o_key = S[j].key;
                        if (o_key != 0xFFFFFFFFu &&
                            before(S[j].weight, S[j].rT, o_key, S[j].pos_t, c.weight, c.rT, c.key, c.pos_t))
                            ++rank;
                    }
                    if (rank >= K) rank = 0xFFu;
                }
            }
            __syncwarp();                       // everybody has read the list entries it ranks against
            if (i < n) S[i].aux = rank;
            nkept += __popc(__ballot_sync(0xffffffffu, rank != 0xFFu));
        }
        __syncwarp();
        if (nkept) {
            float ew = 0.f, es = 0.f, erb = 0.f;
            uint32_t eb = 0, et = 0, npos = 0xFFu;
            if (lane < ntop) {
                ew = top_w[lane], es = top_s[lane], erb = top_rb[lane];
                eb = top_b[lane], et = top_t[lane];
                npos = lane;
                for (uint32_t j = 0; j < n; ++j) {
                    if (S[j].aux == 0xFFu) continue;
                    npos += before(S[j].weight, S[j].rT, S[j].key, S[j].pos_t, ew, es, eb, et) ? 1u : 0u;
                }
            }
            __syncwarp();
            if (npos < K) {
                top_w[npos] = ew, top_s[npos] = es, top_rb[npos] = erb;
                top_b[npos] = eb, top_t[npos] = et;
            }
            for (uint32_t i = lane; i < n; i += 32) {
                const BigRowTriplet c = S[i];
                if (c.aux != 0xFFu) {
                    top_w[c.aux] = c.weight, top_s[c.aux] = c.rT, top_rb[c.aux] = c.curvature;
                    top_b[c.aux] = c.key, top_t[c.aux] = c.pos_t;
                }
            }
            ntop = (ntop + nkept < K) ? (ntop + nkept) : K;
        }
        __syncwarp();
        if (row >= row_begin) found += n;  // earlier rows were counted by k_triplets
    }
    }  // arena room
    // ---- final per-middle selection (seed_filtering.cpp:84-122) ----
    {
        const bool keep = lane < ntop && (lane == 0 || cut_per_middle_sp(cfg, top_rb[lane], top_w[lane]));
        const uint32_t km = __ballot_sync(0xffffffffu, keep);
        const float w_ = keep ? top_w[lane] : 0.f;
        const uint32_t b_ = keep ? top_b[lane] : 0u, t_ = keep ? top_t[lane] : 0u;
        __syncwarp();
        if (keep) {
            const uint32_t o = __popc(km & ltmask);
            a.seed_b[size_t(m) * K + o] = b_;
            a.seed_t[size_t(m) * K + o] = t_;
            a.seed_w[size_t(m) * K + o] = w_;
        }
        if (lane == 0) a.seed_cnt[m] = __popc(km);
    }
    __syncwarp();
  }
    if (lane == 0 && found) atomicAdd(&a.ctrl->n_triplets, found);
}

// Warp per middle spacepoint (atomic ticket queue). For every block of 32 mid-bottom doublets
// (one per lane) the lane binary-searches, in the cotTheta-sorted mid-top list, the window
// outside of which the first scattering cut of triplet_finding_helper::isCompatible (:57-78)
// must fail; the (row, top) pairs inside the windows are flattened with a warp scan and
// evaluated 32 at a time at full lane occupancy with the exact reference arithmetic.
// The window is conservative (see the margin below), so the accepted set is identical to
// testing all nMidBot x nMidTop combinations.
#ifndef B200_TRIPLET_MIN_CTAS
#define B200_TRIPLET_MIN_CTAS (32 / B200_WARPS_PER_CTA)
#endif
// DENSE selects the top-K merge written for busy events (see the flush, step 4): the host uses
// it above 80k spacepoints; the plain variant is 3 % faster on the 10k-particle event.
template <bool DENSE>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, B200_TRIPLET_MIN_CTAS)
k_triplets(const DevCfg cfg, const TripletArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ uint32_t s_ntrip;
    __shared__ unsigned long long s_tests, s_visited;
    __shared__ uint32_t s_pre[WORK_CLASSES + 1];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltmask = lanemask_lt();
    unsigned char* base = s_raw + triplet_smem_per_warp(a.list_cap, DENSE) * warp;
    BlockTriplet* list = reinterpret_cast<BlockTriplet*>(base);
    uint32_t* lpos = reinterpret_cast<uint32_t*>(base + size_t(a.list_cap) * 16);  // pos of top
    uint32_t* lrow = lpos + a.list_cap;  // index of the mid-bottom doublet
    float* top_w = reinterpret_cast<float*>(lrow + a.list_cap);
    float* top_s = top_w + MAX_TOPK;
    float* top_rb = top_s + MAX_TOPK;
    uint32_t* top_b = reinterpret_cast<uint32_t*>(top_rb + MAX_TOPK);
    uint32_t* top_t = top_b + MAX_TOPK;
    float* cot_sm = reinterpret_cast<float*>(top_t + MAX_TOPK);
    uint32_t* pend = reinterpret_cast<uint32_t*>(cot_sm + TCOT_CAP);  // [64] (row << 27) | top
    uint16_t* ord = reinterpret_cast<uint16_t*>(pend + (DENSE ? 64 : 0));  // list order inside a row
    uint8_t* aux = reinterpret_cast<uint8_t*>(ord + a.list_cap);     // bonus count, later rank
    if (threadIdx.x == 0) {
        s_ntrip = 0;
        s_tests = s_visited = 0ull;
    }
    work_prefix(a.ctrl->n_cls, s_pre);
    // (k_triplets_lanes, launched behind this kernel as a programmatic dependent, shares no data with it)
    asm volatile("griddepcontrol.launch_dependents;");
    __syncthreads();
    const uint32_t n_valid = a.ctrl->n_valid;
    const uint32_t K = cfg.maxSeedsPerSpM;
    uint32_t acc_trip = 0;
    unsigned long long acc_tests = 0ull, acc_visited = 0ull;

    // Work items: the active middles k_doublets listed, longest jobs first (see work_class),
    // middles without work never drawn.
    const uint32_t n_work = s_pre[a.heavy_only ? a.heavy_only : WORK_CLASSES];
    while (true) {
        uint32_t m = 0;
        if (lane == 0) m = atomicAdd(&a.ctrl->ticket, 1u);
        m = __shfl_sync(0xffffffffu, m, 0);
#ifdef B200_TAIL_PROBE
        if (!DENSE && lane == 0) {
            const uint32_t w = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
            if (w < 16384u) {
                if (g_tail_probe[1][0][w] == 0ull) g_tail_probe[1][0][w] = probe_now();
                if (m >= n_work) g_tail_probe[1][1][w] = probe_now();
            }
        }
#endif
        if (m >= n_work) break;
        m = work_item(a.active_list, a.n_sp, s_pre, m);
        const uint32_t nb = a.cnt_b[m], nt = a.cnt_t[m];
#ifdef B200_TAIL_PROBE
        if (!DENSE && lane == 0) {
            const uint32_t w = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
            if (w < 16384u) {
                g_tail_probe[1][2][w] = probe_now();
                g_tail_probe[1][3][w] = ((unsigned long long)nb << 32) | nt;
            }
        }
#endif
        acc_tests += (unsigned long long)nb * nt;
        const DoubletRec* LB = a.arena_b + a.off_b[m];
        const DoubletRec* LT = a.arena_t + a.off_t[m];
        const float4 M = __ldg(a.sp4 + m);
        const float2 VM = __ldg(a.var2 + m);  // {varZ, varR}
        const float rM = M.w, varZM = VM.x, varRM = VM.y;
        // order of a spacepoint among the doublet partners of this middle in the reference
        // (canon_key of k_doublets); only evaluated for full ties in the seed ranking
        const uint32_t walk_r0 =
            circular_remap(cfg.nPhi, __ldg(a.sorted_bin + m) % cfg.nPhi, -int(cfg.scope0));
        auto tie_key = [&](uint32_t pos) -> unsigned long long {
            const uint32_t pb = __ldg(a.sorted_bin + pos) % cfg.nPhi;
            const uint32_t w = (pb + cfg.nPhi - walk_r0) % cfg.nPhi;
            return (unsigned long long)w * n_valid + pos;
        };

        // bounds over the mid-tops for the conservative window
        float maxEr = 0.f, minEr = 0.f, maxIDR = 0.f, maxAbsCot = 0.f;
        for (uint32_t t = lane; t < nt; t += 32) {
            const float4 ta = __ldg(&LT[t].a);
            if (t < TCOT_CAP) cot_sm[t] = ta.x;
            maxEr = fmaxf(maxEr, ta.z);
            minEr = fminf(minEr, ta.z);
            maxIDR = fmaxf(maxIDR, ta.y);
            maxAbsCot = fmaxf(maxAbsCot, absf(ta.x));
        }
        maxEr = warp_max(maxEr);
        minEr = warp_min(minEr);
        maxIDR = warp_max(maxIDR);
        maxAbsCot = warp_max(maxAbsCot);
        __syncwarp();
        const TopCot top_cot{cot_sm, LT};
        // with negative or non-finite error terms the reference's sqrt() yields NaN and the
        // cut passes everything: no pruning then
        const bool sane = (varRM >= 0.f) && (varZM >= 0.f) && (minEr >= 0.f) && (maxEr < 1e30f) &&
                          (maxIDR < 1e30f) && (maxAbsCot < 1e30f);

        uint32_t ntop = 0;   // entries in the per-middle top-K (warp-uniform)
        uint32_t nlist = 0;  // triplets waiting in the shared-memory list (complete rows)
        uint32_t row0 = 0;
        uint32_t rows = 32;  // rows of the next block (shrinks only on list overflow)
        while (true) {
            // ---- windows of the next block of mid-bottom doublets (one row per lane) ----
            const bool have_block = row0 < nb;
            const uint32_t nrows = have_block ? ((nb - row0 < rows) ? (nb - row0) : rows) : 0u;
            const bool has_row = lane < nrows;
            uint32_t lo = 0, hi = 0;
            // the rows of the next block into L1 while this block is searched and evaluated (no
            // registers: the block's first use of them is then an L1 hit instead of a DRAM round
            // trip; 116.5 -> 115.2 us on the 10k-particle event)
            if (row0 + nrows + lane < nb)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(&LB[row0 + nrows + lane]));
            if (has_row) {
                const float4 la = __ldg(&LB[row0 + lane].a);
                float iSinTheta2, sir2;
                triplet_row_constants(cfg, la.x, iSinTheta2, sir2);
                // Window half-width. The cut rejects iff d2 - e2 > 0 and (|d| - e)^2 > sir2
                // (d = cot_b - cot_t, e^2 = error2 <= e2max), i.e. surely when
                // |d| > e + sqrt(sir2); the factors absorb float rounding including the
                // cancellation in d2 + e2 - 2|d|e (relative error <= 1e-6 (d^2 + e^2)).
                const float e2max = la.z + maxEr +
                                    2.f * (absf(la.x) * maxAbsCot * varRM + varZM) * la.y * maxIDR;
                const float W = 1.004f * sqrt_rn(e2max) + 1.002f * sqrt_rn(sir2) +
                                4e-6f * (absf(la.x) + maxAbsCot) + 1e-30f;
                const bool prune = sane && (la.z >= 0.f) && (W < 1e30f) && (sir2 >= 0.f);
                if (nt <= TCOT_CAP) {
                    // the whole list is in shared memory: branch-free searches whose trip count
                    // depends on nt only (the same for all rows of the middle: no divergence)
                    lo = prune ? smem_lower_bound(cot_sm, nt, la.x - W) : 0u;
                    hi = prune ? smem_upper_bound(cot_sm, nt, la.x + W) : nt;
                } else {
                    lo = prune ? cot_lower_bound(top_cot, nt, la.x - W) : 0u;
                    hi = prune ? cot_upper_bound(top_cot, lo, nt, la.x + W) : nt;
                }
                if (hi < lo) hi = lo;
            }
            const uint32_t wdt = hi - lo;
            const uint32_t incl = warp_incl_scan(wdt, lane);
            const uint32_t excl = incl - wdt;
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);

            // ---- flush the list when the middle is finished or the next block may not fit ----
            if (nlist != 0 && (!have_block || nlist + total > a.list_cap)) {
                __syncwarp();
                // (1) reference order inside every row: ord[first of row + rank by canon_key]
                for (uint32_t i0 = 0; i0 < nlist; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    if (i < nlist) {
                        const uint32_t row = lrow[i], key = list[i].key;
                        uint32_t sgm = i, rank = 0;
                        while (sgm > 0 && lrow[sgm - 1] == row) {
                            --sgm;
                            rank += (list[sgm].key < key) ? 1u : 0u;
                        }
                        for (uint32_t j = i + 1; j < nlist && lrow[j] == row; ++j)
                            rank += (list[j].key < key) ? 1u : 0u;
                        ord[sgm + rank] = uint16_t(i);
                    }
                }
                __syncwarp();
                // (2) compatible-seed bonus (triplet_finding.hpp:107-179), lane per triplet:
                //     the other triplets of the same mid-bottom doublet in the reference order
                for (uint32_t i0 = 0; i0 < nlist; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    if (i < nlist) {
                        const BlockTriplet cur = list[i];
                        const uint32_t row = lrow[i];
                        uint32_t sgm = i, egm = i + 1;
                        while (sgm > 0 && lrow[sgm - 1] == row) --sgm;
                        while (egm < nlist && lrow[egm] == row) ++egm;
                        const float lower = cur.curvature - cfg.deltaInvHelixDiameter;
                        const float upper = cur.curvature + cfg.deltaInvHelixDiameter;
                        float compat[MAX_COMPAT];
                        uint32_t ncompat = 0;
                        for (uint32_t q = sgm; q < egm; ++q) {
                            const uint32_t j = ord[q];
                            if (j == i) continue;
                            const BlockTriplet o = list[j];
                            const float deltaR = cur.rT - o.rT;
                            if (absf(deltaR) < cfg.filterDeltaRMin) continue;
                            if (o.curvature < lower) continue;
                            if (o.curvature > upper) continue;
                            bool newCompSeed = true;
#pragma unroll
                            for (uint32_t c = 0; c < MAX_COMPAT; ++c) {
                                if (c < ncompat && absf(compat[c] - o.rT) < cfg.filterDeltaRMin)
                                    newCompSeed = false;
                            }
                            if (newCompSeed) {
#pragma unroll
                                for (uint32_t c = 0; c < MAX_COMPAT; ++c)
                                    if (c == ncompat) compat[c] = o.rT;
                                ++ncompat;
                            }
                            if (ncompat >= cfg.compatSeedLimit) break;
                        }
                        aux[i] = uint8_t(ncompat);
                    }
                }
                __syncwarp();
                // (3) final weight, single-seed cut, sorter sum; optional dump
                for (uint32_t i0 = 0; i0 < nlist; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    if (i < nlist) {
                        BlockTriplet cur = list[i];
                        const uint32_t row = lrow[i];
                        // the reference adds compatSeedWeight one at a time (:171)
                        float w = cur.weight;
                        for (uint32_t q = aux[i]; q > 0; --q) w += cfg.compatSeedWeight;
                        const float4 bb = __ldg(&LB[row].b);
                        const uint32_t pos_b = __float_as_uint(bb.w), pos_t = lpos[i];
                        if (a.dump) {
                            const uint32_t d = atomicAdd(&a.ctrl->dump_cursor, 1u);
                            if (d < a.max_dump) {
                                TripletDumpRec r;
                                r.pos_b = pos_b, r.pos_m = m, r.pos_t = pos_t, r.mb_idx = row;
                                r.mt_idx = cur.key, r.curvature = cur.curvature, r.weight = w;
                                r.z_vertex = bb.y;
                                a.dump[d] = r;
                            } else {
                                atomicOr(&a.ctrl->overflow, B200SEED_OVF_DUMP);
                            }
                        }
                        const float rB = bb.z, rT = cur.rT;
                        w += seed_weight_increase(cfg, rB, rT);
                        const bool keep = single_seed_cut(cfg, rB, w);
                        const float4 PB = __ldg(a.sp4 + pos_b);
                        const float4 PT = __ldg(a.sp4 + pos_t);
                        cur.weight = w;
                        cur.rT = sorter_sum(PB.y, PB.z, PT.y, PT.z);
                        cur.curvature = rB;
                        cur.key = keep ? pos_b : 0xFFFFFFFFu;  // pos_b < 2^32 - 1 always
                        list[i] = cur;
                    }
                }
                __syncwarp();
                // (4) merge into the per-middle top-K: every kept triplet computes its position
                //     in the union of the list and the current top-K under triplet_sorter's
                //     order (full ties: the reference's order of discovery)
                auto before = [&](float w1, float s1, uint32_t b1, uint32_t t1, float w2, float s2,
                                  uint32_t b2, uint32_t t2) -> bool {
                    if (w1 != w2 || s1 != s2) return seed_before(w1, s1, w2, s2);
                    const unsigned long long k1 = tie_key(b1), k2 = tie_key(b2);
                    return (k1 != k2) ? (k1 < k2) : (tie_key(t1) < tie_key(t2));
                };
                uint32_t nkept = 0;
                if (DENSE) {
                    // Busy events (long lists, many flushes per middle). Candidates: kept
                    // triplets that sort before the current K-th entry (all kept ones while the
                    // top-K is not full). Everything else can neither enter the top-K nor sort
                    // before a candidate, so ranks are counted among candidates and current
                    // entries only. With a full top-K the candidates are few: they are compacted
                    // first (`ord` is free again) and ranked one per lane.
                    uint32_t ncand = 0;
                    uint16_t* cand = ord;
                    for (uint32_t i0 = 0; i0 < nlist; i0 += 32) {
                        const uint32_t i = i0 + lane;
                        bool in = false;
                        if (i < nlist) {
                            const BlockTriplet c = list[i];
                            in = (c.key != 0xFFFFFFFFu);
                            if (in && ntop == K)  // cannot displace anything otherwise
                                in = before(c.weight, c.rT, c.key, lpos[i], top_w[K - 1],
                                            top_s[K - 1], top_b[K - 1], top_t[K - 1]);
                            aux[i] = 0xFFu;
                        }
                        const uint32_t cm = __ballot_sync(0xffffffffu, in);
                        if (in) cand[ncand + __popc(cm & ltmask)] = uint16_t(i);
                        ncand += __popc(cm);
                    }
                    __syncwarp();
                    if (ncand > 32u) {
                        // Many candidates: K rounds of warp-wide selection instead of ranking
                        // every candidate against all others. Round r picks the best entry of
                        // (candidates + current top-K) that sorts after the winner of round r-1;
                        // lane r keeps it. Same result: `before` is a strict total order.
                        float pw = 0.f, ps = 0.f;        // previous winner
                        uint32_t pb = 0, pt = 0;
                        float mw = 0.f, ms = 0.f, mrb = 0.f;  // the entry this lane will hold
                        uint32_t mb = 0, mt = 0;
                        uint32_t nnew = 0;
                        for (uint32_t r = 0; r < K; ++r) {
                            bool have = false;
                            float bw = 0.f, bs = 0.f;
                            uint32_t bb = 0, bt = 0, bsrc = 0;
                            auto offer = [&](float w, float s_, uint32_t b, uint32_t t, uint32_t src) {
                                if (r != 0 && !before(pw, ps, pb, pt, w, s_, b, t)) return;
                                if (!have || before(w, s_, b, t, bw, bs, bb, bt)) {
                                    have = true;
                                    bw = w, bs = s_, bb = b, bt = t, bsrc = src;
                                }
                            };
                            for (uint32_t q = lane; q < ncand; q += 32) {
                                const uint32_t i = cand[q];
                                const BlockTriplet c = list[i];
                                offer(c.weight, c.rT, c.key, lpos[i], i);
                            }
                            if (lane < ntop)
                                offer(top_w[lane], top_s[lane], top_b[lane], top_t[lane], 0x8000u | lane);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                const bool oh = __shfl_xor_sync(0xffffffffu, have, o);
                                const float ow = __shfl_xor_sync(0xffffffffu, bw, o);
                                const float os = __shfl_xor_sync(0xffffffffu, bs, o);
                                const uint32_t ob = __shfl_xor_sync(0xffffffffu, bb, o);
                                const uint32_t ot = __shfl_xor_sync(0xffffffffu, bt, o);
                                const uint32_t osrc = __shfl_xor_sync(0xffffffffu, bsrc, o);
                                if (oh && (!have || before(ow, os, ob, ot, bw, bs, bb, bt))) {
                                    have = true;
                                    bw = ow, bs = os, bb = ob, bt = ot, bsrc = osrc;
                                }
                            }
                            if (!have) break;  // warp-uniform after the butterfly
                            pw = bw, ps = bs, pb = bb, pt = bt;
                            if (lane == r) {
                                mw = bw, ms = bs, mb = bb, mt = bt;
                                mrb = (bsrc & 0x8000u) ? top_rb[bsrc & 0x7FFFu] : list[bsrc].curvature;
                            }
                            ++nnew;
                        }
                        __syncwarp();
                        if (lane < nnew) {
                            top_w[lane] = mw, top_s[lane] = ms, top_rb[lane] = mrb;
                            top_b[lane] = mb, top_t[lane] = mt;
                        }
                        ntop = nnew;
                        ncand = 0;  // nothing left for the ranking path below
                    }
                    for (uint32_t q0 = 0; q0 < ncand; q0 += 32) {
                        const uint32_t qi = q0 + lane;
                        uint32_t rank = 0xFFu;
                        if (qi < ncand) {
                            const uint32_t i = cand[qi];
                            const BlockTriplet c = list[i];
                            const uint32_t ct = lpos[i];
                            rank = 0;
                            for (uint32_t q = 0; q < ntop; ++q)
                                rank += before(top_w[q], top_s[q], top_b[q], top_t[q], c.weight,
                                               c.rT, c.key, ct) ? 1u : 0u;
                            for (uint32_t jq = 0; jq < ncand && rank < K; ++jq) {
                                const uint32_t j = cand[jq];
                                const BlockTriplet o = list[j];
                                if (j != i && before(o.weight, o.rT, o.key, lpos[j], c.weight, c.rT,
                                                     c.key, ct))
                                    ++rank;
                            }
                            if (rank >= K) rank = 0xFFu;
                            aux[i] = uint8_t(rank);
                        }
                        nkept += __popc(__ballot_sync(0xffffffffu, rank != 0xFFu));
                    }
                    __syncwarp();
                    if (nkept) {
                        // current entries move down by the number of new entries sorted before them
                        float ew = 0.f, es = 0.f, erb = 0.f;
                        uint32_t eb = 0, et = 0, npos = 0xFFu;
                        if (lane < ntop) {
                            ew = top_w[lane], es = top_s[lane], erb = top_rb[lane];
                            eb = top_b[lane], et = top_t[lane];
                            npos = lane;
                            for (uint32_t jq = 0; jq < ncand; ++jq) {
                                const uint32_t j = cand[jq];
                                if (aux[j] == 0xFFu) continue;
                                const BlockTriplet o = list[j];
                                npos += before(o.weight, o.rT, o.key, lpos[j], ew, es, eb, et) ? 1u : 0u;
                            }
                        }
                        __syncwarp();
                        if (npos < K) {
                            top_w[npos] = ew, top_s[npos] = es, top_rb[npos] = erb;
                            top_b[npos] = eb, top_t[npos] = et;
                        }
                        for (uint32_t q0 = 0; q0 < ncand; q0 += 32) {
                            const uint32_t i = (q0 + lane < ncand) ? cand[q0 + lane] : 0xFFFFu;
                            if (i != 0xFFFFu && aux[i] != 0xFFu) {
                                const BlockTriplet c = list[i];
                                const uint32_t r = aux[i];
                                top_w[r] = c.weight, top_s[r] = c.rT, top_rb[r] = c.curvature;
                                top_b[r] = c.key, top_t[r] = lpos[i];
                            }
                        }
                        ntop = (ntop + nkept < K) ? (ntop + nkept) : K;
                    }
                } else {
                    for (uint32_t i0 = 0; i0 < nlist; i0 += 32) {
                        const uint32_t i = i0 + lane;
                        uint32_t rank = 0xFFu;
                        if (i < nlist) {
                            const BlockTriplet c = list[i];
                            const uint32_t ct = lpos[i];
                            bool in = (c.key != 0xFFFFFFFFu);
                            if (in && ntop == K)  // cannot displace anything: skip the ranking
                                in = before(c.weight, c.rT, c.key, ct, top_w[K - 1], top_s[K - 1],
                                            top_b[K - 1], top_t[K - 1]);
                            if (in) {
                                rank = 0;
                                for (uint32_t q = 0; q < ntop; ++q)
                                    rank += before(top_w[q], top_s[q], top_b[q], top_t[q], c.weight, c.rT,
                                                   c.key, ct) ? 1u : 0u;
                                for (uint32_t j = 0; j < nlist && rank < K; ++j) {
                                    const BlockTriplet o = list[j];
                                    if (j != i && o.key != 0xFFFFFFFFu &&
                                        before(o.weight, o.rT, o.key, lpos[j], c.weight, c.rT, c.key, ct))
                                        ++rank;
                                }
                                if (rank >= K) rank = 0xFFu;
                            }
                            aux[i] = uint8_t(rank);
                        }
                        nkept += __popc(__ballot_sync(0xffffffffu, rank != 0xFFu));
                    }
                    __syncwarp();
                    if (nkept) {
                        // current entries move down by the number of new entries sorted before them
                        float ew = 0.f, es = 0.f, erb = 0.f;
                        uint32_t eb = 0, et = 0, npos = 0xFFu;
                        if (lane < ntop) {
                            ew = top_w[lane], es = top_s[lane], erb = top_rb[lane];
                            eb = top_b[lane], et = top_t[lane];
                            npos = lane;
                            for (uint32_t j = 0; j < nlist; ++j) {
                                if (aux[j] == 0xFFu) continue;
                                const BlockTriplet o = list[j];
                                npos += before(o.weight, o.rT, o.key, lpos[j], ew, es, eb, et) ? 1u : 0u;
                            }
                        }
                        __syncwarp();
                        if (npos < K) {
                            top_w[npos] = ew, top_s[npos] = es, top_rb[npos] = erb;
                            top_b[npos] = eb, top_t[npos] = et;
                        }
                        for (uint32_t i0 = 0; i0 < nlist; i0 += 32) {
                            const uint32_t i = i0 + lane;
                            if (i < nlist && aux[i] != 0xFFu) {
                                const BlockTriplet c = list[i];
                                const uint32_t r = aux[i];
                                top_w[r] = c.weight, top_s[r] = c.rT, top_rb[r] = c.curvature;
                                top_b[r] = c.key, top_t[r] = lpos[i];
                            }
                        }
                        ntop = (ntop + nkept < K) ? (ntop + nkept) : K;
                    }
                }
                __syncwarp();
                acc_trip += nlist;
                nlist = 0;
            }
            if (!have_block) break;

            // ---- evaluate the (row, mid-top) pairs inside the windows, 32 at a time ----
            const uint32_t base_n = nlist;
            acc_visited += total;
#ifndef B200_PREFILTER_MIN_PAIRS
#define B200_PREFILTER_MIN_PAIRS 512u
#endif
            if (DENSE && total >= B200_PREFILTER_MIN_PAIRS) {
                // Blocks with hundreds of pairs (busy events). All pairs go through a
                // division-free pre-filter (only U and V of both doublets) 32 at a time; the
                // survivors queue up in `pend` and take the exact cuts 32 at a time, at full lane
                // occupancy, in order (the rows must stay contiguous in the list).
                auto eval_exact = [&](const uint32_t cnt) {
                    bool ok = false;
                    uint32_t key = 0, pos_t = 0, r = 0;
                    float curvature = 0.f, impact = 0.f, rT = 0.f;
                    if (lane < cnt) {
                        const uint32_t e = pend[lane];
                        r = e >> 27;
                        const uint32_t tt = e & 0x7FFFFFFu;
                        const float4 ba = __ldg(&LB[row0 + r].a);
                        const float4 bb = __ldg(&LB[row0 + r].b);
                        const float4 ta = __ldg(&LT[tt].a);
                        const float4 tb = __ldg(&LT[tt].b);
                        LinCircle lb, lt;
                        lb.cotTheta = ba.x, lb.iDeltaR = ba.y, lb.Er = ba.z, lb.U = ba.w;
                        lb.V = bb.x, lb.Zo = bb.y;
                        lt.cotTheta = ta.x, lt.iDeltaR = ta.y, lt.Er = ta.z, lt.U = ta.w;
                        lt.V = tb.x, lt.Zo = 0.f;
                        float is2, s2;
                        triplet_row_constants(cfg, lb.cotTheta, is2, s2);
                        ok = triplet_is_compatible(cfg, rM, varRM, varZM, lb, lt, is2, s2, curvature,
                                                   impact);
                        key = __float_as_uint(tb.y);
                        rT = tb.z;
                        pos_t = __float_as_uint(tb.w);
                    }
                    const uint32_t mk = __ballot_sync(0xffffffffu, ok);
                    const uint32_t k = nlist + __popc(mk & ltmask);
                    if (ok && k < a.list_cap) {
                        BlockTriplet e;
                        e.key = key;
                        e.curvature = curvature;
                        e.weight = -impact * cfg.impactWeightFactor;
                        e.rT = rT;
                        list[k] = e;
                        lpos[k] = pos_t;
                        lrow[k] = row0 + r;
                    }
                    nlist += __popc(mk);
                };
                uint32_t npend = 0;
                for (uint32_t p0 = 0; p0 < total; p0 += 32) {
                    const uint32_t p = p0 + lane;
                    const uint32_t r = owner_lane(incl, p);
                    const uint32_t er = __shfl_sync(0xffffffffu, excl, r);
                    const uint32_t lor = __shfl_sync(0xffffffffu, lo, r);
                    bool maybe = false;
                    uint32_t tt = 0;
                    if (p < total) {
                        tt = lor + (p - er);
                        const float Ub = __ldg(&LB[row0 + r].a.w), Vb = __ldg(&LB[row0 + r].b.x);
                        const float Ut = __ldg(&LT[tt].a.w), Vt = __ldg(&LT[tt].b.x);
                        maybe = !triplet_certainly_rejected(cfg, rM, Ub, Vb, Ut, Vt);
                    }
                    const uint32_t mm = __ballot_sync(0xffffffffu, maybe);
                    if (maybe) pend[npend + __popc(mm & ltmask)] = (r << 27) | (tt & 0x7FFFFFFu);
                    npend += __popc(mm);
                    __syncwarp();
                    if (npend >= 32u) {
                        eval_exact(32u);  // first in, first out
                        __syncwarp();
                        npend -= 32u;
                        const uint32_t mv = (lane < npend) ? pend[32u + lane] : 0u;
                        __syncwarp();
                        if (lane < npend) pend[lane] = mv;
                        __syncwarp();
                    }
                }
                if (npend) eval_exact(npend);
            } else {
                for (uint32_t p0 = 0; p0 < total; p0 += 32) {
                    const uint32_t p = p0 + lane;
                    const bool valid = p < total;
                    const uint32_t r = owner_lane(incl, p);
                    const uint32_t er = __shfl_sync(0xffffffffu, excl, r);
                    const uint32_t lor = __shfl_sync(0xffffffffu, lo, r);
                    bool ok = false;
                    uint32_t key = 0, pos_t = 0;
                    float curvature = 0.f, impact = 0.f, rT = 0.f;
                    if (valid) {
                        const uint32_t tt = lor + (p - er);
                        const float4 ba = __ldg(&LB[row0 + r].a);
                        const float4 bb = __ldg(&LB[row0 + r].b);
                        const float4 ta = __ldg(&LT[tt].a);
                        const float4 tb = __ldg(&LT[tt].b);
                        LinCircle lb, lt;
                        lb.cotTheta = ba.x, lb.iDeltaR = ba.y, lb.Er = ba.z, lb.U = ba.w;
                        lb.V = bb.x, lb.Zo = bb.y;
                        lt.cotTheta = ta.x, lt.iDeltaR = ta.y, lt.Er = ta.z, lt.U = ta.w;
                        lt.V = tb.x, lt.Zo = 0.f;
                        float is2, s2;
                        triplet_row_constants(cfg, lb.cotTheta, is2, s2);
                        ok = triplet_is_compatible(cfg, rM, varRM, varZM, lb, lt, is2, s2, curvature,
                                                   impact);
                        key = __float_as_uint(tb.y);
                        rT = tb.z;
                        pos_t = __float_as_uint(tb.w);
                    }
                    const uint32_t mk = __ballot_sync(0xffffffffu, ok);
                    const uint32_t k = nlist + __popc(mk & ltmask);
                    if (ok && k < a.list_cap) {
                        BlockTriplet e;
                        e.key = key;
                        e.curvature = curvature;
                        e.weight = -impact * cfg.impactWeightFactor;
                        e.rT = rT;
                        list[k] = e;
                        lpos[k] = pos_t;
                        lrow[k] = row0 + r;
                    }
                    nlist += __popc(mk);
                }
            }
            __syncwarp();
            if (nlist > a.list_cap) {  // only possible when the list was empty before the block
                if (rows > 1) {
                    rows >>= 1;  // redo this block with fewer rows
                    nlist = base_n;
                    continue;
                }
                // a single row with more triplets than the list holds: the slow path
                // (triplets_slow_middles) redoes this middle row by row through global memory;
                // rows before row0 were already counted and dumped here. The row loop ends through
                // its normal exit (no rows left), the final selection below is skipped.
                if (lane == 0) {
                    const uint32_t e = atomicAdd(&a.ctrl->n_slow, 1u);
                    a.slow_list[2 * e] = m;
                    a.slow_list[2 * e + 1] = row0;
                }
                nlist = 0;
                row0 = HANDED_OVER - nrows;
            }
            row0 += nrows;
            rows = 32;
        }
        if (row0 == HANDED_OVER) {
            __syncwarp();
            continue;
        }
        // ---- final per-middle selection (seed_filtering.cpp:84-122) ----
        {
            // lane i decides entry i; survivors keep their order
            const bool keep = lane < ntop && (lane == 0 || cut_per_middle_sp(cfg, top_rb[lane], top_w[lane]));
            const uint32_t km = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const uint32_t o = __popc(km & ltmask);
                a.seed_b[size_t(m) * K + o] = top_b[lane];
                a.seed_t[size_t(m) * K + o] = top_t[lane];
                a.seed_w[size_t(m) * K + o] = top_w[lane];
            }
            if (lane == 0) a.seed_cnt[m] = __popc(km);
        }
        __syncwarp();
    }
    if (lane == 0) {
        atomicAdd(&s_ntrip, acc_trip);
        atomicAdd(&s_tests, acc_tests);
        atomicAdd(&s_visited, acc_visited);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_ntrip) atomicAdd(&a.ctrl->n_triplets, s_ntrip);
        if (s_tests) atomicAdd(&a.ctrl->triplet_tests, s_tests);
        if (s_visited) atomicAdd(&a.ctrl->triplet_visited, s_visited);
    }
}

// ---------------------------------------------------------------------------
// seeds in CPU order + counters. The offset of a middle's seeds is the exclusive prefix sum of
// the per-middle counts: computed here in a single pass with decoupled look-back (status word
// {state:2 @32, value:32}; the words and the ticket counter live in the part of the workspace
// that is cleared at the start of every event).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BIN_THREADS)
k_seed_gather(const uint32_t n_sp, const uint32_t K, const Control* __restrict__ ctrl,
              const uint32_t* __restrict__ seed_cnt, const uint32_t* __restrict__ seed_b,
              const uint32_t* __restrict__ seed_t, const float* __restrict__ seed_w,
              const uint32_t* __restrict__ sorted_index, const uint32_t seed_capacity,
              uint32_t* __restrict__ out_b, uint32_t* __restrict__ out_m,
              uint32_t* __restrict__ out_t, float* __restrict__ out_q, uint32_t* __restrict__ out_n,
              b200seed_counters* __restrict__ counters, const uint32_t* __restrict__ n_sp_dev,
              unsigned long long* __restrict__ status, unsigned long long* __restrict__ ticket_ctr,
              uint32_t* __restrict__ sticky_overflow, const __grid_constant__ DevCfg cfg,
              const __grid_constant__ TripletArgs ta) {
    __shared__ uint32_t s_tile, s_warp[BIN_THREADS / 32], s_prefix;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = uint32_t(atomicAdd(ticket_ctr, 1ull));
    __syncthreads();
    const uint32_t tile = s_tile;
    // Middles the triplet kernel handed over (a mid-bottom doublet with more accepted triplets than
    // its shared-memory list holds; none for ordinary events): the CTA that drew tile 0 — it is
    // running, tiles are handed out in the order the CTAs start — finishes them first, the others
    // wait for it before they read the per-middle seed counts.
    if (ta.ctrl->n_slow != 0u) {
        __shared__ uint32_t s_top[(BIN_THREADS / 32) * 5 * MAX_TOPK];
        volatile uint32_t* done = &ta.ctrl->slow_done;
        if (tile == 0) {
            triplets_slow_middles(cfg, ta, s_top);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) *done = 1u;
        } else {
            if (threadIdx.x == 0)
                while (*done == 0u) __nanosleep(200);
            __syncthreads();
            __threadfence();
        }
    }
    const uint32_t m = tile * BIN_THREADS + threadIdx.x;
    const uint32_t n_valid = ctrl->n_valid;
    const uint32_t n = (m < n_valid) ? seed_cnt[m] : 0u;
    const uint32_t incl = warp_incl_scan(n, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < BIN_THREADS / 32; ++w) {
        const uint32_t c = s_warp[w];
        if (w < int(warp)) before += c;
        total += c;
    }
    if (warp == 0) {
        volatile unsigned long long* st = status;
        uint32_t prefix = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = (FORM_PREFIX << 32) | total;
        } else {
            if (lane == 0) st[tile] = (FORM_AGGREGATE << 32) | total;
            int base = int(tile) - 1;
            while (true) {
                const int j = base - int(lane);
                unsigned long long w = FORM_PREFIX << 32;  // before tile 0: prefix 0
                if (j >= 0) {
                    do {
                        w = st[j];
                    } while ((w >> 32) == 0ull);
                }
                const uint32_t is_pref = __ballot_sync(0xffffffffu, (w >> 32) == FORM_PREFIX);
                const uint32_t first = is_pref ? uint32_t(__ffs(int(is_pref)) - 1) : 32u;
                uint32_t v = (lane <= first) ? uint32_t(w) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                prefix += v;
                if (is_pref) break;
                base -= 32;
            }
            if (lane == 0) st[tile] = (FORM_PREFIX << 32) | (prefix + total);
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == gridDim.x - 1) {  // the last tile knows the total
                const uint32_t all = prefix + total;
                const uint32_t nout = all < seed_capacity ? all : seed_capacity;
                *out_n = nout;
                // a truncated event is never silent: the handle's host-mapped word collects the
                // overflow bits even when the caller passed no counters record
                // (volatile: the slow path above may have updated the block from another SM)
                const volatile Control* vc = ctrl;
                const uint32_t ovf = vc->overflow | (all > seed_capacity ? B200SEED_OVF_SEEDS : 0u);
                if (ovf != 0u && sticky_overflow) atomicOr_system(sticky_overflow, ovf);
                if (counters) {
                    b200seed_counters c;
                    c.n_spacepoints = dev_count(n_sp, n_sp_dev);
                    c.n_valid = n_valid;
                    c.n_active_middles = vc->n_active;
                    c.n_mid_bot = vc->n_mid_bot;
                    c.n_mid_top = vc->n_mid_top;
                    c.n_triplets = vc->n_triplets;
                    c.n_seeds = nout;
                    c.overflow = ovf;
                    c.pair_tests = vc->pair_tests;
                    c.triplet_tests = vc->triplet_tests;
                    c.pair_visited = vc->pair_visited;
                    c.n_fallback_middles = vc->n_fallback;
                    c.reserved_ = 0u;
                    c.triplet_visited = vc->triplet_visited;
                    *counters = c;
                }
            }
        }
    }
    __syncthreads();
    if (n == 0) return;
    const uint32_t off = s_prefix + before + incl - n;
    const uint32_t mi = sorted_index[m];
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t o = off + k;
        if (o < seed_capacity) {
            out_b[o] = sorted_index[seed_b[size_t(m) * K + k]];
            out_m[o] = mi;
            out_t[o] = sorted_index[seed_t[size_t(m) * K + k]];
            out_q[o] = seed_w[size_t(m) * K + k];
        }
    }
}

// ---------------------------------------------------------------------------
// seed parameter estimation: one thread per seed
// (track_params_estimation_helper.hpp:47-129, estimate_track_params.ipp:29-87)
// ---------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float v3dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 v3cross(V3 a, V3 b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
__device__ __forceinline__ V3 v3normalize(V3 a) {
    const float s = 1.f / sqrt_rn(v3dot(a, a));
    return {s * a.x, s * a.y, s * a.z};
}
__device__ __forceinline__ float perp2(float x, float y) { return sqrt_rn(x * x + y * y); }

// 4x4 determinant / inverse by cofactor expansion, m[col][row] — the "hard-coded" 4x4 forms of the
// array plugin's transform3 (restated from the published implementation; the library is absent).
__device__ __forceinline__ float det44(const float (&m)[4][4]) {
        return m[0][3] * m[1][2] * m[2][1] * m[3][0] - m[0][2] * m[1][3] * m[2][1] * m[3][0] -
               m[0][3] * m[1][1] * m[2][2] * m[3][0] + m[0][1] * m[1][3] * m[2][2] * m[3][0] +
               m[0][2] * m[1][1] * m[2][3] * m[3][0] - m[0][1] * m[1][2] * m[2][3] * m[3][0] -
               m[0][3] * m[1][2] * m[2][0] * m[3][1] + m[0][2] * m[1][3] * m[2][0] * m[3][1] +
               m[0][3] * m[1][0] * m[2][2] * m[3][1] - m[0][0] * m[1][3] * m[2][2] * m[3][1] -
               m[0][2] * m[1][0] * m[2][3] * m[3][1] + m[0][0] * m[1][2] * m[2][3] * m[3][1] +
               m[0][3] * m[1][1] * m[2][0] * m[3][2] - m[0][1] * m[1][3] * m[2][0] * m[3][2] -
               m[0][3] * m[1][0] * m[2][1] * m[3][2] + m[0][0] * m[1][3] * m[2][1] * m[3][2] +
               m[0][1] * m[1][0] * m[2][3] * m[3][2] - m[0][0] * m[1][1] * m[2][3] * m[3][2] -
               m[0][2] * m[1][1] * m[2][0] * m[3][3] + m[0][1] * m[1][2] * m[2][0] * m[3][3] +
               m[0][2] * m[1][0] * m[2][1] * m[3][3] - m[0][0] * m[1][2] * m[2][1] * m[3][3] -
               m[0][1] * m[1][0] * m[2][2] * m[3][3] + m[0][0] * m[1][1] * m[2][2] * m[3][3];
}
__device__ __forceinline__ void inverse44(const float (&m)[4][4], float (&i)[4][4]) {
        i[0][0] = m[1][2] * m[2][3] * m[3][1] - m[1][3] * m[2][2] * m[3][1] + m[1][3] * m[2][1] * m[3][2] -
                  m[1][1] * m[2][3] * m[3][2] - m[1][2] * m[2][1] * m[3][3] + m[1][1] * m[2][2] * m[3][3];
        i[0][1] = m[0][3] * m[2][2] * m[3][1] - m[0][2] * m[2][3] * m[3][1] - m[0][3] * m[2][1] * m[3][2] +
                  m[0][1] * m[2][3] * m[3][2] + m[0][2] * m[2][1] * m[3][3] - m[0][1] * m[2][2] * m[3][3];
        i[0][2] = m[0][2] * m[1][3] * m[3][1] - m[0][3] * m[1][2] * m[3][1] + m[0][3] * m[1][1] * m[3][2] -
                  m[0][1] * m[1][3] * m[3][2] - m[0][2] * m[1][1] * m[3][3] + m[0][1] * m[1][2] * m[3][3];
        i[0][3] = m[0][3] * m[1][2] * m[2][1] - m[0][2] * m[1][3] * m[2][1] - m[0][3] * m[1][1] * m[2][2] +
                  m[0][1] * m[1][3] * m[2][2] + m[0][2] * m[1][1] * m[2][3] - m[0][1] * m[1][2] * m[2][3];
        i[1][0] = m[1][3] * m[2][2] * m[3][0] - m[1][2] * m[2][3] * m[3][0] - m[1][3] * m[2][0] * m[3][2] +
                  m[1][0] * m[2][3] * m[3][2] + m[1][2] * m[2][0] * m[3][3] - m[1][0] * m[2][2] * m[3][3];
        i[1][1] = m[0][2] * m[2][3] * m[3][0] - m[0][3] * m[2][2] * m[3][0] + m[0][3] * m[2][0] * m[3][2] -
                  m[0][0] * m[2][3] * m[3][2] - m[0][2] * m[2][0] * m[3][3] + m[0][0] * m[2][2] * m[3][3];
        i[1][2] = m[0][3] * m[1][2] * m[3][0] - m[0][2] * m[1][3] * m[3][0] - m[0][3] * m[1][0] * m[3][2] +
                  m[0][0] * m[1][3] * m[3][2] + m[0][2] * m[1][0] * m[3][3] - m[0][0] * m[1][2] * m[3][3];
        i[1][3] = m[0][2] * m[1][3] * m[2][0] - m[0][3] * m[1][2] * m[2][0] + m[0][3] * m[1][0] * m[2][2] -
                  m[0][0] * m[1][3] * m[2][2] - m[0][2] * m[1][0] * m[2][3] + m[0][0] * m[1][2] * m[2][3];
        i[2][0] = m[1][1] * m[2][3] * m[3][0] - m[1][3] * m[2][1] * m[3][0] + m[1][3] * m[2][0] * m[3][1] -
                  m[1][0] * m[2][3] * m[3][1] - m[1][1] * m[2][0] * m[3][3] + m[1][0] * m[2][1] * m[3][3];
        i[2][1] = m[0][3] * m[2][1] * m[3][0] - m[0][1] * m[2][3] * m[3][0] - m[0][3] * m[2][0] * m[3][1] +
                  m[0][0] * m[2][3] * m[3][1] + m[0][1] * m[2][0] * m[3][3] - m[0][0] * m[2][1] * m[3][3];
        i[2][2] = m[0][1] * m[1][3] * m[3][0] - m[0][3] * m[1][1] * m[3][0] + m[0][3] * m[1][0] * m[3][1] -
                  m[0][0] * m[1][3] * m[3][1] - m[0][1] * m[1][0] * m[3][3] + m[0][0] * m[1][1] * m[3][3];
        i[2][3] = m[0][3] * m[1][1] * m[2][0] - m[0][1] * m[1][3] * m[2][0] - m[0][3] * m[1][0] * m[2][1] +
                  m[0][0] * m[1][3] * m[2][1] + m[0][1] * m[1][0] * m[2][3] - m[0][0] * m[1][1] * m[2][3];
        i[3][0] = m[1][2] * m[2][1] * m[3][0] - m[1][1] * m[2][2] * m[3][0] - m[1][2] * m[2][0] * m[3][1] +
                  m[1][0] * m[2][2] * m[3][1] + m[1][1] * m[2][0] * m[3][2] - m[1][0] * m[2][1] * m[3][2];
        i[3][1] = m[0][1] * m[2][2] * m[3][0] - m[0][2] * m[2][1] * m[3][0] + m[0][2] * m[2][0] * m[3][1] -
                  m[0][0] * m[2][2] * m[3][1] - m[0][1] * m[2][0] * m[3][2] + m[0][0] * m[2][1] * m[3][2];
        i[3][2] = m[0][2] * m[1][1] * m[3][0] - m[0][1] * m[1][2] * m[3][0] - m[0][2] * m[1][0] * m[3][1] +
                  m[0][0] * m[1][2] * m[3][1] + m[0][1] * m[1][0] * m[3][2] - m[0][0] * m[1][1] * m[3][2];
        i[3][3] = m[0][1] * m[1][2] * m[2][0] - m[0][2] * m[1][1] * m[2][0] + m[0][2] * m[1][0] * m[2][1] -
                  m[0][0] * m[1][2] * m[2][1] - m[0][1] * m[1][0] * m[2][2] + m[0][0] * m[1][1] * m[2][2];
    const float s = 1.f / det44(m);
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) i[c][r] *= s;
}

// Inhomogeneous field on a regular grid: the reference's cuda::inhom_global_bfield_backend_t =
// covfie affine< linear< clamp< strided< device array of float3 >>>>
// (device/cuda/src/utils/magnetic_field_types.hpp:27-32), sampled at the bottom spacepoint
// (estimate_track_params.ipp:45-50). covfie is third-party and absent here; restated from its
// published semantics: index-space coordinate c = A (x, y, z, 1); trilinear interpolation
// between floor(c) and floor(c) + 1 with the indices clamped to the grid; row-major storage.
__device__ __forceinline__ V3 field_at(const b200seed_field_grid& fg, V3 p) {
    float c[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        c[i] = ((fg.affine[4 * i] * p.x + fg.affine[4 * i + 1] * p.y) + fg.affine[4 * i + 2] * p.z) +
               fg.affine[4 * i + 3];
    int i0[3], i1[3];
    float w1[3], w0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float fl = floorf(c[k]);
        w1[k] = c[k] - fl;
        w0[k] = 1.f - w1[k];
        const int hi = int(fg.size[k]) - 1;
        // NaN / out-of-range coordinates clamp like covfie's clamp backend
        const float lo_f = (fl >= 0.f) ? ((fl <= float(hi)) ? fl : float(hi)) : 0.f;
        const float up = fl + 1.f;
        const float hi_f = (up >= 0.f) ? ((up <= float(hi)) ? up : float(hi)) : 0.f;
        i0[k] = int(lo_f);
        i1[k] = int(hi_f);
    }
    V3 r{0.f, 0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int ix = (n & 4) ? i1[0] : i0[0], iy = (n & 2) ? i1[1] : i0[1],
                  iz = (n & 1) ? i1[2] : i0[2];
        const float w = (((n & 4) ? w1[0] : w0[0]) * ((n & 2) ? w1[1] : w0[1])) *
                        ((n & 1) ? w1[2] : w0[2]);
        const float* f = fg.data + 3 * ((size_t(ix) * fg.size[1] + iy) * fg.size[2] + iz);
        r.x = r.x + w * __ldg(f);
        r.y = r.y + w * __ldg(f + 1);
        r.z = r.z + w * __ldg(f + 2);
    }
    return r;
}

__global__ void __launch_bounds__(128)
k_estimate_params(const b200seed_tpe_cfg cfg, const uint32_t* __restrict__ n_seeds_dev,
                  const uint32_t seed_capacity, const uint32_t* __restrict__ sd_b,
                  const uint32_t* __restrict__ sd_m, const uint32_t* __restrict__ sd_t,
                  const float* __restrict__ xyz, const uint32_t* __restrict__ sp_meas,
                  const float* __restrict__ meas_local, const uint64_t* __restrict__ meas_surface,
                  const float bx, const float by, const float bz, const b200seed_field_grid fg,
                  b200seed_bound_params* __restrict__ out,
                  b200seed_bound_params_diag* __restrict__ out_diag,
                  b200seed_seed_params* __restrict__ out_compact,
                  b200seed_bound_params_packed* __restrict__ out_packed) {
    // Records are 176 B: written one per lane they would cost 32 sectors per store
    // instruction. Each warp builds its 32 records (5632 contiguous bytes) in shared memory
    // and streams them out as float4 rows.
    __shared__ __align__(16) float s_rec[4][32 * 44];
    uint32_t n = *n_seeds_dev;
    if (n > seed_capacity) n = seed_capacity;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i0 = blockIdx.x * blockDim.x + warp * 32;  // first seed of this warp
    if (i0 >= n) return;
    const uint32_t i = i0 + lane;
    const bool live = i < n;
    float* rec = s_rec[warp];
    {
        float4* z4 = reinterpret_cast<float4*>(rec);
#pragma unroll
        for (int k = 0; k < 11; ++k) z4[k * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    const uint32_t ib = live ? sd_b[i] : 0u, im = live ? sd_m[i] : 0u, it = live ? sd_t[i] : 0u;
    const V3 p0{xyz[3 * size_t(ib)], xyz[3 * size_t(ib) + 1], xyz[3 * size_t(ib) + 2]};
    const V3 p1{xyz[3 * size_t(im)], xyz[3 * size_t(im) + 1], xyz[3 * size_t(im) + 2]};
    const V3 p2{xyz[3 * size_t(it)], xyz[3 * size_t(it) + 1], xyz[3 * size_t(it) + 2]};
    const V3 bfield = fg.data ? field_at(fg, p0) : V3{bx, by, bz};
    const V3 relVec = v3sub(p1, p0);
    const V3 newZ = v3normalize(bfield);
    const V3 newY = v3normalize(v3cross(newZ, relVec));
    const V3 newX = v3cross(newY, newZ);
    // transform3(translation = p0, x, y, z): 4x4 matrix and its cofactor-expansion inverse;
    // point_to_local(p) = rotate(inverse, p) + translation column of the inverse
    // (track_params_estimation_helper.hpp:78-85; the array plugin's transform3)
    float tm[4][4], ti[4][4];
    tm[0][0] = newX.x, tm[0][1] = newX.y, tm[0][2] = newX.z, tm[0][3] = 0.f;
    tm[1][0] = newY.x, tm[1][1] = newY.y, tm[1][2] = newY.z, tm[1][3] = 0.f;
    tm[2][0] = newZ.x, tm[2][1] = newZ.y, tm[2][2] = newZ.z, tm[2][3] = 0.f;
    tm[3][0] = p0.x, tm[3][1] = p0.y, tm[3][2] = p0.z, tm[3][3] = 1.f;
    inverse44(tm, ti);
    auto to_local = [&](const V3& p) {
        return V3{(ti[0][0] * p.x + ti[1][0] * p.y + ti[2][0] * p.z) + ti[3][0],
                  (ti[0][1] * p.x + ti[1][1] * p.y + ti[2][1] * p.z) + ti[3][1],
                  (ti[0][2] * p.x + ti[1][2] * p.y + ti[2][2] * p.z) + ti[3][2]};
    };
    const V3 local1 = to_local(p1);
    const V3 local2 = to_local(p2);
    const float den1 = local1.x * local1.x + local1.y * local1.y;
    const float den2 = local2.x * local2.x + local2.y * local2.y;
    const float u1 = local1.x / den1, v1 = local1.y / den1;
    const float u2 = local2.x / den2, v2 = local2.y / den2;
    const float A = (v2 - v1) / (u2 - u1);
    const float B = v2 - A * u2;
    const float R = -perp2(1.f, A) / (2.f * B);
    const float invTanTheta =
        local2.z / (2.f * R * asinf(perp2(local2.x, local2.y) / (2.f * R)));
    const V3 td{1.f, A, perp2(1.f, A) * invTanTheta};
    const V3 nd = v3normalize(td);
    const V3 dir{newX.x * nd.x + newY.x * nd.y + newZ.x * nd.z,
                 newX.y * nd.x + newY.y * nd.y + newZ.y * nd.z,
                 newX.z * nd.x + newY.z * nd.y + newZ.z * nd.z};
    const float phi = atan2f(dir.y, dir.x);
    const float theta = atan2f(perp2(dir.x, dir.y), dir.z);
    const float qOverPt = 1.f / (R * sqrt_rn(v3dot(bfield, bfield)));
    const float qop = qOverPt / perp2(1.f, invTanTheta);
    const uint32_t mi = sp_meas ? sp_meas[ib] : ib;

    // the record of this lane inside the warp's staging buffer (same layout as the output)
    const float sigma_qopt = cfg.initial_sigma_qopt * sinf(theta);
    const float sigma_pt_rel = cfg.initial_sigma_pt_rel * qop;
    const float sigma_theta = qop / tanf(theta);
    float var[6];
    float var_theta = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        float v = cfg.initial_sigma[j] * cfg.initial_sigma[j];
        if (j == 4) {
            v += sigma_qopt * sigma_qopt;
            v += sigma_pt_rel * sigma_pt_rel;
            v += var_theta * sigma_theta * sigma_theta;
        }
        v *= cfg.initial_inflation[j];
        if (j == 3) var_theta = v;
        var[j] = v;
    }
    if (out_compact) {
        // what only the device can compute: 16 bytes per seed (the rest of the record is copied
        // from the bottom spacepoint's measurement or is a constant of the configuration)
        if (live) reinterpret_cast<float4*>(out_compact)[i] = make_float4(phi, theta, qop, var[4]);
        return;
    }
    const uint64_t link = meas_surface ? meas_surface[mi] : 0ull;
    const float loc0 = meas_local ? meas_local[2 * size_t(mi)] : 0.f;
    const float loc1 = meas_local ? meas_local[2 * size_t(mi) + 1] : 0.f;
    const uint32_t nrec = (n - i0 < 32u) ? (n - i0) : 32u;
    if (out_packed) {
        // 32-byte records (no constant variances, no time): 8 floats per lane, streamed out as float4 rows
        b200seed_bound_params_packed* o = reinterpret_cast<b200seed_bound_params_packed*>(rec + lane * 8);
        o->surface_link = link;
        o->loc0 = loc0, o->loc1 = loc1, o->phi = phi, o->theta = theta, o->qop = qop, o->var_qop = var[4];
        __syncwarp();
        const float4* s4 = reinterpret_cast<const float4*>(rec);
        float4* d4 = reinterpret_cast<float4*>(out_packed + i0);
        for (uint32_t k = lane; k < nrec * 2u; k += 32) d4[k] = s4[k];
        return;
    }
    if (out_diag) {
        // 56-byte diagonal records: 14 floats per lane, streamed out as float2 rows
        b200seed_bound_params_diag* o = reinterpret_cast<b200seed_bound_params_diag*>(rec + lane * 14);
        o->surface_link = link;
        o->vec[0] = loc0, o->vec[1] = loc1, o->vec[2] = phi, o->vec[3] = theta, o->vec[4] = qop;
        o->vec[5] = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) o->cov_diag[j] = var[j];
        __syncwarp();
        const float2* s2 = reinterpret_cast<const float2*>(rec);
        float2* d2 = reinterpret_cast<float2*>(out_diag + i0);
        for (uint32_t k = lane; k < nrec * 7u; k += 32) d2[k] = s2[k];
        return;
    }
    b200seed_bound_params* o = reinterpret_cast<b200seed_bound_params*>(rec + lane * 44);
    o->surface_link = link;
    o->vec[0] = loc0;
    o->vec[1] = loc1;
    o->vec[2] = phi;
    o->vec[3] = theta;
    o->vec[4] = qop;
    o->vec[5] = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) o->cov[j * 6 + j] = var[j];
    __syncwarp();
    float* dst = reinterpret_cast<float*>(out + i0);
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(rec);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (uint32_t k = lane; k < nrec * 11u; k += 32) d4[k] = s4[k];
    } else {  // caller buffer only 8-byte aligned
        const float2* s2 = reinterpret_cast<const float2*>(rec);
        float2* d2 = reinterpret_cast<float2*>(dst);
        for (uint32_t k = lane; k < nrec * 22u; k += 32) d2[k] = s2[k];
    }
}

// ---------------------------------------------------------------------------
// FP32 issue-rate probe for the roofline denominator: 8 independent chains of
// non-fused FMUL + FADD per thread (this library is built with -fmad=false because the
// cut arithmetic may not contract), 16 ops per inner iteration.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp32_probe(float* out, const int iters, const float a,
                                                    const float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f,
          x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
        x0 = x0 * a + b, x1 = x1 * a + b, x2 = x2 * a + b, x3 = x3 * a + b;
        x4 = x4 * a + b, x5 = x5 * a + b, x6 = x6 * a + b, x7 = x7 * a + b;
    }
    const float r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (r == 12345.678f) out[0] = r;  // never true; keeps the chains alive
}

}  // namespace b200seed
