// b200seed_api.cu — C-ABI of libb200seed.so (include/b200seed.h): handle, workspace
// layout, kernel launch sequence, host-buffer convenience path, host probes.
//
// Host-side logic mirrors device::triplet_seeding_algorithm::operator()
// (device/common/src/seeding/triplet_seeding_algorithm.cpp:56-262) minus its seven
// blocking device->host reads: every size the reference reads back is consumed on the
// device instead, and arenas are capacity bounded with an overflow flag.
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/b200seed.h"
#include "../../include/b200seed_probes.h"
#include "seed_kernels.cuh"
#include "seed_tile.cuh"
#include "seed_pool.cuh"
#include "seed_lanes.cuh"

using namespace b200seed;

namespace {

constexpr float unit_mm = 1.f;
constexpr float unit_GeV = 1.f;
constexpr float unit_MeV = 1e-3f;
constexpr float unit_T = static_cast<float>(0.000299792458);
constexpr float unit_degree = static_cast<float>(0.017453292519943295);
constexpr float unit_ns = static_cast<float>(1e-9 * 299792458000.0);

thread_local std::string g_create_error;

inline size_t align_up(size_t v, size_t a) {
    return (v + a - 1) / a * a;
}

struct Layout {
    size_t control, cell_cnt, bin_tot, bin_of, blk_hist, bin_off, cell_off, sorted_index, sorted_bin, sp4,
        var2, csp4, ccanon, cnt, off, seed_cnt, seed_b, seed_t, seed_w, arena_b, arena_t,
        dump, gather_state, spill_list, active_list, group_list, fallback_list, slow_list, seg_info, mid_order, total;
    size_t zero_bytes;  // control block, cell populations, look-back state: cleared per event
    uint32_t nblk;
    uint64_t max_doublets, max_dump;
    CellGrid g;
};

struct TimingSlot {
    const char* name;
    cudaEvent_t start, stop;
};

}  // namespace

struct b200seed_handle {
    b200seed_finder_cfg finder;
    b200seed_grid_cfg grid;
    b200seed_filter_cfg filter;
    b200seed_tpe_cfg tpe;
    DevCfg dev;
    int device = 0;
    uint32_t nbins = 0;
    uint64_t max_doublets_user = 0;
    uint64_t max_dump = 0;
    uint32_t stage_cap_user = 0;
    uint32_t list_cap_user = 0;
    // doublet search: 2 = warp-per-middle k_doublets<0> (default: the faster one),
    // 0 = k_doublets_tile (groups of middles, cp.async.bulk staging), 1 = the same with 16-byte
    // cp.async (B200SEED_DOUBLETS=warp|tile|ldgsts, read at b200seed_create)
    int doublet_mode = 2;
    // k_doublets<0> draws its tickets in cost order (longest middles first: k_cell_scan's classes)
    // instead of grid order: 1 = for events of at least 8k spacepoints (default), 0 / 2 = never /
    // always (B200SEED_DOUBLET_ORDER=grid / cost, for A/B runs and the tests)
    int ordered_tickets = 1;
    // ... and the classes with a scarce side in a launch of their own, from 20k spacepoints on
    // (B200SEED_DOUBLET_SIDES=0: off, =force: at any size)
    int split_sides = 1;
    int pdl = 1;  // ... overlapping the end of k_doublets<0> (B200SEED_PDL=0: plain stream order)
    // triplet search of the light middles: 0 = k_triplets for all (default: faster),
    // 1 = k_triplets_pool (several middles per warp); B200SEED_TRIPLETS=warp|pool
    int triplet_pool = 0;
    // B200SEED_TRIPLETS=lanes: the light middles with fewer than 32 rows go to k_triplets_lanes (one
    // middle per lane), launched as a programmatic dependent of k_triplets. Bit-identical, measured
    // slower (a lane's serial program over <= 31 rows is ~450 us of dependent loads and there are
    // only ~300 such warps per event: triplet stage 110 -> 546 us): off by default.
    int triplet_lanes = 0;
    uint32_t group_max = 0;      // 0 = automatic (by event size)
    float group_zspan_mm = 0.f;  // 0 = default
    int num_sms = 148;
    int smem_optin = 0;
    bool timing = false;
    std::vector<TimingSlot> slots;
    int n_slots_used = 0;
    mutable std::string error;
    // staging for b200seed_run_host
    void* d_stage = nullptr;
    size_t d_stage_bytes = 0;
    b200seed_counters* h_pinned = nullptr;  // counters + n_seeds read-back
    // host-buffer path: pinned landing area of the 16-byte b200seed_seed_params records (and of the
    // bottom indices when the caller did not ask for the seed columns), expanded on the host
    void* h_compact = nullptr;
    size_t h_compact_bytes = 0;
    bool pcie_compact = false;  // B200SEED_PCIE_PARAMS=compact
    bool pcie_packed = false;   // B200SEED_PCIE_PARAMS=packed: 32-byte packed records over PCIe
    // OR of the overflow masks of the events run on this handle since the last
    // b200seed_check_overflow: one pinned, device-mapped word that k_seed_gather writes only when
    // an event was truncated (so a caller that passes d_counters == NULL still learns about it)
    uint32_t* h_sticky = nullptr;
    // look-back state of b200seed_form_spacepoints (status words + ticket counter)
    cudaEvent_t ev_host = nullptr;  // blocking-sync event of the host-buffer path
    unsigned long long* d_form = nullptr;
    size_t form_tiles = 0;
    unsigned long long form_ticket_base = 0;
    uint32_t form_epoch = 0;
};

namespace {

int fail(const b200seed_handle* h, int code, const std::string& msg) {
    if (h)
        h->error = msg;
    else
        g_create_error = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                 \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail(h, B200SEED_ECUDA,                                                \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

std::string overflow_message(uint32_t ovf) {
    std::string m = "event truncated:";
    if (ovf & B200SEED_OVF_DOUBLETS) m += " doublet arena too small (raise b200seed_set_max_doublets);";
    if (ovf & B200SEED_OVF_SEEDS) m += " seed_capacity too small;";
    if (ovf & B200SEED_OVF_DUMP) m += " triplet dump buffer too small;";
    if (ovf & B200SEED_OVF_TRIPLETS) m += " a mid-bottom doublet has more triplets than the list holds;";
    return m;
}

uint64_t default_max_doublets(uint32_t max_sp) {
    // mid-top lists that outgrow the staging area are allocated twice (sort space), and on the
    // densest events nearly all do: 100k particles in |eta| < 1 (N = 4e5) needs 5.9e8 mid-bottom and
    // 2 x 3.2e8 mid-top entries, 4e-3 N^2 was 0.4 % short
    const double q = 5e-3 * double(max_sp) * double(max_sp);
    uint64_t v = q < double(1u << 20) ? (1u << 20) : uint64_t(q);
    if (v > 0xFFFF0000ull) v = 0xFFFF0000ull;
    return v;
}

// Fine (r, z) cells per reference bin (seed_math.cuh, CellGrid): about two cells per
// spacepoint of an average bin, at most 32 x 128, and at most 2^21 cells in total.
CellGrid make_cell_grid(const DevCfg& dev, const b200seed_finder_cfg& finder, uint32_t n_sp) {
    const uint64_t nb64 = uint64_t(dev.nPhi) * dev.nZ;
    const uint32_t nbins = nb64 ? uint32_t(nb64) : 1u;
    const double per_bin = double(n_sp) / double(nbins);
    uint32_t cpb = 1;
#ifndef B200_CELL_DENSITY
#define B200_CELL_DENSITY 2.0
#endif
#ifndef B200_NR_BIG
#define B200_NR_BIG 16u
#endif
    while (cpb < 4096u && double(cpb) < B200_CELL_DENSITY * per_bin) cpb <<= 1;
    if (cpb < 64u) cpb = 64u;
    while (cpb > 1u && uint64_t(cpb) * nbins > (1ull << 21)) cpb >>= 1;
#ifndef B200_NR_MID
#define B200_NR_MID 16u
#endif
    uint32_t nr = cpb >= 2048u ? B200_NR_BIG : (cpb >= 512u ? B200_NR_MID : (cpb >= 64u ? 8u : 1u));
    if (nr > cpb) nr = cpb;
#ifdef B200_FORCE_NR
    nr = B200_FORCE_NR;
    cpb = B200_FORCE_NR * B200_FORCE_NZC;
#endif
    CellGrid g{};
    g.NR = nr;
    g.NZc = cpb / nr;
    g.CPB = cpb;
    g.NZg = dev.nZ * g.NZc;
    const float beam = std::sqrt(finder.beamPos[0] * finder.beamPos[0] +
                                 finder.beamPos[1] * finder.beamPos[1]);
    float rspan = finder.rMax + 2.f * beam + 1.f;
    if (!(rspan > 1.f) || !(rspan < 1e30f)) rspan = 1.f;
    g.rw = rspan / float(nr);
    g.invRw = 1.f / g.rw;
    g.zMin = dev.zAxisMin;
    float zspan = dev.zAxisMax - dev.zAxisMin;
    if (!(zspan > 0.f) || !(zspan < 1e30f)) zspan = 1.f;
    g.invZw = float(g.NZg) / zspan;
    return g;
}

Layout make_layout(const b200seed_handle* h, uint32_t max_sp) {
    Layout L{};
    const size_t n = max_sp ? max_sp : 1;
    L.g = make_cell_grid(h->dev, h->finder, max_sp);
    const size_t ncells = size_t(h->nbins) * L.g.CPB;
    L.nblk = uint32_t((n + BIN_THREADS - 1) / BIN_THREADS);
    L.max_doublets = h->max_doublets_user ? h->max_doublets_user : default_max_doublets(max_sp);
    L.max_dump = h->max_dump;
    const size_t K = h->finder.maxSeedsPerSpM ? h->finder.maxSeedsPerSpM : 1;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    L.control = take(sizeof(Control));
    L.cell_cnt = take(ncells * 4);
    L.bin_tot = take(size_t(h->nbins) * 4);
    // look-back status words of k_seed_gather (one per 256 middles) + its ticket counter
    L.gather_state = take((size_t(L.nblk) + 1) * sizeof(unsigned long long));
    L.zero_bytes = o;
    L.cell_off = take((ncells + 1) * 4);
    L.csp4 = take(n * 16);
    L.ccanon = take(n * 4);
    L.bin_of = take(n * 4);
    L.blk_hist = take(size_t(h->nbins) * L.nblk * 4);
    L.bin_off = take((size_t(h->nbins) + 1) * 4);
    L.sorted_index = take(n * 4);
    L.sorted_bin = take(n * 4);
    L.sp4 = take(n * 16);
    L.var2 = take(n * 8);
    L.cnt = take(2 * n * 4);
    L.off = take(2 * n * 4);
    L.spill_list = take(n * 4);
    L.active_list = take(n * 4 * WORK_CLASSES);
    L.mid_order = take(n * 4);
    L.seg_info = take(size_t(h->nbins) * L.g.NR * 4);
    L.group_list = take(n * 4);
    L.fallback_list = take(n * 4);
    L.seed_cnt = take(n * 4);
    L.seed_b = take(n * K * 4);
    L.seed_t = take(n * K * 4);
    L.seed_w = take(n * K * 4);
    L.slow_list = take(2 * n * 4);
    L.arena_b = take(L.max_doublets * sizeof(DoubletRec));
    L.arena_t = take(L.max_doublets * sizeof(DoubletRec));
    L.dump = take(L.max_dump * sizeof(TripletDumpRec));
    L.total = o;
    return L;
}

// get_axes — core/include/traccc/seeding/spacepoint_binning_helper.hpp:22-110
int compute_axes(const b200seed_grid_cfg& g, DevCfg& d, std::string& why) {
    uint32_t phiBins;
    if (g.bFieldInZ == 0) {
        phiBins = 100;
    } else {
        float minHelixRadius = g.minPt / g.bFieldInZ;
        if (minHelixRadius < g.rMax / 2) {
            why =
                "The value of minHelixRadius cannot be smaller than rMax / 2. Please check the "
                "configuration of bFieldInZ and minPt";
            return -1;
        }
        float maxR2 = g.rMax * g.rMax;
        float xOuter = maxR2 / (2 * minHelixRadius);
        float yOuter = std::sqrt(maxR2 - xOuter * xOuter);
        float outerAngle = std::atan(xOuter / yOuter);
        float innerAngle = 0;
        float rMin = g.rMax;
        if (g.rMax > g.deltaRMax) {
            rMin = g.rMax - g.deltaRMax;
            float innerCircleR2 = (g.rMax - g.deltaRMax) * (g.rMax - g.deltaRMax);
            float xInner = innerCircleR2 / (2 * minHelixRadius);
            float yInner = std::sqrt(innerCircleR2 - xInner * xInner);
            innerAngle = std::atan(xInner / yInner);
        }
        float deltaAngleWithMaxD0 =
            std::fabs(std::asin(g.impactMax / (rMin)) - std::asin(g.impactMax / g.rMax));
        float deltaPhi = (outerAngle - innerAngle + deltaAngleWithMaxD0) /
                         static_cast<float>(g.phiBinDeflectionCoverage);
        if (!(deltaPhi > 0.f) || !std::isfinite(deltaPhi)) {
            why =
                "Delta phi value is equal to or less than zero, leading to an impossible number "
                "of bins (negative or infinite)";
            return -1;
        }
        const long long nb = std::llround(2 * M_PI / deltaPhi + 0.5);
        if (nb < 1 || nb > (1ll << 30)) {
            why = "Delta phi value leads to an impossible number of phi bins";
            return -1;
        }
        phiBins = static_cast<uint32_t>(nb);
    }
    if (!(g.zMax > g.zMin) || !(g.cotThetaMax * g.deltaRMax > 0.f)) {
        why = "empty z range or non-positive cotThetaMax * deltaRMax";
        return -1;
    }
    float zBinSize = g.cotThetaMax * g.deltaRMax;
    uint32_t zBins = std::max(static_cast<uint32_t>(1),
                              static_cast<uint32_t>(std::floor((g.zMax - g.zMin) / zBinSize)));
    d.nPhi = phiBins;
    d.phiAxisMin = g.phiMin;
    d.phiAxisMax = g.phiMax;
    d.nZ = zBins;
    d.zAxisMin = g.zMin;
    d.zAxisMax = g.zMax;
    return 0;
}

void fill_devcfg(const b200seed_finder_cfg& f, const b200seed_filter_cfg& fl, DevCfg& d) {
    d.zMin = f.zMin;
    d.zMax = f.zMax;
    d.phiMin = f.phiMin;
    d.phiMax = f.phiMax;
    d.beamX = f.beamPos[0];
    d.beamY = f.beamPos[1];
    // get_num_rbins: size_t(rMax + vector::norm(beamPos)) — seeding_config.hpp:110-113
    d.numRBins = static_cast<size_t>(
        f.rMax + std::sqrt(f.beamPos[0] * f.beamPos[0] + f.beamPos[1] * f.beamPos[1]));
    d.scope0 = f.neighbor_scope[0];
    d.scope1 = f.neighbor_scope[1];
    d.deltaRMin = f.deltaRMin;
    d.deltaRMax = f.deltaRMax;
    d.cotThetaMax = f.cotThetaMax;
    d.collisionRegionMin = f.collisionRegionMin;
    d.collisionRegionMax = f.collisionRegionMax;
    d.deltaZMax = f.deltaZMax;
    d.minHelixRadius2 = f.minHelixRadius * f.minHelixRadius;
    d.helixImpactMargin2 = (f.minHelixRadius - f.impactMax) * (f.minHelixRadius - f.impactMax);
    d.maxScatteringAngle2 = f.maxScatteringAngle2;
    d.sigmaScattering = f.sigmaScattering;
    d.sigmaScattering2 = f.sigmaScattering * f.sigmaScattering;
    d.minHelixDiameter2 = f.minHelixDiameter2;
    d.pT2perRadius = f.pT2perRadius;
    d.pTPerHelixRadius = f.pTPerHelixRadius;
    d.maxPtScattering = f.maxPtScattering;
    {
        float pTscatter = f.highland / f.maxPtScattering;
        d.pT2scatterMax = pTscatter * pTscatter;
    }
    d.impactMax = f.impactMax;
    d.impactWeightFactor = fl.impactWeightFactor;
    d.deltaInvHelixDiameter = fl.deltaInvHelixDiameter;
    d.compatSeedWeight = fl.compatSeedWeight;
    d.filterDeltaRMin = fl.deltaRMin;
    d.compatSeedLimit = static_cast<uint32_t>(fl.compatSeedLimit);
    d.maxSeedsPerSpM = f.maxSeedsPerSpM;
    d.good_spB_min_radius = fl.good_spB_min_radius;
    d.good_spB_weight_increase = fl.good_spB_weight_increase;
    d.good_spT_max_radius = fl.good_spT_max_radius;
    d.good_spT_weight_increase = fl.good_spT_weight_increase;
    d.good_spB_min_weight = fl.good_spB_min_weight;
    d.seed_min_weight = fl.seed_min_weight;
    d.spB_min_radius = fl.spB_min_radius;
    {
        // bound on |x|, |y| of a valid spacepoint (is_valid_sp: perp to the beam < numRBins)
        const double rb = double(d.numRBins) + std::sqrt(double(f.beamPos[0]) * f.beamPos[0] +
                                                         double(f.beamPos[1]) * f.beamPos[1]) + 1.0;
        const double R2 = double(d.minHelixRadius2);
        d.fast_bounded = (R2 > 0.0 && R2 < 1e12 && rb < 1e5 && rb < 0.99 * std::sqrt(R2)) ? 1u : 0u;
    }
}

struct KernelTimer {
    b200seed_handle* h;
    cudaStream_t s;
    int idx = -1;
    KernelTimer(b200seed_handle* h_, cudaStream_t s_, const char* name) : h(h_), s(s_) {
        if (!h->timing) return;
        if (h->n_slots_used >= int(h->slots.size())) {
            TimingSlot t{name, nullptr, nullptr};
            cudaEventCreate(&t.start);
            cudaEventCreate(&t.stop);
            h->slots.push_back(t);
        }
        idx = h->n_slots_used++;
        h->slots[idx].name = name;
        cudaEventRecord(h->slots[idx].start, s);
    }
    ~KernelTimer() {
        if (idx >= 0) cudaEventRecord(h->slots[idx].stop, s);
    }
};

// Doublet staging capacity per warp (mid-bottoms; mid-tops: half of it), shared memory.
// 384: 4 CTAs x 31 KB fit the 132-KB shared-memory carve-out, which leaves 124 KB of L1
// (512: 200 KB / 56 KB); busier events need the longer lists more than the cache (a 15k-particle
// event loses 20 % with 384: too many middles take the spill pass). Longer lists than 512 cost
// occupancy and lose even on the busiest events (100k particles in |eta| < 1, k_doublets:
// 2048 -> 105 ms at one CTA per SM, 1024 -> 60 ms, 512 -> 40 ms): the spill pass is cheaper.
uint32_t doublet_stage_cap(uint32_t n_sp) {
    return n_sp <= 55000 ? 384 : 512;
}
// Triplet list per warp of k_triplets (shared memory). Small lists mean more resident warps and
// more L1: 96 entries below 80k spacepoints (192 needed the 233-KB carve-out: 153 -> 148 us on the
// 10k-particle event), 128 above — the busiest event (100k particles in |eta| < 1) takes
// 306 ms in k_triplets with 768 entries (one CTA per SM), 144 / 118 / 106 / 110 / 167 ms with
// 256 / 192 / 128 / 96 / 64.
#ifndef B200_LIST_CAP_SMALL
#define B200_LIST_CAP_SMALL 96
#endif
#ifndef B200_LIST_CAP_BIG
#define B200_LIST_CAP_BIG 128
#endif
uint32_t triplet_list_cap(uint32_t n_sp) {
    return n_sp <= 80000 ? B200_LIST_CAP_SMALL : B200_LIST_CAP_BIG;
}

}  // namespace

extern "C" {

const char* b200seed_version(void) {
    return "b200seed 0.1 (sm_100a)";
}

// seedfinder_config::setup() — seeding_config.hpp:123-138
void b200seed_finder_cfg_setup(b200seed_finder_cfg* c) {
    c->highland = 13.6f * unit_MeV * std::sqrt(c->radLengthPerSeed) *
                  (1.f + 0.038f * std::log(c->radLengthPerSeed));
    float maxScatteringAngle = c->highland / c->minPt;
    c->maxScatteringAngle2 = maxScatteringAngle * maxScatteringAngle;
    c->pTPerHelixRadius = c->bFieldInZ;
    c->minHelixDiameter2 = std::pow(c->minPt * 2.f / c->pTPerHelixRadius, 2.f);
    c->minHelixRadius = std::sqrt(c->minHelixDiameter2) / 2.f;
    c->pT2perRadius = std::pow(c->highland / c->pTPerHelixRadius, 2.f);
}

// seedfinder_config in-class defaults — seeding_config.hpp:22-108
void b200seed_finder_cfg_defaults(b200seed_finder_cfg* c) {
    std::memset(c, 0, sizeof(*c));
    c->zMin = -2000.f * unit_mm;
    c->zMax = 2000.f * unit_mm;
    c->rMax = 200.f * unit_mm;
    c->rMin = 33.f * unit_mm;
    c->collisionRegionMin = -250 * unit_mm;
    c->collisionRegionMax = +250 * unit_mm;
    c->phiMin = static_cast<float>(-M_PI);
    c->phiMax = static_cast<float>(M_PI);
    c->minPt = 500.f * unit_MeV;
    c->cotThetaMax = 27.2845f;
    c->deltaRMin = 20 * unit_mm;
    c->deltaRMax = 80 * unit_mm;
    c->deltaZMax = 450 * unit_mm;
    c->impactMax = 10.f * unit_mm;
    c->sigmaScattering = 3.0f;
    c->maxPtScattering = 10.f * unit_GeV;
    c->maxSeedsPerSpM = 5;
    c->bFieldInZ = 1.99724f * unit_T;
    c->beamPos[0] = -.0f * unit_mm;
    c->beamPos[1] = -.0f * unit_mm;
    c->radLengthPerSeed = 0.05f;
    c->zAlign = 0 * unit_mm;
    c->rAlign = 0 * unit_mm;
    c->sigmaError = 5;
    c->phiBinDeflectionCoverage = 1;
    c->neighbor_scope[0] = 1;
    c->neighbor_scope[1] = 1;
    b200seed_finder_cfg_setup(c);
}

// spacepoint_grid_config(const seedfinder_config&) — seeding_config.hpp:145-156
void b200seed_grid_cfg_from_finder(const b200seed_finder_cfg* f, b200seed_grid_cfg* g) {
    g->bFieldInZ = f->bFieldInZ;
    g->minPt = f->minPt;
    g->rMax = f->rMax;
    g->zMax = f->zMax;
    g->zMin = f->zMin;
    g->deltaRMax = f->deltaRMax;
    g->cotThetaMax = f->cotThetaMax;
    g->impactMax = f->impactMax;
    g->phiMin = f->phiMin;
    g->phiMax = f->phiMax;
    g->phiBinDeflectionCoverage = f->phiBinDeflectionCoverage;
}

// seedfilter_config — seeding_config.hpp:191-219
void b200seed_filter_cfg_defaults(b200seed_filter_cfg* c) {
    std::memset(c, 0, sizeof(*c));
    c->deltaInvHelixDiameter = 0.00003f / unit_mm;
    c->impactWeightFactor = 1.f;
    c->compatSeedWeight = 200.f;
    c->deltaRMin = 5.f * unit_mm;
    c->compatSeedLimit = 2;
    c->good_spB_min_radius = 150.f * unit_mm;
    c->good_spB_weight_increase = 400.f;
    c->good_spT_max_radius = 150.f * unit_mm;
    c->good_spT_weight_increase = 200.f;
    c->good_spB_min_weight = 380.f;
    c->seed_min_weight = 200.f;
    c->spB_min_radius = 43.f * unit_mm;
}

// track_params_estimation_config — detail/track_params_estimation_config.hpp:18-33
void b200seed_tpe_cfg_defaults(b200seed_tpe_cfg* c) {
    const float s[6] = {1.f * unit_mm,     1.f * unit_mm,        1.f * unit_degree,
                        1.f * unit_degree, 0.f * 1.f / unit_GeV, 1.f * unit_ns};
    const float infl[6] = {1.f, 1.f, 1.f, 1.f, 1.f, 100.f};
    for (int i = 0; i < 6; ++i) {
        c->initial_sigma[i] = s[i];
        c->initial_inflation[i] = infl[i];
    }
    c->initial_sigma_qopt = 0.1f * 1.f / unit_GeV;
    c->initial_sigma_pt_rel = 0.1f;
}

int b200seed_create(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                    const b200seed_filter_cfg* filter, const b200seed_tpe_cfg* tpe, int device,
                    b200seed_handle** out) {
    if (!finder || !grid || !filter || !out)
        return fail(nullptr, B200SEED_EINVAL, "b200seed_create: null argument");
    *out = nullptr;
    DevCfg d{};
    std::string why;
    if (compute_axes(*grid, d, why) != 0) return fail(nullptr, B200SEED_EINVAL, why);
    fill_devcfg(*finder, *filter, d);
    const uint64_t nbins = uint64_t(d.nPhi) * d.nZ;
    if (d.nPhi == 0 || nbins > 8192)
        return fail(nullptr, B200SEED_EINVAL,
                    "unsupported grid: " + std::to_string(d.nPhi) + " x " + std::to_string(d.nZ) +
                        " bins (limit 8192)");
    if (finder->maxSeedsPerSpM > uint32_t(MAX_TOPK))
        return fail(nullptr, B200SEED_EINVAL, "maxSeedsPerSpM > 16 is not supported");
    if (finder->maxSeedsPerSpM == 0)
        return fail(nullptr, B200SEED_EINVAL, "maxSeedsPerSpM == 0: no seed could ever be kept");
    if (filter->compatSeedLimit > size_t(MAX_COMPAT))
        return fail(nullptr, B200SEED_EINVAL, "compatSeedLimit > 8 is not supported");
    if (!(finder->deltaRMin >= 0.f))
        return fail(nullptr, B200SEED_EINVAL, "deltaRMin must be >= 0");
    if (d.scope0 + d.scope1 + 1u > d.nPhi)
        return fail(nullptr, B200SEED_EINVAL, "neighbor_scope wider than the phi axis");

    // The product path needs a CUDA device: fail loudly, never fall back to the CPU.
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, B200SEED_ECUDA,
                    std::string("no CUDA device available: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev)
        return fail(nullptr, B200SEED_EINVAL, "invalid device ordinal");
    e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return fail(nullptr, B200SEED_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));

    b200seed_handle* h = new b200seed_handle();
    h->finder = *finder;
    h->grid = *grid;
    h->filter = *filter;
    if (tpe)
        h->tpe = *tpe;
    else
        b200seed_tpe_cfg_defaults(&h->tpe);
    h->dev = d;
    h->device = device;
    h->nbins = uint32_t(nbins);
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    // allow the large dynamic shared memory configurations
    // (static shared memory counts against the same limit, hence the margin)
    cudaFuncSetAttribute(k_doublets<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    cudaFuncSetAttribute(k_doublets<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    cudaFuncSetAttribute(k_doublets<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    cudaFuncSetAttribute(k_doublets<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    cudaFuncSetAttribute(k_doublets_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    cudaFuncSetAttribute(k_doublets_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    if (const char* m = std::getenv("B200SEED_PCIE_PARAMS")) {
        h->pcie_compact = !std::strcmp(m, "compact");
        h->pcie_packed = !std::strcmp(m, "packed");
    }
    if (const char* m = std::getenv("B200SEED_DOUBLET_SIDES"))
        h->split_sides = !std::strcmp(m, "0") ? 0 : (!std::strcmp(m, "force") ? 2 : 1);
    if (const char* m = std::getenv("B200SEED_PDL")) h->pdl = std::strcmp(m, "0") != 0;
    if (const char* m = std::getenv("B200SEED_DOUBLET_ORDER"))
        h->ordered_tickets = !std::strcmp(m, "grid") ? 0 : (!std::strcmp(m, "cost") ? 2 : 1);
    if (const char* m = std::getenv("B200SEED_DOUBLETS")) {
        if (!std::strcmp(m, "tile")) h->doublet_mode = 0;
        else if (!std::strcmp(m, "ldgsts")) h->doublet_mode = 1;
        else h->doublet_mode = 2;  // "warp" / "legacy"
    }
    if (const char* m = std::getenv("B200SEED_TRIPLETS")) {
        h->triplet_pool = std::strcmp(m, "pool") == 0;
        h->triplet_lanes = std::strcmp(m, "lanes") == 0;
    }
    cudaFuncSetAttribute(k_triplets_pool, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    if (const char* m = std::getenv("B200SEED_GROUP_MAX")) h->group_max = uint32_t(std::atoi(m));
    if (const char* m = std::getenv("B200SEED_GROUP_ZSPAN")) h->group_zspan_mm = float(std::atof(m));
    cudaFuncSetAttribute(k_triplets<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    cudaFuncSetAttribute(k_triplets<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         h->smem_optin - 1024);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::string msg = std::string("kernel image not usable on this device (built for sm_100a): ") +
                          cudaGetErrorString(e);
        delete h;
        return fail(nullptr, B200SEED_ECUDA, msg);
    }
    if (cudaHostAlloc(reinterpret_cast<void**>(&h->h_sticky), 64, cudaHostAllocMapped | cudaHostAllocPortable) !=
        cudaSuccess) {
        delete h;
        return fail(nullptr, B200SEED_ECUDA, "cudaHostAlloc (overflow word) failed");
    }
    *h->h_sticky = 0u;
    *out = h;
    return B200SEED_OK;
}

void b200seed_destroy(b200seed_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto& s : h->slots) {
        cudaEventDestroy(s.start);
        cudaEventDestroy(s.stop);
    }
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_form) cudaFree(h->d_form);
    if (h->ev_host) cudaEventDestroy(h->ev_host);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    if (h->h_compact) cudaFreeHost(h->h_compact);
    if (h->h_sticky) cudaFreeHost(h->h_sticky);
    delete h;
}

const char* b200seed_last_error(const b200seed_handle* h) {
    return h ? h->error.c_str() : g_create_error.c_str();
}

int b200seed_get_axes(const b200seed_handle* h, uint32_t* n_phi, float* phi_min, float* phi_max,
                      uint32_t* n_z, float* z_min, float* z_max) {
    if (!h) return B200SEED_EINVAL;
    if (n_phi) *n_phi = h->dev.nPhi;
    if (phi_min) *phi_min = h->dev.phiAxisMin;
    if (phi_max) *phi_max = h->dev.phiAxisMax;
    if (n_z) *n_z = h->dev.nZ;
    if (z_min) *z_min = h->dev.zAxisMin;
    if (z_max) *z_max = h->dev.zAxisMax;
    return B200SEED_OK;
}

// Axes for a grid config without creating a handle (no GPU needed).
int b200seed_axes_for(const b200seed_grid_cfg* grid, uint32_t* n_phi, uint32_t* n_z) {
    DevCfg d{};
    std::string why;
    if (!grid || compute_axes(*grid, d, why) != 0) return fail(nullptr, B200SEED_EINVAL, why);
    *n_phi = d.nPhi;
    *n_z = d.nZ;
    return B200SEED_OK;
}

int b200seed_set_max_doublets(b200seed_handle* h, uint64_t max_doublets) {
    if (!h) return B200SEED_EINVAL;
    if (max_doublets > 0xFFFF0000ull) return fail(h, B200SEED_EINVAL, "max_doublets too large");
    h->max_doublets_user = max_doublets;
    return B200SEED_OK;
}

int b200seed_set_stage_cap(b200seed_handle* h, uint32_t cap) {
    if (!h) return B200SEED_EINVAL;
    if (cap != 0 && (cap < 16 || cap > 4096 || (cap & 15u)))
        return fail(h, B200SEED_EINVAL, "stage cap must be 0 or a multiple of 16 in [16, 4096]");
    // k_doublets needs doublet_smem_words(cap, cap / 2) words per warp
    if (cap != 0 && size_t(WARPS_PER_CTA) * doublet_smem_words(cap, cap / 2) * 4 + 1024 >
                        size_t(h->smem_optin > 0 ? h->smem_optin : 0))
        return fail(h, B200SEED_EINVAL,
                    "stage cap " + std::to_string(cap) + " needs more shared memory than the " +
                        std::to_string(h->smem_optin) + " bytes a CTA can have on this device");
    h->stage_cap_user = cap;
    return B200SEED_OK;
}

int b200seed_set_triplet_list_cap(b200seed_handle* h, uint32_t cap) {
    if (!h) return B200SEED_EINVAL;
    // (the per-warp shared-memory block must stay 16-byte aligned: 27 bytes per entry)
    if (cap != 0 && (cap < 16 || cap > 1024 || (cap & 15u)))
        return fail(h, B200SEED_EINVAL, "triplet list cap must be 0 or a multiple of 16 in [16, 1024]");
    if (cap != 0 && triplet_smem_per_warp(cap, true) * WARPS_PER_CTA + 1024 >
                        size_t(h->smem_optin > 0 ? h->smem_optin : 0))
        return fail(h, B200SEED_EINVAL, "triplet list cap needs more shared memory than a CTA can have");
    h->list_cap_user = cap;
    return B200SEED_OK;
}

int b200seed_set_triplet_dump(b200seed_handle* h, uint64_t max_triplets) {
    if (!h) return B200SEED_EINVAL;
    if (max_triplets > 0xFFFF0000ull) return fail(h, B200SEED_EINVAL, "max_triplets too large");
    h->max_dump = max_triplets;
    return B200SEED_OK;
}

size_t b200seed_workspace_bytes(const b200seed_handle* h, uint32_t max_spacepoints) {
    if (!h) return 0;
    return make_layout(h, max_spacepoints).total;
}

int b200seed_workspace_layout(const b200seed_handle* h, uint32_t max_spacepoints,
                              b200seed_ws_layout* out) {
    if (!h || !out) return B200SEED_EINVAL;
    const Layout L = make_layout(h, max_spacepoints);
    const size_t n = max_spacepoints ? max_spacepoints : 1;
    (void)n;
    out->bin_offsets = L.bin_off;
    out->sorted_index = L.sorted_index;
    out->sp_xyzr = L.sp4;
    out->mid_counts = L.cnt;
    out->mid_offsets = L.off;
    out->doublets = L.arena_b;
    out->triplet_dump = L.dump;
    out->triplet_dump_count = L.control + offsetof(Control, dump_cursor);
    out->max_doublets = L.max_doublets;
    out->max_triplet_dump = L.max_dump;
    out->n_bins = h->nbins;
    out->max_spacepoints = max_spacepoints;
    return B200SEED_OK;
}

int b200seed_set_timing(b200seed_handle* h, int enabled) {
    if (!h) return B200SEED_EINVAL;
    h->timing = enabled != 0;
    h->n_slots_used = 0;
    return B200SEED_OK;
}

int b200seed_get_timings(b200seed_handle* h, const char** names, float* ms, int cap) {
    if (!h) return B200SEED_EINVAL;
    int n = 0;
    for (int i = 0; i < h->n_slots_used && n < cap; ++i, ++n) {
        cudaError_t e = cudaEventSynchronize(h->slots[i].stop);
        if (e != cudaSuccess) return fail(h, B200SEED_ECUDA, cudaGetErrorString(e));
        float t = 0.f;
        cudaEventElapsedTime(&t, h->slots[i].start, h->slots[i].stop);
        if (names) names[n] = h->slots[i].name;
        if (ms) ms[n] = t;
    }
    return n;
}

int b200seed_launches_per_event(const b200seed_handle* h, int with_params) {
    // k_bin_count, k_cell_scan, k_bin_scatter, k_doublets<0>, k_doublets<1>, k_triplets,
    // k_seed_gather; the group kernel adds one launch (k_doublets_tile + k_doublets<2>)
    // (events of at least 20k spacepoints: + k_doublets<3> for the middles with a scarce side)
    const bool sides = h && h->doublet_mode == 2 && h->ordered_tickets && h->split_sides && h->finder.deltaRMin >= 0.f;
    const int doublets = (h && h->doublet_mode != 2) ? 3 : (sides ? 4 : 2);
    const int triplets = (h && (h->triplet_pool || h->triplet_lanes)) ? 2 : 1;  // (lanes: not for > 80k spacepoints)
    return 4 + doublets + triplets + (with_params ? 1 : 0);
}

}  // extern "C"

namespace {
// b200seed_run / b200seed_run_n_on_device: n_sp is the number of spacepoints, or — when
// d_n_sp != nullptr — an upper bound of the number stored in *d_n_sp on the device.
int run_impl(b200seed_handle* h, void* stream, uint32_t n_sp, const uint32_t* d_n_sp,
             const float* d_xyz, const float* d_var_z, const float* d_var_r, void* d_workspace,
             size_t workspace_bytes, uint32_t seed_capacity, uint32_t* d_bottom,
             uint32_t* d_middle, uint32_t* d_top, float* d_quality, uint32_t* d_n_seeds,
             b200seed_counters* d_counters) {
    if (!h) return B200SEED_EINVAL;
    if (!d_n_seeds) return fail(h, B200SEED_EINVAL, "b200seed_run: d_n_seeds is null");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->timing) h->n_slots_used = 0;
    if (n_sp == 0) {
        // "If there are no spacepoints, return right away" — triplet_seeding_algorithm.cpp:75-77
        CUDA_TRY(h, cudaMemsetAsync(d_n_seeds, 0, sizeof(uint32_t), s));
        if (d_counters) CUDA_TRY(h, cudaMemsetAsync(d_counters, 0, sizeof(b200seed_counters), s));
        return B200SEED_OK;
    }
    if (!d_xyz || !d_workspace || (seed_capacity && (!d_bottom || !d_middle || !d_top || !d_quality)))
        return fail(h, B200SEED_EINVAL, "b200seed_run: null device pointer");
    if (reinterpret_cast<uintptr_t>(d_workspace) % 256 != 0)
        return fail(h, B200SEED_EINVAL, "b200seed_run: workspace must be 256-byte aligned");
    const Layout L = make_layout(h, n_sp);
    if (workspace_bytes < L.total)
        return fail(h, B200SEED_ENOMEM,
                    "b200seed_run: workspace of " + std::to_string(workspace_bytes) +
                        " bytes, need " + std::to_string(L.total));
    unsigned char* ws = static_cast<unsigned char*>(d_workspace);
    auto at = [&](size_t off) { return ws + off; };
    Control* ctrl = reinterpret_cast<Control*>(at(L.control));
    uint32_t* bin_of = reinterpret_cast<uint32_t*>(at(L.bin_of));
    uint32_t* blk_hist = reinterpret_cast<uint32_t*>(at(L.blk_hist));
    uint32_t* bin_tot = reinterpret_cast<uint32_t*>(at(L.bin_tot));
    uint32_t* bin_off = reinterpret_cast<uint32_t*>(at(L.bin_off));
    uint32_t* sorted_index = reinterpret_cast<uint32_t*>(at(L.sorted_index));
    uint32_t* sorted_bin = reinterpret_cast<uint32_t*>(at(L.sorted_bin));
    float4* sp4 = reinterpret_cast<float4*>(at(L.sp4));
    float2* var2 = reinterpret_cast<float2*>(at(L.var2));
    uint32_t* cnt = reinterpret_cast<uint32_t*>(at(L.cnt));
    uint32_t* off = reinterpret_cast<uint32_t*>(at(L.off));
    uint32_t* seed_cnt = reinterpret_cast<uint32_t*>(at(L.seed_cnt));
    uint32_t* seed_b = reinterpret_cast<uint32_t*>(at(L.seed_b));
    uint32_t* seed_t = reinterpret_cast<uint32_t*>(at(L.seed_t));
    float* seed_w = reinterpret_cast<float*>(at(L.seed_w));
    const uint32_t nblk = L.nblk;
    const uint32_t K = h->finder.maxSeedsPerSpM;

    uint32_t* cell_cnt = reinterpret_cast<uint32_t*>(at(L.cell_cnt));
    uint32_t* cell_off = reinterpret_cast<uint32_t*>(at(L.cell_off));
    float4* csp4 = reinterpret_cast<float4*>(at(L.csp4));
    uint32_t* ccanon = reinterpret_cast<uint32_t*>(at(L.ccanon));
    // ticket order of k_doublets<0> (cost classes per (bin, r row); B200SEED_DOUBLET_ORDER=grid: off)
    // (from 8k spacepoints on: below that the launch is too short for its end to matter; =cost
    // forces it)
    const bool ordered = h->ordered_tickets == 2 || (h->ordered_tickets == 1 && n_sp >= 8192u);
    uint32_t* seg_info = ordered ? reinterpret_cast<uint32_t*>(at(L.seg_info)) : nullptr;
    uint32_t* mid_order = ordered ? reinterpret_cast<uint32_t*>(at(L.mid_order)) : nullptr;
    // the classes with an (almost) empty side get their own launch, which looks at that side first
    // (k_doublets<3>; needs deltaRMin >= 0: bottoms below, tops above the middle's row)
    // From 20k spacepoints on: a batch of 32 lane-serial pre-screenings takes 35-60 us, which only
    // hides behind a main launch that long (2000 particles per event: doublet stage 36 -> 63 us with
    // it, 5000: 70 -> 63 us, 10000: 148 -> 130 us).
    // (B200SEED_DOUBLET_ORDER=cost or B200SEED_DOUBLET_SIDES=force: at any size, for the tests)
    const bool split_sides = ordered && h->split_sides && h->finder.deltaRMin >= 0.f && h->doublet_mode == 2 &&
                             (h->split_sides == 2 || h->ordered_tickets == 2 || n_sp >= 20000u);
    uint32_t row_reach = uint32_t(h->finder.deltaRMax * L.g.invRw) + 1u;  // rows a partner can be away
    if (!(h->finder.deltaRMax >= 0.f) || row_reach > 31u) row_reach = 31u;
    // canon_key (seed_kernels.cuh) is a 32-bit word
    if (uint64_t(h->dev.scope0 + h->dev.scope1 + 1u) * n_sp > 0xFFFFFFFFull)
        return fail(h, B200SEED_EINVAL, "b200seed_run: too many spacepoints for this neighbor_scope");
    // k_triplets<DENSE> packs (row, mid-top index) into one word: 27 bits for the index
    if (n_sp >= (1u << 27)) return fail(h, B200SEED_EINVAL, "b200seed_run: more than 2^27 spacepoints");

    CUDA_TRY(h, cudaMemsetAsync(ws, 0, L.zero_bytes, s));
    {
        KernelTimer t(h, s, "bin_count");
        k_bin_count<<<nblk, BIN_THREADS, h->nbins * sizeof(uint32_t), s>>>(
            h->dev, L.g, n_sp, d_xyz, bin_of, blk_hist, cell_cnt, h->nbins, nblk, d_n_sp, bin_tot);
    }
    {
        KernelTimer t(h, s, "cell_scan");
        // groups of neighbouring middles for k_doublets_tile: as many members as its mid-top
        // segments hold for this occupancy (~7e-4 N mid-tops per active middle, with headroom)
        uint32_t gmax = h->group_max ? h->group_max : uint32_t(double(TILE_QT) * 1500.0 / double(n_sp));
        gmax = gmax < 1u ? 1u : (gmax > TILE_GCAP ? TILE_GCAP : gmax);
        const float zspan_mm = h->group_zspan_mm > 0.f ? h->group_zspan_mm : 64.f;
        uint32_t zspan = uint32_t(zspan_mm * L.g.invZw + 0.5f);
        if (zspan < 1u) zspan = 1u;
        k_cell_scan<<<h->nbins, 256, h->doublet_mode == 2 ? 0 : (L.g.CPB + 1) * sizeof(uint32_t), s>>>(
            cell_cnt, cell_off, bin_off, bin_tot, blk_hist, nblk, L.g.CPB, h->nbins,
            h->doublet_mode == 2 ? nullptr : reinterpret_cast<uint32_t*>(at(L.group_list)), ctrl,
            L.g.NZc, gmax, zspan, gmax >= 4u ? gmax / 2u : 2u, n_sp, seg_info, row_reach, split_sides);
    }
    {
        KernelTimer t(h, s, "bin_scatter");
        k_bin_scatter<<<nblk, BIN_THREADS, h->nbins * sizeof(uint32_t), s>>>(
            h->dev, L.g, n_sp, d_xyz, d_var_z, d_var_r, bin_of, blk_hist, nblk, sp4, var2, sorted_index,
            sorted_bin, cell_off, cell_cnt, csp4, ccanon, d_n_sp, ctrl, h->nbins, seg_info, mid_order);
    }
    {
        DoubletArgs a{};
        a.bin_off = bin_off;
        a.sp4 = sp4;
        a.var2 = var2;
        a.sorted_bin = sorted_bin;
        a.cell_off = cell_off;
        a.csp4 = csp4;
        a.ccanon = ccanon;
        a.cnt_b = cnt;
        a.cnt_t = cnt + n_sp;
        a.off_b = off;
        a.off_t = off + n_sp;
        a.arena_b = reinterpret_cast<DoubletRec*>(at(L.arena_b));
        a.arena_t = reinterpret_cast<DoubletRec*>(at(L.arena_t));
        a.ctrl = ctrl;
        a.g = L.g;
        a.max_doublets = uint32_t(L.max_doublets);
        a.cap_b = h->stage_cap_user ? h->stage_cap_user : doublet_stage_cap(n_sp);
        a.cap_t = a.cap_b / 2;
        const size_t smem = size_t(WARPS_PER_CTA) * doublet_smem_words(a.cap_b, a.cap_t) * 4;
        uint32_t grid = (n_sp + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        // One wave: as many CTAs as can be resident (the warps draw tickets, so more CTAs only
        // start, find the queue empty and leave — with several events in flight they take slots
        // other events' kernels could use: 2x the resident CTAs cost 2.6 % of the event rate).
#ifndef B200_DOUBLET_GRID_X2
#define B200_DOUBLET_GRID_X2 2
#endif
        const uint32_t max_grid = uint32_t(h->num_sms) * (32 / WARPS_PER_CTA) * B200_DOUBLET_GRID_X2 / 2;
        if (grid > max_grid) grid = max_grid;
        KernelTimer t(h, s, "doublets");
        a.spill_list = reinterpret_cast<uint32_t*>(at(L.spill_list));
        a.active_list = reinterpret_cast<uint32_t*>(at(L.active_list));
        a.seed_cnt = seed_cnt;
        a.n_sp = n_sp;
        a.fallback_list = reinterpret_cast<const uint32_t*>(at(L.fallback_list));
        a.mid_order = mid_order;
        a.split_sides = split_sides ? 1u : 0u;
        const uint32_t grid_s = grid < uint32_t(h->num_sms) * 4u ? grid : uint32_t(h->num_sms) * 4u;
        if (h->doublet_mode == 2) {
            k_doublets<0><<<grid, WARPS_PER_CTA * 32, smem, s>>>(h->dev, a);
            if (split_sides) {
                // no data dependence on k_doublets<0> (disjoint middles, shared atomics only): its
                // CTAs start as soon as that launch frees slots, i.e. they fill its tail
                cudaLaunchConfig_t lc{};
                // a warp takes 32 middles at a time: no more CTAs than there can be batches
                const uint32_t grid_e = n_sp / (32u * WARPS_PER_CTA) + 1u;
                lc.gridDim = dim3(grid_e < grid_s ? grid_e : grid_s);
                lc.blockDim = dim3(WARPS_PER_CTA * 32);
                lc.dynamicSmemBytes = smem;
                lc.stream = s;
                cudaLaunchAttribute at1{};
                at1.id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at1.val.programmaticStreamSerializationAllowed = h->pdl ? 1 : 0;
                lc.attrs = &at1;
                lc.numAttrs = 1;
                cudaError_t le = cudaLaunchKernelEx(&lc, k_doublets<3>, h->dev, a);
                if (le != cudaSuccess && h->pdl) {
                    // a driver without programmatic dependent launch: plain stream order from now on
                    (void)cudaGetLastError();
                    h->pdl = 0;
                    at1.val.programmaticStreamSerializationAllowed = 0;
                    le = cudaLaunchKernelEx(&lc, k_doublets<3>, h->dev, a);
                }
                CUDA_TRY(h, le);
                // survivors beyond SIDED_INLINE per batch (normally none: the CTAs find an empty
                // list and exit): warp per middle
                const uint32_t grid_f = grid_s < uint32_t(h->num_sms) ? grid_s : uint32_t(h->num_sms);
                k_doublets<2><<<grid_f, WARPS_PER_CTA * 32, smem, s>>>(h->dev, a);
            }
        } else {
            TileArgs ta{};
            ta.bin_off = bin_off;
            ta.sorted_bin = sorted_bin;
            ta.var2 = var2;
            ta.cell_off = cell_off;
            ta.csp4 = csp4;
            ta.ccanon = ccanon;
            ta.group_list = reinterpret_cast<const uint32_t*>(at(L.group_list));
            ta.cnt_b = a.cnt_b, ta.cnt_t = a.cnt_t, ta.off_b = a.off_b, ta.off_t = a.off_t;
            ta.arena_b = a.arena_b, ta.arena_t = a.arena_t;
            ta.ctrl = ctrl;
            ta.g = L.g;
            ta.max_doublets = a.max_doublets;
            ta.fallback_list = reinterpret_cast<uint32_t*>(at(L.fallback_list));
            ta.active_list = a.active_list;
            ta.seed_cnt = seed_cnt;
            ta.n_sp = n_sp;
            const size_t tsmem = size_t(TILE_WARPS) * sizeof(TileWarp);
            uint32_t tgrid = uint32_t(h->num_sms) * B200_TILE_MIN_CTAS;
            const uint32_t need = (n_sp + TILE_WARPS - 1) / TILE_WARPS;
            if (tgrid > need) tgrid = need;
            if (h->doublet_mode == 1)
                k_doublets_tile<1><<<tgrid, TILE_WARPS * 32, tsmem, s>>>(h->dev, ta);
            else
                k_doublets_tile<0><<<tgrid, TILE_WARPS * 32, tsmem, s>>>(h->dev, ta);
            // groups it handed back (doublets that outgrow its queues): warp per middle
            k_doublets<2><<<grid_s, WARPS_PER_CTA * 32, smem, s>>>(h->dev, a);
        }
        // middles whose lists outgrew the staging area (none for ordinary events: the CTAs
        // find an empty list and exit)
        k_doublets<1><<<grid_s, WARPS_PER_CTA * 32, smem, s>>>(h->dev, a);
    }
    TripletArgs ta{};  // also read by k_seed_gather (slow path of the triplet search)
    {
        TripletArgs& a = ta;
        a.sp4 = sp4;
        a.var2 = var2;
        a.sorted_bin = sorted_bin;
        a.cnt_b = cnt;
        a.cnt_t = cnt + n_sp;
        a.off_b = off;
        a.off_t = off + n_sp;
        a.arena_b = reinterpret_cast<DoubletRec*>(at(L.arena_b));
        a.arena_t = reinterpret_cast<DoubletRec*>(at(L.arena_t));
        a.ctrl = ctrl;
        a.seed_cnt = seed_cnt;
        a.seed_b = seed_b;
        a.seed_t = seed_t;
        a.seed_w = seed_w;
        a.dump = L.max_dump ? reinterpret_cast<TripletDumpRec*>(at(L.dump)) : nullptr;
        a.max_dump = uint32_t(L.max_dump);
        a.list_cap = h->list_cap_user ? h->list_cap_user : triplet_list_cap(n_sp);
        a.scratch_t = reinterpret_cast<DoubletRec*>(at(L.arena_t));
        a.max_doublets = uint32_t(L.max_doublets);
        a.slow_list = reinterpret_cast<uint32_t*>(at(L.slow_list));
        a.active_list = reinterpret_cast<const uint32_t*>(at(L.active_list));
        a.n_sp = n_sp;
        const bool dense = n_sp > 80000u;
        const size_t smem = triplet_smem_per_warp(a.list_cap, dense) * WARPS_PER_CTA;
        uint32_t grid = (n_sp + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
#ifndef B200_TRIPLET_GRID_X2
#define B200_TRIPLET_GRID_X2 2
#endif
        const uint32_t max_grid = uint32_t(h->num_sms) * (32 / WARPS_PER_CTA) * B200_TRIPLET_GRID_X2 / 2;
        if (grid > max_grid) grid = max_grid;
        KernelTimer t(h, s, "triplets");
        const bool lanes = h->triplet_lanes && !h->triplet_pool && !dense;
        a.heavy_only = h->triplet_pool ? WORK_HEAVY_CLASSES : (lanes ? LANES_FIRST_CLASS : 0u);
        if (dense)
            k_triplets<true><<<grid, WARPS_PER_CTA * 32, smem, s>>>(h->dev, a);
        else
            k_triplets<false><<<grid, WARPS_PER_CTA * 32, smem, s>>>(h->dev, a);
        if (h->triplet_pool) {
            // the light middles, POOL_G per warp
            const size_t psmem = pool_smem_per_warp(K) * POOL_WARPS;
            uint32_t pgrid = (n_sp / POOL_G + POOL_WARPS) / POOL_WARPS;
            const uint32_t pmax = uint32_t(h->num_sms) * B200_POOL_MIN_CTAS;
            if (pgrid > pmax) pgrid = pmax;
            k_triplets_pool<<<pgrid, POOL_WARPS * 32, psmem, s>>>(h->dev, a);
        }
        if (lanes) {
            // the light middles, one per lane; no data dependence on k_triplets (disjoint middles):
            // launched as a programmatic dependent, its CTAs fill that launch's tail
            cudaLaunchConfig_t lc{};
            uint32_t lgrid = (n_sp / 32u + LANES_WARPS) / LANES_WARPS;
            const uint32_t lmax = uint32_t(h->num_sms) * 4u;
            lc.gridDim = dim3(lgrid < lmax ? lgrid : lmax);
            lc.blockDim = dim3(LANES_WARPS * 32);
            lc.dynamicSmemBytes = lanes_smem_bytes();
            lc.stream = s;
            cudaLaunchAttribute at1{};
            at1.id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at1.val.programmaticStreamSerializationAllowed = h->pdl ? 1 : 0;
            lc.attrs = &at1;
            lc.numAttrs = 1;
            cudaError_t le = cudaLaunchKernelEx(&lc, k_triplets_lanes, h->dev, a);
            if (le != cudaSuccess && h->pdl) {
                (void)cudaGetLastError();
                h->pdl = 0;
                at1.val.programmaticStreamSerializationAllowed = 0;
                le = cudaLaunchKernelEx(&lc, k_triplets_lanes, h->dev, a);
            }
            CUDA_TRY(h, le);
        }


    }
    {
        // exclusive scan of the per-middle seed counts fused into the gather (single pass,
        // decoupled look-back): seeds come out in the reference CPU's order
        KernelTimer t(h, s, "seed_gather");
        unsigned long long* gs = reinterpret_cast<unsigned long long*>(at(L.gather_state));
        k_seed_gather<<<nblk, BIN_THREADS, 0, s>>>(
            n_sp, K, ctrl, seed_cnt, seed_b, seed_t, seed_w, sorted_index, seed_capacity, d_bottom,
            d_middle, d_top, d_quality, d_n_seeds, d_counters, d_n_sp, gs + 1, gs, h->h_sticky, h->dev, ta);
    }
    CUDA_TRY(h, cudaGetLastError());
    return B200SEED_OK;
}
}  // namespace

extern "C" {

int b200seed_run(b200seed_handle* h, void* stream, uint32_t n_sp, const float* d_xyz,
                 const float* d_var_z, const float* d_var_r, void* d_workspace,
                 size_t workspace_bytes, uint32_t seed_capacity, uint32_t* d_bottom,
                 uint32_t* d_middle, uint32_t* d_top, float* d_quality, uint32_t* d_n_seeds,
                 b200seed_counters* d_counters) {
    return run_impl(h, stream, n_sp, nullptr, d_xyz, d_var_z, d_var_r, d_workspace, workspace_bytes,
                    seed_capacity, d_bottom, d_middle, d_top, d_quality, d_n_seeds, d_counters);
}

int b200seed_run_n_on_device(b200seed_handle* h, void* stream, uint32_t max_sp,
                             const uint32_t* d_n_sp, const float* d_xyz, const float* d_var_z,
                             const float* d_var_r, void* d_workspace, size_t workspace_bytes,
                             uint32_t seed_capacity, uint32_t* d_bottom, uint32_t* d_middle,
                             uint32_t* d_top, float* d_quality, uint32_t* d_n_seeds,
                             b200seed_counters* d_counters) {
    if (h && !d_n_sp) return fail(h, B200SEED_EINVAL, "b200seed_run_n_on_device: d_n_sp is null");
    return run_impl(h, stream, max_sp, d_n_sp, d_xyz, d_var_z, d_var_r, d_workspace, workspace_bytes,
                    seed_capacity, d_bottom, d_middle, d_top, d_quality, d_n_seeds, d_counters);
}

int b200seed_form_spacepoints(b200seed_handle* h, void* stream, uint32_t n_meas,
                              const float* d_meas_local, const uint32_t* d_meas_dim,
                              const uint32_t* d_meas_surface_index,
                              const b200seed_surface* d_surfaces, uint32_t n_surfaces, float* d_xyz,
                              float* d_var_z, float* d_var_r, uint32_t* d_meas_index_1,
                              uint32_t* d_meas_index_2, uint32_t* d_n_sp) {
    if (!h) return B200SEED_EINVAL;
    if (!d_n_sp) return fail(h, B200SEED_EINVAL, "b200seed_form_spacepoints: d_n_sp is null");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (n_meas == 0) {
        // "If there are no measurements, return right away"
        // (silicon_pixel_spacepoint_formation_algorithm.cpp:37-40)
        CUDA_TRY(h, cudaMemsetAsync(d_n_sp, 0, sizeof(uint32_t), s));
        return B200SEED_OK;
    }
    if (!d_meas_local || !d_meas_surface_index || !d_surfaces || !d_xyz)
        return fail(h, B200SEED_EINVAL, "b200seed_form_spacepoints: null device pointer");
    const size_t tiles = (size_t(n_meas) + FORM_TILE - 1) / FORM_TILE;
    if (tiles > h->form_tiles) {
        // grows rarely; a fresh buffer is zero == "epoch 0", which no call ever uses
        CUDA_TRY(h, cudaStreamSynchronize(s));
        if (h->d_form) CUDA_TRY(h, cudaFree(h->d_form));
        h->d_form = nullptr;
        h->form_tiles = 0;
        const size_t want = tiles + tiles / 2 + 64;
        CUDA_TRY(h, cudaMalloc(&h->d_form, (want + 1) * sizeof(unsigned long long)));
        CUDA_TRY(h, cudaMemset(h->d_form, 0, (want + 1) * sizeof(unsigned long long)));
        h->form_tiles = want;
        h->form_ticket_base = 0;
        h->form_epoch = 0;
    }
    if (++h->form_epoch >= (1u << 30)) {  // epoch field is 30 bits wide: start over
        CUDA_TRY(h, cudaMemsetAsync(h->d_form + 1, 0, h->form_tiles * sizeof(unsigned long long), s));
        h->form_epoch = 1;
    }
    {
        KernelTimer t(h, s, "form_spacepoints");
        k_form_spacepoints<<<uint32_t(tiles), FORM_THREADS, 0, s>>>(
            n_meas, d_meas_local, d_meas_dim, d_meas_surface_index, d_surfaces, n_surfaces, d_xyz,
            d_var_z, d_var_r, d_meas_index_1, d_meas_index_2, d_n_sp, h->d_form + 1, h->d_form,
            h->form_ticket_base, h->form_epoch);
    }
    h->form_ticket_base += tiles;
    CUDA_TRY(h, cudaGetLastError());
    return B200SEED_OK;
}

}  // extern "C"

namespace {
int estimate_impl(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                  uint32_t seed_capacity, const uint32_t* d_bottom, const uint32_t* d_middle,
                  const uint32_t* d_top, const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                  const float* d_meas_local, const uint64_t* d_meas_surface, const float bfield[3],
                  const b200seed_field_grid& fg, b200seed_bound_params* d_params,
                  b200seed_bound_params_diag* d_params_diag = nullptr,
                  b200seed_seed_params* d_params_compact = nullptr,
                  b200seed_bound_params_packed* d_params_packed = nullptr) {
    if (!h) return B200SEED_EINVAL;
    if (seed_capacity == 0) return B200SEED_OK;
    if (!d_xyz) return B200SEED_OK;  // no spacepoints => no seeds (…estimation_algorithm.cpp:49-51)
    if (!d_n_seeds || !d_bottom || !d_middle || !d_top || !bfield ||
        (!d_params && !d_params_diag && !d_params_compact && !d_params_packed))
        return fail(h, B200SEED_EINVAL, "b200seed_estimate_params: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUDA_TRY(h, cudaSetDevice(h->device));
    {
        KernelTimer t(h, s, "estimate_params");
        k_estimate_params<<<(seed_capacity + 127) / 128, 128, 0, s>>>(
            h->tpe, d_n_seeds, seed_capacity, d_bottom, d_middle, d_top, d_xyz, d_sp_meas_index_1,
            d_meas_local, d_meas_surface, bfield[0], bfield[1], bfield[2], fg, d_params, d_params_diag,
            d_params_compact, d_params_packed);
    }
    CUDA_TRY(h, cudaGetLastError());
    return B200SEED_OK;
}
}  // namespace

extern "C" {

int b200seed_estimate_params(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                             uint32_t seed_capacity, const uint32_t* d_bottom,
                             const uint32_t* d_middle, const uint32_t* d_top,
                             const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                             const float* d_meas_local, const uint64_t* d_meas_surface,
                             const float bfield[3], b200seed_bound_params* d_params) {
    b200seed_field_grid none{};
    return estimate_impl(h, stream, d_n_seeds, seed_capacity, d_bottom, d_middle, d_top, d_xyz,
                         d_sp_meas_index_1, d_meas_local, d_meas_surface, bfield, none, d_params);
}

int b200seed_estimate_params_diag(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                                  uint32_t seed_capacity, const uint32_t* d_bottom,
                                  const uint32_t* d_middle, const uint32_t* d_top,
                                  const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                                  const float* d_meas_local, const uint64_t* d_meas_surface,
                                  const float bfield[3], b200seed_bound_params_diag* d_params) {
    b200seed_field_grid none{};
    return estimate_impl(h, stream, d_n_seeds, seed_capacity, d_bottom, d_middle, d_top, d_xyz,
                         d_sp_meas_index_1, d_meas_local, d_meas_surface, bfield, none, nullptr, d_params);
}

int b200seed_estimate_params_compact(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                                     uint32_t seed_capacity, const uint32_t* d_bottom,
                                     const uint32_t* d_middle, const uint32_t* d_top,
                                     const float* d_xyz, const float bfield[3],
                                     b200seed_seed_params* d_params) {
    b200seed_field_grid none{};
    return estimate_impl(h, stream, d_n_seeds, seed_capacity, d_bottom, d_middle, d_top, d_xyz, nullptr,
                         nullptr, nullptr, bfield, none, nullptr, nullptr, d_params);
}

int b200seed_estimate_params_packed(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                                    uint32_t seed_capacity, const uint32_t* d_bottom,
                                    const uint32_t* d_middle, const uint32_t* d_top,
                                    const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                                    const float* d_meas_local, const uint64_t* d_meas_surface,
                                    const float bfield[3], b200seed_bound_params_packed* d_params) {
    b200seed_field_grid none{};
    return estimate_impl(h, stream, d_n_seeds, seed_capacity, d_bottom, d_middle, d_top, d_xyz,
                         d_sp_meas_index_1, d_meas_local, d_meas_surface, bfield, none, nullptr, nullptr,
                         nullptr, d_params);
}

void b200seed_expand_packed_params(const b200seed_handle* h, uint32_t n,
                                   const b200seed_bound_params_packed* in,
                                   b200seed_bound_params* out_full, b200seed_bound_params_diag* out_diag) {
    if (!h || !in) return;
    // the variances that do not depend on the seed: the same two float multiplications as in
    // k_estimate_params (track_params_estimation.cpp:64-86)
    float var[6];
    for (int j = 0; j < 6; ++j) {
        float v = h->tpe.initial_sigma[j] * h->tpe.initial_sigma[j];
        v *= h->tpe.initial_inflation[j];
        var[j] = v;
    }
    for (uint32_t i = 0; i < n; ++i) {
        const b200seed_bound_params_packed p = in[i];
        const float vec[6] = {p.loc0, p.loc1, p.phi, p.theta, p.qop, 0.f};
        if (out_diag) {
            b200seed_bound_params_diag& o = out_diag[i];
            o.surface_link = p.surface_link;
            for (int k = 0; k < 6; ++k) o.vec[k] = vec[k], o.cov_diag[k] = (k == 4) ? p.var_qop : var[k];
        }
        if (out_full) {
            b200seed_bound_params& o = out_full[i];
            std::memset(&o, 0, sizeof(o));
            o.surface_link = p.surface_link;
            for (int k = 0; k < 6; ++k) o.vec[k] = vec[k], o.cov[k * 7] = (k == 4) ? p.var_qop : var[k];
        }
    }
}

void b200seed_expand_seed_params(const b200seed_handle* h, uint32_t n, const uint32_t* bottom,
                                 const b200seed_seed_params* in, const uint32_t* sp_meas_index_1,
                                 const float* meas_local, const uint64_t* meas_surface,
                                 b200seed_bound_params* out_full, b200seed_bound_params_diag* out_diag) {
    if (!h || !bottom || !in) return;
    // the variances that do not depend on the seed: the same two float multiplications as in
    // k_estimate_params (track_params_estimation.cpp:64-86)
    float var[6];
    for (int j = 0; j < 6; ++j) {
        float v = h->tpe.initial_sigma[j] * h->tpe.initial_sigma[j];
        v *= h->tpe.initial_inflation[j];
        var[j] = v;
    }
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t ib = bottom[i];
        const uint32_t mi = sp_meas_index_1 ? sp_meas_index_1[ib] : ib;
        const uint64_t link = meas_surface ? meas_surface[mi] : 0ull;
        const float loc0 = meas_local ? meas_local[2 * size_t(mi)] : 0.f;
        const float loc1 = meas_local ? meas_local[2 * size_t(mi) + 1] : 0.f;
        const float vec[6] = {loc0, loc1, in[i].phi, in[i].theta, in[i].qop, 0.f};
        if (out_diag) {
            b200seed_bound_params_diag& o = out_diag[i];
            o.surface_link = link;
            for (int k = 0; k < 6; ++k) o.vec[k] = vec[k], o.cov_diag[k] = (k == 4) ? in[i].var_qop : var[k];
        }
        if (out_full) {
            b200seed_bound_params& o = out_full[i];
            std::memset(&o, 0, sizeof(o));
            o.surface_link = link;
            for (int k = 0; k < 6; ++k) o.vec[k] = vec[k], o.cov[k * 7] = (k == 4) ? in[i].var_qop : var[k];
        }
    }
}

void b200seed_expand_params(const b200seed_bound_params_diag* in, uint32_t n, b200seed_bound_params* out) {
    for (uint32_t i = 0; i < n; ++i) {
        b200seed_bound_params& o = out[i];
        std::memset(&o, 0, sizeof(o));
        o.surface_link = in[i].surface_link;
        for (int k = 0; k < 6; ++k) {
            o.vec[k] = in[i].vec[k];
            o.cov[k * 7] = in[i].cov_diag[k];
        }
    }
}

int b200seed_estimate_params_inhom(b200seed_handle* h, void* stream, const uint32_t* d_n_seeds,
                                   uint32_t seed_capacity, const uint32_t* d_bottom,
                                   const uint32_t* d_middle, const uint32_t* d_top,
                                   const float* d_xyz, const uint32_t* d_sp_meas_index_1,
                                   const float* d_meas_local, const uint64_t* d_meas_surface,
                                   const b200seed_field_grid* field,
                                   b200seed_bound_params* d_params) {
    if (!h) return B200SEED_EINVAL;
    if (!field || !field->data || !field->size[0] || !field->size[1] || !field->size[2])
        return fail(h, B200SEED_EINVAL, "b200seed_estimate_params_inhom: empty field grid");
    const float unused[3] = {0.f, 0.f, 0.f};
    return estimate_impl(h, stream, d_n_seeds, seed_capacity, d_bottom, d_middle, d_top, d_xyz,
                         d_sp_meas_index_1, d_meas_local, d_meas_surface, unused, *field, d_params);
}

}  // extern "C"

namespace {

// One event of the host-buffer path, split in two so that callers can keep several events
// in flight: submit = H->D + all kernels + read-back of the counters (asynchronous);
// finish = wait for the counters, sized D->H copies of seeds and parameters, wait.
struct HostEvent {
    uint32_t n_sp = 0, n_meas = 0, seed_capacity = 0;
    const float* h_xyz = nullptr;
    const float* h_var_z = nullptr;
    const float* h_var_r = nullptr;
    const uint32_t* h_smi = nullptr;
    const float* h_ml = nullptr;
    const uint64_t* h_ms = nullptr;
    float bfield[3] = {0.f, 0.f, 0.f};
    uint32_t* h_bottom = nullptr;
    uint32_t* h_middle = nullptr;
    uint32_t* h_top = nullptr;
    float* h_quality = nullptr;
    b200seed_bound_params* h_params = nullptr;
    b200seed_bound_params_diag* h_params_diag = nullptr;  // when set, the parameters cross PCIe as
                                                          // 56-byte diagonal records
    b200seed_bound_params_packed* h_params_packed = nullptr;  // ... as 32-byte packed records, delivered as such
    // device staging of the outputs (set by submit)
    uint32_t *d_b = nullptr, *d_m = nullptr, *d_t = nullptr;
    float* d_q = nullptr;
    b200seed_bound_params* d_p = nullptr;
    bool compact = false;  // the parameters cross PCIe as b200seed_seed_params
    bool packed = false;   // ... as b200seed_bound_params_packed
    const uint32_t* h_bot_stage = nullptr;  // bottom indices on the host (for host_expand)
    bool submitted = false;
};

// The host threads of a throughput job wait on an event created with cudaEventBlockingSync:
// they sleep instead of spinning in cudaStreamSynchronize, so a node running one pool per GPU
// (8 ranks x several workers) does not oversubscribe its cores with busy-waiting threads.
cudaError_t host_wait_point(b200seed_handle* h, cudaStream_t s) {
    if (!h->ev_host) {
        const cudaError_t e =
            cudaEventCreateWithFlags(&h->ev_host, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    return cudaEventRecord(h->ev_host, s);
}

int host_submit(b200seed_handle* h, cudaStream_t s, HostEvent& e) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    e.submitted = false;
    if (e.n_sp == 0) return B200SEED_OK;
    if (!e.h_xyz) return fail(h, B200SEED_EINVAL, "b200seed_run_host: h_xyz is null");
    const bool diag = e.h_params_diag != nullptr;
    const bool want_params = e.h_params != nullptr || diag || e.h_params_packed != nullptr;
    const uint32_t n_sp = e.n_sp, n_meas = e.n_meas, seed_capacity = e.seed_capacity;
    // B200SEED_PCIE_PARAMS=compact: only phi, theta, q/p and var(q/p) are computed on the device
    // (16 bytes per seed); the records are completed on the host from the caller's own measurement
    // columns, which then never travel to the device. Less than half the PCIe bytes, but the host
    // pays ~0.9 ms of cache-cold gathers per 10k-particle event: on one GPU 3.67k instead of 3.76k
    // events/s (3.90k if the expansion cost nothing) — for hosts whose D->H rate is the limit, not
    // the default.
    // (read at b200seed_create)
    e.compact = want_params && h->pcie_compact && !e.h_params_packed;  // (a packed delivery is its own wire form)
    // B200SEED_PCIE_PARAMS=packed: 32-byte records (no constant variances), completed on the host by
    // a sequential copy (b200seed_expand_packed_params). Measured like the compact form: the host
    // writing the records costs more than the copy engine saving a third of the bytes — 3.90k vs
    // 4.00k events/s on one GPU, 13.0k vs 18.4k on eight (whose host is the limit either way).
    e.packed = want_params && !e.compact && (h->pcie_packed || e.h_params_packed != nullptr);

    // device staging: inputs | outputs | workspace
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    const size_t o_xyz = take(size_t(n_sp) * 12), o_vz = take(size_t(n_sp) * 4),
                 o_vr = take(size_t(n_sp) * 4), o_smi = take(size_t(n_sp) * 4),
                 o_ml = take(size_t(n_meas) * 8), o_ms = take(size_t(n_meas) * 8),
                 o_b = take(size_t(seed_capacity) * 4), o_m = take(size_t(seed_capacity) * 4),
                 o_t = take(size_t(seed_capacity) * 4), o_q = take(size_t(seed_capacity) * 4),
                 o_p = take(want_params ? size_t(seed_capacity) *
                                              (e.compact ? sizeof(b200seed_seed_params)
                                               : e.packed ? sizeof(b200seed_bound_params_packed)
                                                          : (diag ? sizeof(b200seed_bound_params_diag)
                                                                  : sizeof(b200seed_bound_params)))
                                        : 0),
                 o_n = take(256), o_c = take(sizeof(b200seed_counters));
    const size_t ws_bytes = b200seed_workspace_bytes(h, n_sp);
    const size_t o_ws = take(ws_bytes);
    if (o > h->d_stage_bytes) {
        if (h->d_stage) CUDA_TRY(h, cudaFree(h->d_stage));
        h->d_stage = nullptr;
        h->d_stage_bytes = 0;
        const size_t want = o + o / 4;
        CUDA_TRY(h, cudaMalloc(&h->d_stage, want));
        h->d_stage_bytes = want;
    }
    if (!h->h_pinned) CUDA_TRY(h, cudaMallocHost(&h->h_pinned, 256));
    if (e.compact || (e.packed && !e.h_params_packed)) {
        const size_t need = size_t(seed_capacity) * (e.packed ? sizeof(b200seed_bound_params_packed)
                                                              : sizeof(b200seed_seed_params) + 4);
        if (need > h->h_compact_bytes) {
            if (h->h_compact) CUDA_TRY(h, cudaFreeHost(h->h_compact));
            h->h_compact = nullptr;
            h->h_compact_bytes = 0;
            CUDA_TRY(h, cudaMallocHost(&h->h_compact, need + need / 4));
            h->h_compact_bytes = need + need / 4;
        }
    }
    unsigned char* d = static_cast<unsigned char*>(h->d_stage);
    float* d_xyz = reinterpret_cast<float*>(d + o_xyz);
    float* d_vz = e.h_var_z ? reinterpret_cast<float*>(d + o_vz) : nullptr;
    float* d_vr = e.h_var_r ? reinterpret_cast<float*>(d + o_vr) : nullptr;
    uint32_t* d_smi = e.h_smi ? reinterpret_cast<uint32_t*>(d + o_smi) : nullptr;
    float* d_ml = e.h_ml ? reinterpret_cast<float*>(d + o_ml) : nullptr;
    uint64_t* d_ms = e.h_ms ? reinterpret_cast<uint64_t*>(d + o_ms) : nullptr;
    e.d_b = reinterpret_cast<uint32_t*>(d + o_b);
    e.d_m = reinterpret_cast<uint32_t*>(d + o_m);
    e.d_t = reinterpret_cast<uint32_t*>(d + o_t);
    e.d_q = reinterpret_cast<float*>(d + o_q);
    e.d_p = reinterpret_cast<b200seed_bound_params*>(d + o_p);
    uint32_t* d_n = reinterpret_cast<uint32_t*>(d + o_n);
    b200seed_counters* d_c = reinterpret_cast<b200seed_counters*>(d + o_c);

    CUDA_TRY(h, cudaMemcpyAsync(d_xyz, e.h_xyz, size_t(n_sp) * 12, cudaMemcpyHostToDevice, s));
    if (d_vz) CUDA_TRY(h, cudaMemcpyAsync(d_vz, e.h_var_z, size_t(n_sp) * 4, cudaMemcpyHostToDevice, s));
    if (d_vr) CUDA_TRY(h, cudaMemcpyAsync(d_vr, e.h_var_r, size_t(n_sp) * 4, cudaMemcpyHostToDevice, s));
    if (want_params && !e.compact) {
        if (d_smi)
            CUDA_TRY(h, cudaMemcpyAsync(d_smi, e.h_smi, size_t(n_sp) * 4, cudaMemcpyHostToDevice, s));
        if (d_ml)
            CUDA_TRY(h, cudaMemcpyAsync(d_ml, e.h_ml, size_t(n_meas) * 8, cudaMemcpyHostToDevice, s));
        if (d_ms)
            CUDA_TRY(h, cudaMemcpyAsync(d_ms, e.h_ms, size_t(n_meas) * 8, cudaMemcpyHostToDevice, s));
    }
    // The 56-byte counters record (it carries n_seeds, which sizes the copies below) is written by
    // k_seed_gather straight into the handle's pinned, device-mapped host word: a D->H copy of it
    // would queue behind other events' 11-MB parameter copies on the copy engine.
    static const bool mapped_counters = std::getenv("B200SEED_COPY_COUNTERS") == nullptr;
    if (mapped_counters) d_c = h->h_pinned;
    int rc = b200seed_run(h, s, n_sp, d_xyz, d_vz, d_vr, d + o_ws, ws_bytes, seed_capacity, e.d_b,
                          e.d_m, e.d_t, e.d_q, d_n, d_c);
    if (rc != B200SEED_OK) return rc;
    if (e.compact) {
        rc = b200seed_estimate_params_compact(h, s, d_n, seed_capacity, e.d_b, e.d_m, e.d_t, d_xyz, e.bfield,
                                              reinterpret_cast<b200seed_seed_params*>(e.d_p));
        if (rc != B200SEED_OK) return rc;
    } else if (e.packed) {
        rc = b200seed_estimate_params_packed(h, s, d_n, seed_capacity, e.d_b, e.d_m, e.d_t, d_xyz, d_smi, d_ml,
                                             d_ms, e.bfield,
                                             reinterpret_cast<b200seed_bound_params_packed*>(e.d_p));
        if (rc != B200SEED_OK) return rc;
    } else if (want_params) {
        rc = diag ? b200seed_estimate_params_diag(h, s, d_n, seed_capacity, e.d_b, e.d_m, e.d_t, d_xyz, d_smi,
                                                  d_ml, d_ms, e.bfield,
                                                  reinterpret_cast<b200seed_bound_params_diag*>(e.d_p))
                  : b200seed_estimate_params(h, s, d_n, seed_capacity, e.d_b, e.d_m, e.d_t, d_xyz, d_smi,
                                             d_ml, d_ms, e.bfield, e.d_p);
        if (rc != B200SEED_OK) return rc;
    }
    // the counters struct carries n_seeds: one small read-back, then the sized copies
    if (!mapped_counters)
        CUDA_TRY(h, cudaMemcpyAsync(h->h_pinned, d_c, sizeof(b200seed_counters), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(h, host_wait_point(h, s));
    e.submitted = true;
    return B200SEED_OK;
}

int host_finish(b200seed_handle* h, cudaStream_t s, HostEvent& e, uint32_t* h_n_seeds,
                b200seed_counters* h_counters) {
    *h_n_seeds = 0;
    if (h_counters) std::memset(h_counters, 0, sizeof(*h_counters));
    if (!e.submitted) return B200SEED_OK;
    e.submitted = false;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaEventSynchronize(h->ev_host));
    const uint32_t n = h->h_pinned->n_seeds;
    *h_n_seeds = n;
    if (h_counters) *h_counters = *h->h_pinned;
    if (n) {
        if (e.h_bottom) CUDA_TRY(h, cudaMemcpyAsync(e.h_bottom, e.d_b, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
        if (e.h_middle) CUDA_TRY(h, cudaMemcpyAsync(e.h_middle, e.d_m, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
        if (e.h_top) CUDA_TRY(h, cudaMemcpyAsync(e.h_top, e.d_t, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
        if (e.h_quality) CUDA_TRY(h, cudaMemcpyAsync(e.h_quality, e.d_q, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
        b200seed_seed_params* h_sp = static_cast<b200seed_seed_params*>(h->h_compact);
        const uint32_t* h_bot = e.h_bottom;
        if (e.compact) {
            CUDA_TRY(h, cudaMemcpyAsync(h_sp, e.d_p, size_t(n) * sizeof(b200seed_seed_params),
                                        cudaMemcpyDeviceToHost, s));
            if (!h_bot) {  // the expansion needs the bottom spacepoint of every seed
                uint32_t* stage = reinterpret_cast<uint32_t*>(h_sp + e.seed_capacity);
                CUDA_TRY(h, cudaMemcpyAsync(stage, e.d_b, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
                h_bot = stage;
            }
        } else if (e.packed) {
            // straight into the caller's buffer if that form was asked for
            CUDA_TRY(h, cudaMemcpyAsync(e.h_params_packed ? static_cast<void*>(e.h_params_packed) : h->h_compact,
                                        e.d_p, size_t(n) * sizeof(b200seed_bound_params_packed),
                                        cudaMemcpyDeviceToHost, s));
        } else if (e.h_params_diag) {
            CUDA_TRY(h, cudaMemcpyAsync(e.h_params_diag, e.d_p, size_t(n) * sizeof(b200seed_bound_params_diag),
                                        cudaMemcpyDeviceToHost, s));
        } else if (e.h_params) {
            CUDA_TRY(h, cudaMemcpyAsync(e.h_params, e.d_p, size_t(n) * sizeof(b200seed_bound_params),
                                        cudaMemcpyDeviceToHost, s));
        }
        CUDA_TRY(h, host_wait_point(h, s));
        CUDA_TRY(h, cudaEventSynchronize(h->ev_host));
        e.h_bot_stage = h_bot;
    }
    // The device buffers of this handle are free again from here on; what is left (host_expand) is
    // host work on the pinned landing area, which the next host_submit on this handle leaves alone.
    if (const uint32_t ovf = h->h_pinned->overflow) {
        *h->h_sticky = 0u;  // reported here
        return fail(h, B200SEED_EOVERFLOW, overflow_message(ovf));
    }
    return B200SEED_OK;
}

// Second half of a host-buffer event: the parameter records, completed on the host.
void host_expand(b200seed_handle* h, const HostEvent& e, uint32_t n) {
    if (n == 0) return;
    if (e.compact)
        b200seed_expand_seed_params(h, n, e.h_bot_stage, static_cast<const b200seed_seed_params*>(h->h_compact),
                                    e.h_smi, e.h_ml, e.h_ms, e.h_params, e.h_params_diag);
    else if (e.packed) {
        if (e.h_params || e.h_params_diag)
            b200seed_expand_packed_params(
                h, n,
                e.h_params_packed ? e.h_params_packed : static_cast<const b200seed_bound_params_packed*>(h->h_compact),
                e.h_params, e.h_params_diag);
    }
    // both forms requested: the full records are expanded on the host
    else if (e.h_params_diag && e.h_params)
        b200seed_expand_params(e.h_params_diag, n, e.h_params);
}

// host_submit would have to enlarge the pinned landing area of this handle (which may still hold an
// event waiting for host_expand)
bool host_submit_regrows_landing(const b200seed_handle* h, const HostEvent& e) {
    return size_t(e.seed_capacity) * sizeof(b200seed_bound_params_packed) > h->h_compact_bytes;
}

}  // namespace

extern "C" int b200seed_check_overflow(b200seed_handle* h, uint32_t* mask_out) {
    if (!h) return B200SEED_EINVAL;
    const uint32_t ovf = h->h_sticky ? *static_cast<volatile uint32_t*>(h->h_sticky) : 0u;
    if (mask_out) *mask_out = ovf;
    if (ovf == 0u) return B200SEED_OK;
    *h->h_sticky = 0u;
    return fail(h, B200SEED_EOVERFLOW, overflow_message(ovf));
}

namespace {

HostEvent host_event_of(const b200seed_event_io& io) {
    HostEvent e;
    e.n_sp = io.n_spacepoints, e.n_meas = io.n_measurements, e.seed_capacity = io.seed_capacity;
    e.h_xyz = io.xyz, e.h_var_z = io.var_z, e.h_var_r = io.var_r, e.h_smi = io.sp_meas_index_1;
    e.h_ml = io.meas_local, e.h_ms = io.meas_surface;
    e.bfield[0] = io.bfield[0], e.bfield[1] = io.bfield[1], e.bfield[2] = io.bfield[2];
    e.h_bottom = io.bottom, e.h_middle = io.middle, e.h_top = io.top, e.h_quality = io.quality;
    e.h_params = io.params;
    e.h_params_diag = io.params_diag;
    e.h_params_packed = io.params_packed;
    return e;
}

}  // namespace

// ---------------------------------------------------------------------------
// Event pool: the host side of a throughput job on one device — worker threads, each with
// two algorithm instances + streams, so that a worker always has one event computing while
// it collects the previous one (the reference's throughput_mt app runs one
// full_chain_algorithm per TBB thread: examples/run/common/.../throughput_mt.ipp:170-298).
// ---------------------------------------------------------------------------
struct b200seed_pool {
    struct Slot {
        b200seed_handle* h = nullptr;
        cudaStream_t s = nullptr;
        HostEvent ev;
        long pending = -1;
    };
    struct Worker {
        Slot slot[2];
        std::thread th;
    };
    int device = 0;
    std::vector<Worker> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    b200seed_event_io* events = nullptr;
    uint32_t n_events = 0;
    std::atomic<uint32_t> next{0};
    uint64_t generation = 0;
    int running = 0;
    bool stop = false;
    std::string error;

    void work(Worker& w) {
        uint64_t seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
            }
            cudaSetDevice(device);
            int k = 0;
            // an event whose copies are done and whose records still have to be completed on the
            // host: that happens after the next event went to the device, not before
            Slot* todo = nullptr;
            HostEvent todo_ev;
            uint32_t todo_n = 0;
            auto expand_todo = [&] {
                if (todo) host_expand(todo->h, todo_ev, todo_n);
                todo = nullptr;
            };
            while (true) {
                Slot& cur = w.slot[k];
                const uint32_t idx = next.fetch_add(1);
                if (idx < n_events) {
                    const HostEvent ne = host_event_of(events[idx]);
                    if (todo == &cur && host_submit_regrows_landing(cur.h, ne)) expand_todo();
                    cur.ev = ne;
                    const int rc = host_submit(cur.h, cur.s, cur.ev);
                    events[idx].status = rc;
                    cur.pending = idx;
                }
                expand_todo();
                Slot& oth = w.slot[k ^ 1];
                if (oth.pending >= 0) {
                    b200seed_event_io& io = events[oth.pending];
                    if (io.status == B200SEED_OK) {
                        io.status = host_finish(oth.h, oth.s, oth.ev, &io.n_seeds, &io.counters);
                        if (io.status == B200SEED_OK || io.status == B200SEED_EOVERFLOW) {
                            todo = &oth;
                            todo_ev = oth.ev;
                            todo_n = io.n_seeds;
                        }
                    }
                    oth.pending = -1;
                }
                if (idx >= n_events && cur.pending < 0) break;
                k ^= 1;
            }
            expand_todo();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--running == 0) cv_done.notify_all();
            }
        }
    }
};

extern "C" {

int b200seed_pool_create(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                         const b200seed_filter_cfg* filter, const b200seed_tpe_cfg* tpe, int device,
                         int n_workers, b200seed_pool** out) {
    if (!out || n_workers < 1 || n_workers > 64)
        return fail(nullptr, B200SEED_EINVAL, "b200seed_pool_create: bad argument");
    *out = nullptr;
    b200seed_pool* p = new b200seed_pool();
    p->device = device;
    p->workers = std::vector<b200seed_pool::Worker>(size_t(n_workers));
    for (auto& w : p->workers) {
        for (auto& sl : w.slot) {
            int rc = b200seed_create(finder, grid, filter, tpe, device, &sl.h);
            if (rc == B200SEED_OK && cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking) != cudaSuccess)
                rc = fail(nullptr, B200SEED_ECUDA, "cudaStreamCreate failed");
            if (rc != B200SEED_OK) {
                b200seed_pool_destroy(p);
                return rc;
            }
        }
    }
    for (auto& w : p->workers) w.th = std::thread([p, &w] { p->work(w); });
    *out = p;
    return B200SEED_OK;
}

int b200seed_pool_process(b200seed_pool* p, b200seed_event_io* events, uint32_t n_events) {
    if (!p || (!events && n_events)) return B200SEED_EINVAL;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->events = events;
        p->n_events = n_events;
        p->next.store(0);
        p->running = int(p->workers.size());
        ++p->generation;
    }
    p->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lk(p->mu);
        p->cv_done.wait(lk, [&] { return p->running == 0; });
    }
    for (uint32_t i = 0; i < n_events; ++i)
        if (events[i].status != B200SEED_OK) {
            p->error = "event " + std::to_string(i) + " failed with status " + std::to_string(events[i].status);
            return events[i].status;
        }
    return B200SEED_OK;
}

const char* b200seed_pool_last_error(const b200seed_pool* p) {
    return p ? p->error.c_str() : g_create_error.c_str();
}

void b200seed_pool_destroy(b200seed_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_job.notify_all();
    for (auto& w : p->workers) {
        if (w.th.joinable()) w.th.join();
        for (auto& sl : w.slot) {
            if (sl.s) cudaStreamDestroy(sl.s);
            if (sl.h) b200seed_destroy(sl.h);
        }
    }
    delete p;
}

int b200seed_run_host(b200seed_handle* h, void* stream, uint32_t n_sp, const float* h_xyz,
                      const float* h_var_z, const float* h_var_r,
                      const uint32_t* h_sp_meas_index_1, uint32_t n_meas,
                      const float* h_meas_local, const uint64_t* h_meas_surface,
                      const float bfield[3], uint32_t seed_capacity, uint32_t* h_bottom,
                      uint32_t* h_middle, uint32_t* h_top, float* h_quality,
                      b200seed_bound_params* h_params, uint32_t* h_n_seeds,
                      b200seed_counters* h_counters) {
    if (!h) return B200SEED_EINVAL;
    if (!h_n_seeds) return fail(h, B200SEED_EINVAL, "b200seed_run_host: h_n_seeds is null");
    if (h_params && !bfield) return fail(h, B200SEED_EINVAL, "b200seed_run_host: bfield is null");
    HostEvent e;
    e.n_sp = n_sp, e.n_meas = n_meas, e.seed_capacity = seed_capacity;
    e.h_xyz = h_xyz, e.h_var_z = h_var_z, e.h_var_r = h_var_r, e.h_smi = h_sp_meas_index_1;
    e.h_ml = h_meas_local, e.h_ms = h_meas_surface;
    if (bfield) e.bfield[0] = bfield[0], e.bfield[1] = bfield[1], e.bfield[2] = bfield[2];
    e.h_bottom = h_bottom, e.h_middle = h_middle, e.h_top = h_top, e.h_quality = h_quality;
    e.h_params = h_params;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    *h_n_seeds = 0;
    if (h_counters) std::memset(h_counters, 0, sizeof(*h_counters));
    int rc = host_submit(h, s, e);
    if (rc != B200SEED_OK) return rc;
    rc = host_finish(h, s, e, h_n_seeds, h_counters);
    if (rc == B200SEED_OK || rc == B200SEED_EOVERFLOW) host_expand(h, e, *h_n_seeds);
    return rc;
}

// Measured non-fused FP32 rate of the device (ops/s) — the denominator bench.py uses for the
// FP32-issue roofline of the doublet / triplet kernels.
int b200seed_measure_fp32_peak(int device, double* ops_per_s) {
    if (!ops_per_s) return B200SEED_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, B200SEED_ECUDA, "cudaSetDevice");
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float* d = nullptr;
    if (cudaMalloc(&d, 256) != cudaSuccess) return fail(nullptr, B200SEED_ECUDA, "cudaMalloc");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000, blocks = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_fp32_probe<<<blocks, 256>>>(d, iters, 1.0000001f, 1e-7f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = double(blocks) * 256.0 * iters * 16.0;
        if (ms > 0.f && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, B200SEED_ECUDA, cudaGetErrorString(e));
    *ops_per_s = best;
    return B200SEED_OK;
}

// ---------------------------------------------------------------------------
// Host probes: the device cut arithmetic (seed_math.cuh) compiled for the host, so the
// CPU test-suite can compare it with the oracle bit for bit without a GPU. Not used by
// any product path.
// ---------------------------------------------------------------------------
int b200seed_host_probe_devcfg(const b200seed_finder_cfg* f, const b200seed_grid_cfg* g,
                               const b200seed_filter_cfg* fl, void* out, size_t out_bytes) {
    if (!f || !g || !fl || !out || out_bytes < sizeof(DevCfg)) return B200SEED_EINVAL;
    DevCfg d{};
    std::string why;
    if (compute_axes(*g, d, why) != 0) return fail(nullptr, B200SEED_EINVAL, why);
    fill_devcfg(*f, *fl, d);
    std::memcpy(out, &d, sizeof(d));
    return int(sizeof(DevCfg));
}
float b200seed_host_probe_atan2f(float y, float x) {
    return fd_atan2f(y, x);
}
// bins of n spacepoints (0xFFFFFFFF = invalid)
void b200seed_host_probe_bins(const void* devcfg, uint32_t n, const float* xyz, uint32_t* bins) {
    const DevCfg& d = *static_cast<const DevCfg*>(devcfg);
    for (uint32_t i = 0; i < n; ++i) bins[i] = sp_bin(d, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
// doublet decision for n (middle, other) pairs: 0 none, 1 bottom, 2 top; lin_circle of
// accepted pairs (Zo,cotTheta,iDeltaR,Er,U,V). m/o = {x,y,z,varZ,varR} per pair.
void b200seed_host_probe_doublets(const void* devcfg, uint32_t n, const float* m, const float* o,
                                  int32_t* kind, float* lc) {
    const DevCfg& d = *static_cast<const DevCfg*>(devcfg);
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = m + 5 * size_t(i);
        const float* b = o + 5 * size_t(i);
        const float rM = sp_radius(a[0], a[1]), r2 = sp_radius(b[0], b[1]);
        int st = doublet_stage1(d, rM, a[2], r2, b[2]);
        if (st && !doublet_stage2(d, a[0], a[1], b[0], b[1])) st = 0;
        kind[i] = st;
        if (st) {
            const LinCircle l = transform_coordinates(st == 1, a[0], a[1], a[2], rM, a[3], a[4], b[0],
                                                      b[1], b[2], b[3], b[4]);
            float* q = lc + 6 * size_t(i);
            q[0] = l.Zo, q[1] = l.cotTheta, q[2] = l.iDeltaR, q[3] = l.Er, q[4] = l.U, q[5] = l.V;
        }
    }
}
// Helix-radius cut of n pairs {x1,y1,x2,y2}: exact[i] = doublet_stage2, fast[i] = the
// division-free pre-decision (0 fail, 1 pass, 2 undecided).
void b200seed_host_probe_stage2(const void* devcfg, uint32_t n, const float* xy, int32_t* exact,
                                int32_t* fast) {
    const DevCfg& d = *static_cast<const DevCfg*>(devcfg);
    for (uint32_t i = 0; i < n; ++i) {
        const float* q = xy + 4 * size_t(i);
        exact[i] = doublet_stage2(d, q[0], q[1], q[2], q[3]) ? 1 : 0;
        fast[i] = doublet_stage2_fast(d, q[0], q[1], q[2], q[3]);
    }
}
void b200seed_host_probe_stage2_bounded(const void* devcfg, uint32_t n, const float* xy, int32_t* fast) {
    const DevCfg& d = *static_cast<const DevCfg*>(devcfg);
    for (uint32_t i = 0; i < n; ++i) {
        const float* q = xy + 4 * size_t(i);
        fast[i] = doublet_stage2_fast_bounded(d, q[0], q[1], q[2], q[3]);
    }
}
// Pruning index: for n (middle, other) pairs, whether the other spacepoint's cell lies inside
// the cell window k_doublets visits for that middle (grid sized for n_sp spacepoints).
// m/o = {x,y,z,varZ,varR}; bins = reference bin of each `other` (from ..._probe_bins).
// grid_out (optional) receives {NR, NZc}.
void b200seed_host_probe_cell_window(const void* devcfg, const b200seed_finder_cfg* finder,
                                     uint32_t n_sp, uint32_t n, const float* m, const float* o,
                                     const uint32_t* o_bin, int32_t* visited, uint32_t* grid_out) {
    const DevCfg& d = *static_cast<const DevCfg*>(devcfg);
    const CellGrid g = make_cell_grid(d, *finder, n_sp);
    if (grid_out) {
        grid_out[0] = g.NR;
        grid_out[1] = g.NZc;
    }
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = m + 5 * size_t(i);
        const float* b = o + 5 * size_t(i);
        const float rM = sp_radius(a[0], a[1]), r2 = sp_radius(b[0], b[1]);
        const float er = 1e-2f + 1e-5f * (rM + absf(d.deltaRMax));
        const uint32_t row_lo = cell_row(g, rM - d.deltaRMax - er);
        const uint32_t row_hi = cell_row(g, rM + d.deltaRMax + er);
        const uint32_t row = cell_row(g, r2);
        const uint32_t zb = o_bin[i] / d.nPhi;
        int v = 0;
        float L, U;
        if (row >= row_lo && row <= row_hi && cell_row_window(d, g, rM, a[2], row, L, U)) {
            const uint32_t c = cell_z(g, zb, b[2]);
            v = (c >= cell_z(g, zb, L) && c <= cell_z(g, zb, U)) ? 1 : 0;
        }
        visited[i] = v;
    }
}
// triplet decision for n (middle, lb, lt) combinations; out = {curvature, impact}
void b200seed_host_probe_triplets(const void* devcfg, uint32_t n, const float* m, const float* lb,
                                  const float* lt, int32_t* ok, int32_t* cut1, float* out) {
    const DevCfg& d = *static_cast<const DevCfg*>(devcfg);
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = m + 5 * size_t(i);
        const float* pb = lb + 6 * size_t(i);
        const float* pt = lt + 6 * size_t(i);
        const LinCircle b{pb[0], pb[1], pb[2], pb[3], pb[4], pb[5]};
        const LinCircle t{pt[0], pt[1], pt[2], pt[3], pt[4], pt[5]};
        float is2, s2;
        triplet_row_constants(d, b.cotTheta, is2, s2);
        cut1[i] = triplet_cut1(b.cotTheta, b.iDeltaR, b.Er, t.cotTheta, t.iDeltaR, t.Er, a[4], a[3], s2)
                      ? 1
                      : 0;
        float c = 0.f, ip = 0.f;
        ok[i] = triplet_is_compatible(d, sp_radius(a[0], a[1]), a[4], a[3], b, t, is2, s2, c, ip) ? 1 : 0;
        out[2 * size_t(i)] = c;
        out[2 * size_t(i) + 1] = ip;
    }
}

// division-free pre-filter of the triplet cuts for n (middle, lb, lt) combinations:
// rej[i] = 1 if triplet_certainly_rejected (then the exact cuts must reject as well)
void b200seed_host_probe_triplet_prefilter(const void* devcfg, uint32_t n, const float* m,
                                           const float* lb, const float* lt, int32_t* rej) {
    const DevCfg& d = *static_cast<const DevCfg*>(devcfg);
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = m + 5 * size_t(i);
        const float* pb = lb + 6 * size_t(i);
        const float* pt = lt + 6 * size_t(i);
        rej[i] = triplet_certainly_rejected(d, sp_radius(a[0], a[1]), pb[4], pb[5], pt[4], pt[5]) ? 1 : 0;
    }
}

}  // extern "C"


#ifdef B200_TAIL_PROBE
// debug build only (tools/tail_probe.py, tools/middle_cost.py)
extern "C" int b200seed_debug_tail_probe(unsigned long long* out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, b200seed::g_tail_probe, sizeof(b200seed::g_tail_probe));
    if (reset) {
        void* p = nullptr;
        cudaGetSymbolAddress(&p, b200seed::g_tail_probe);
        cudaMemset(p, 0, sizeof(b200seed::g_tail_probe));
    }
    return 0;
}
extern "C" int b200seed_debug_middle_cycles(uint32_t* out, uint32_t n) {
    return int(cudaMemcpyFromSymbol(out, b200seed::g_middle_cycles, size_t(n) * 4));
}
#endif
