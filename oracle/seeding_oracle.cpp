// seeding_oracle.cpp — CPU ORACLE for the traccc triplet-seeding hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it. The product
// (traccc_b200/, libb200seed.so) never links, imports or falls back to it.
//
// It restates, function by function, the reference's *host* algorithms
//   traccc::host::seeding_algorithm        core/src/seeding/seeding_algorithm.cpp:24-28
//   traccc::host::track_params_estimation  core/src/seeding/track_params_estimation.cpp:23-91
// in dependency-free C++ (the reference itself needs vecmem/detray/Acts, which are
// fetched from the network at configure time and are absent here).
// Each function cites the reference file:line it follows. Build flags mimic the
// reference CPU build: -O2 -march=x86-64-v2 (no FMA) -ffp-contract=off
// (cmake/traccc-compiler-options-cpp.cmake:56-64).
//
// Parity pinning: the oracle is checked against the reference's own known answers
// (tests/cpu/test_seeding.cpp, test_track_params_estimation.cpp, test_axis.cpp) in
// tests/test_oracle_kat.py, and its cut arithmetic against the reference's helper
// headers compiled verbatim from /root/reference (oracle/_ref, see oracle/Makefile).
// Third-party arithmetic (detray/algebra-plugins vector ops inside
// seed_to_bound_param_vector) is restated from its published definitions: that part
// is "parity unpinned" below the reference's own 2e-4 |p| tolerance.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

#include "../include/b200seed.h"  // POD config structs only (type definitions)

namespace {

// ---------------------------------------------------------------------------
// Units — detray::unit<float> (detray/definitions/units.hpp, external, restated):
// mm = 1, GeV = 1, MeV = 1e-3, e = 1, T = 0.000299792458 GeV/(e mm), degree = pi/180,
// s = 299792458000 mm (c = 1), ns = 1e-9 s.
// ---------------------------------------------------------------------------
constexpr float unit_mm = 1.f;
constexpr float unit_GeV = 1.f;
constexpr float unit_MeV = 1e-3f;
constexpr float unit_T = static_cast<float>(0.000299792458);
constexpr float unit_degree = static_cast<float>(0.017453292519943295);
constexpr float unit_ns = static_cast<float>(1e-9 * 299792458000.0);

// ---------------------------------------------------------------------------
// atan2f — the classic Sun fdlibm float algorithm (e_atan2f.c + s_atanf.c), which is
// what glibc 2.39's atan2f computes on x86-64 (no FMA/multiarch variant). Restated so
// the oracle does not depend on the libm of the box it runs on (glibc >= 2.41 ships a
// correctly rounded CORE-MATH atan2f). oracle_selftest_atan2f() reports whether the
// local libm agrees. Used for spacepoint phi: edm/impl/spacepoint_collection.ipp:59-63.
// ---------------------------------------------------------------------------
inline uint32_t f2u(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f,
                         1.5707962513e+00f};
const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f,
                         7.5497894159e-08f};
const float aT[11] = {3.3333334327e-01f,  -2.0000000298e-01f, 1.4285714924e-01f,
                      -1.1111110449e-01f, 9.0908870101e-02f,  -7.6918758452e-02f,
                      6.6610731184e-02f,  -5.8335702866e-02f, 4.9768779427e-02f,
                      -3.6531571299e-02f, 1.6285819933e-02f};

float fd_atanf(float x) {
    float w, s1, s2, z;
    int32_t ix, hx, id;
    hx = static_cast<int32_t>(f2u(x));
    ix = hx & 0x7fffffff;
    if (ix >= 0x4c800000) { /* |x| >= 2^26 */
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return atanhi[3] + atanlo[3];
        return -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {     /* |x| < 0.4375 */
        if (ix < 0x31000000) { /* |x| < 2^-29 */
            return x;
        }
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3f980000) {     /* |x| < 1.1875 */
            if (ix < 0x3f300000) { /* 7/16 <= |x| < 11/16 */
                id = 0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else { /* 11/16 <= |x| < 19/16 */
                id = 1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000) { /* |x| < 2.4375 */
                id = 2;
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else { /* 2.4375 <= |x| < 2^26 */
                id = 3;
                x = -1.0f / x;
            }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return (hx < 0) ? -z : z;
}

float fd_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = 7.8539818525e-01f;
    const float pi_o_2 = 1.5707963705e+00f;
    const float pi = 3.1415927410e+00f;
    const float pi_lo = -8.7422776573e-08f;
    float z;
    int32_t k, m, hx, hy, ix, iy;
    hx = static_cast<int32_t>(f2u(x));
    ix = hx & 0x7fffffff;
    hy = static_cast<int32_t>(f2u(y));
    iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return fd_atanf(y);
    m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
            case 0:
            case 1:
                return y;
            case 2:
                return pi + tiny;
            default:
                return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
                case 0:
                    return pi_o_4 + tiny;
                case 1:
                    return -pi_o_4 - tiny;
                case 2:
                    return 3.0f * pi_o_4 + tiny;
                default:
                    return -3.0f * pi_o_4 - tiny;
            }
        } else {
            switch (m) {
                case 0:
                    return 0.0f;
                case 1:
                    return -0.0f;
                case 2:
                    return pi + tiny;
                default:
                    return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    k = (iy - ix) >> 23;
    if (k > 60)
        z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60)
        z = 0.0f;
    else
        z = fd_atanf(std::fabs(y / x));
    switch (m) {
        case 0:
            return z;
        case 1:
            return u2f(f2u(z) ^ 0x80000000u);
        case 2:
            return pi - (z - pi_lo);
        default:
            return (z - pi_lo) - pi;
    }
}

bool g_use_libm_atan2 = false;
inline float oracle_atan2f(float y, float x) {
    return g_use_libm_atan2 ? std::atan2(y, x) : fd_atan2f(y, x);
}

// ---------------------------------------------------------------------------
// Axes — core/include/traccc/seeding/grids/axis.hpp
// ---------------------------------------------------------------------------
struct Axis {
    uint32_t n_bins;
    float min, max;
};

// axis2::regular::bin (axis.hpp:87-99) / axis2::circular::bin (axis.hpp:277-289)
inline int axis_ibin(const Axis& a, float v) {
    return static_cast<int>((v - a.min) / (a.max - a.min) * static_cast<float>(a.n_bins));
}
uint32_t regular_bin(const Axis& a, float v) {
    int ibin = axis_ibin(a, v);
    if (ibin >= 0 && ibin < static_cast<int>(a.n_bins)) return static_cast<uint32_t>(ibin);
    if (ibin < 0) return 0;
    return a.n_bins - 1u;
}
uint32_t circular_bin(const Axis& a, float v) {
    int ibin = axis_ibin(a, v);
    if (ibin >= 0 && ibin < static_cast<int>(a.n_bins)) return static_cast<uint32_t>(ibin);
    if (ibin < 0) return a.n_bins + static_cast<uint32_t>(ibin);
    return static_cast<uint32_t>(ibin) - a.n_bins;
}
// axis2::regular::range, binned neighbourhood (axis.hpp:109-123)
void regular_range(const Axis& a, float v, const uint32_t nhood[2], uint32_t out[2]) {
    int ibin = axis_ibin(a, v);
    int ibinmin = ibin - static_cast<int>(nhood[0]);
    int ibinmax = ibin + static_cast<int>(nhood[1]);
    out[0] = (ibinmin >= 0) ? static_cast<uint32_t>(ibinmin) : 0u;
    out[1] = (ibinmax < static_cast<int>(a.n_bins)) ? static_cast<uint32_t>(ibinmax)
                                                    : a.n_bins - 1u;
}
// axis2::circular::remap (axis.hpp:394-404)
uint32_t circular_remap(const Axis& a, uint32_t ibin, int shood) {
    int opt_bin = static_cast<int>(ibin) + shood;
    if (opt_bin >= 0 && opt_bin < static_cast<int>(a.n_bins))
        return static_cast<uint32_t>(opt_bin);
    if (opt_bin < 0) return static_cast<uint32_t>(static_cast<int>(a.n_bins) + opt_bin);
    return static_cast<uint32_t>(opt_bin) - a.n_bins;
}
// axis2::circular::range, binned neighbourhood (axis.hpp:299-306)
void circular_range(const Axis& a, float v, const uint32_t nhood[2], uint32_t out[2]) {
    uint32_t gbin = circular_bin(a, v);
    out[0] = circular_remap(a, gbin, -static_cast<int>(nhood[0]));
    out[1] = circular_remap(a, gbin, static_cast<int>(nhood[1]));
}
// axis2::regular::zone_t (axis.hpp:162-173)
std::vector<uint32_t> regular_zone(const Axis& a, float v, const uint32_t nhood[2]) {
    uint32_t r[2];
    regular_range(a, v, nhood, r);
    std::vector<uint32_t> seq(static_cast<size_t>(r[1] - r[0] + 1u), r[0]);
    uint32_t m = 0;
    for (auto& n : seq) n += m++;
    return seq;
}
// axis2::circular::zone_t (axis.hpp:336-362)
std::vector<uint32_t> circular_zone(const Axis& a, float v, const uint32_t nhood[2]) {
    uint32_t r[2];
    circular_range(a, v, nhood, r);
    if (r[0] < r[1]) {
        std::vector<uint32_t> seq(static_cast<size_t>(r[1] - r[0] + 1u), r[0]);
        uint32_t m = 0;
        for (auto& n : seq) n += m++;
        return seq;
    }
    uint32_t vl = a.n_bins - r[0] + r[1] + 1u;
    uint32_t mi = 0, mo = 0;
    std::vector<uint32_t> seq(static_cast<size_t>(vl), r[0]);
    for (auto& n : seq) {
        n += mi++;
        if (n > a.n_bins - 1u) n = mo++;
    }
    return seq;
}

// get_axes — core/include/traccc/seeding/spacepoint_binning_helper.hpp:22-110.
// Returns 0, or -1 where the reference throws std::domain_error (:33-38, :77-81).
int get_axes(const b200seed_grid_cfg& g, Axis& phi_axis, Axis& z_axis) {
    uint32_t phiBins;
    if (g.bFieldInZ == 0) {
        phiBins = 100;
    } else {
        float minHelixRadius = g.minPt / g.bFieldInZ;
        if (minHelixRadius < g.rMax / 2) return -1;
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
        if (deltaPhi <= 0.) return -1;
        phiBins = static_cast<uint32_t>(std::llround(2 * M_PI / deltaPhi + 0.5));
    }
    phi_axis = Axis{phiBins, g.phiMin, g.phiMax};
    float zBinSize = g.cotThetaMax * g.deltaRMax;
    uint32_t zBins = std::max(
        static_cast<uint32_t>(1),
        static_cast<uint32_t>(std::floor((g.zMax - g.zMin) / zBinSize)));
    z_axis = Axis{zBins, g.zMin, g.zMax};
    return 0;
}

// ---------------------------------------------------------------------------
// Spacepoint accessors — core/include/traccc/edm/impl/spacepoint_collection.ipp:51-63
// ---------------------------------------------------------------------------
struct Sp {
    float x, y, z, varZ, varR;
    float radius() const {
        const float xx = x, yy = y;
        return std::sqrt(xx * xx + yy * yy);
    }
    float phi() const { return oracle_atan2f(y, x); }
};

// is_valid_sp — spacepoint_binning_helper.hpp:112-127. vector::perp (algebra-plugins,
// external) restated as sqrt(x^2 + y^2); get_num_rbins = size_t(rMax + |beamPos|)
// (seeding_config.hpp:110-113).
bool is_valid_sp(const b200seed_finder_cfg& c, const Sp& sp) {
    if (sp.z > c.zMax || sp.z < c.zMin) return false;
    float spPhi = oracle_atan2f(sp.y, sp.x);
    if (spPhi > c.phiMax || spPhi < c.phiMin) return false;
    const float px = sp.x - c.beamPos[0], py = sp.y - c.beamPos[1];
    const float perp = std::sqrt(px * px + py * py);
    const float bnorm = std::sqrt(c.beamPos[0] * c.beamPos[0] + c.beamPos[1] * c.beamPos[1]);
    return static_cast<size_t>(perp) < static_cast<size_t>(c.rMax + bnorm);
}

// lin_circle — core/include/traccc/seeding/detail/lin_circle.hpp
struct LinCircle {
    float Zo, cotTheta, iDeltaR, Er, U, V;
};

// doublet_finding_helper::isCompatible — doublet_finding_helper.hpp:51-216.
// `bottom` selects details::spacepoint_type::bottom. *stage1 reports whether the pair
// survived the first block of cuts (:77-84), for the work counters only.
bool doublet_is_compatible(bool bottom, const Sp& sp1, const Sp& sp2,
                           const b200seed_finder_cfg& config, bool* stage1) {
    float deltaR, cotTheta, zOrigin;
    if (bottom) {
        deltaR = sp1.radius() - sp2.radius();
        cotTheta = sp1.z - sp2.z;
        zOrigin = sp1.z * deltaR - sp1.radius() * cotTheta;
    } else {
        deltaR = sp2.radius() - sp1.radius();
        cotTheta = (sp2.z - sp1.z);
        zOrigin = sp1.z * deltaR - sp1.radius() * cotTheta;
    }
    if ((deltaR >= config.deltaRMax) || (deltaR <= config.deltaRMin) ||
        (std::fabs(cotTheta) >= config.cotThetaMax * deltaR) ||
        (zOrigin <= config.collisionRegionMin * deltaR) ||
        (zOrigin >= config.collisionRegionMax * deltaR) ||
        std::fabs(cotTheta) >= config.deltaZMax) {
        *stage1 = false;
        return false;
    }
    *stage1 = true;

    float midX = 0.5f * (sp1.x + sp2.x);
    float midY = 0.5f * (sp1.y + sp2.y);
    float slope = (sp2.y - sp1.y) / (sp2.x - sp1.x);
    float deltaX = sp2.x - sp1.x;
    float deltaY = sp2.y - sp1.y;
    float deltaXY2 = deltaX * deltaX + deltaY * deltaY;
    float sagittaLength =
        std::sqrt(config.minHelixRadius * config.minHelixRadius - deltaXY2 / 4.f);
    float denom = std::sqrt((slope * slope + 1) / (slope * slope));
    float cosCentralAngle = 1.f / denom;
    float sinCentralAngle = -1.f / (slope * denom);
    float mpDeltaX = sagittaLength * cosCentralAngle;
    float mpDeltaY = sagittaLength * sinCentralAngle;
    float mp1X = midX + mpDeltaX;
    float mp2X = midX - mpDeltaX;
    float mp1Y = midY + mpDeltaY;
    float mp2Y = midY - mpDeltaY;
    float mp1R2 = mp1X * mp1X + mp1Y * mp1Y;
    float mp2R2 = mp2X * mp2X + mp2Y * mp2Y;
    // math::min == std::min on the host (definitions/math.hpp:52-53): (b < a) ? b : a
    if (std::min(mp1R2, mp2R2) <= ((config.minHelixRadius - config.impactMax) *
                                   (config.minHelixRadius - config.impactMax))) {
        return false;
    }
    return true;
}

// doublet_finding_helper::transform_coordinates — doublet_finding_helper.hpp:218-272
LinCircle transform_coordinates(bool bottom, const Sp& sp1, const Sp& sp2) {
    const float xM = sp1.x;
    const float yM = sp1.y;
    const float zM = sp1.z;
    const float rM = sp1.radius();
    const float varianceZM = sp1.varZ;
    const float varianceRM = sp1.varR;
    float cosPhiM = xM / rM;
    float sinPhiM = yM / rM;
    float deltaX = sp2.x - xM;
    float deltaY = sp2.y - yM;
    float deltaZ = sp2.z - zM;
    float x = deltaX * cosPhiM + deltaY * sinPhiM;
    float y = deltaY * cosPhiM - deltaX * sinPhiM;
    float iDeltaR2 = 1.f / (deltaX * deltaX + deltaY * deltaY);
    float iDeltaR = std::sqrt(iDeltaR2);
    float cot_theta = deltaZ * iDeltaR;
    if (bottom) cot_theta = -cot_theta;
    LinCircle l;
    l.cotTheta = cot_theta;
    l.Zo = zM - rM * cot_theta;
    l.iDeltaR = iDeltaR;
    l.U = x * iDeltaR2;
    l.V = y * iDeltaR2;
    l.Er = ((varianceZM + sp2.varZ) + (cot_theta * cot_theta) * (varianceRM + sp2.varR)) *
           iDeltaR2;
    return l;
}

// triplet_finding_helper::isCompatible — triplet_finding_helper.hpp:42-133.
// *cut1 reports survival of the first scattering cut (:57-78), for counters only.
bool triplet_is_compatible(const Sp& spM, const LinCircle& lb, const LinCircle& lt,
                           const b200seed_finder_cfg& config, const float iSinTheta2,
                           const float scatteringInRegion2, float& curvature,
                           float& impact_parameter, bool* cut1) {
    float error2 = lt.Er + lb.Er +
                   2.f * (lb.cotTheta * lt.cotTheta * spM.varR + spM.varZ) * lb.iDeltaR *
                       lt.iDeltaR;
    float deltaCotTheta = lb.cotTheta - lt.cotTheta;
    float deltaCotTheta2 = deltaCotTheta * deltaCotTheta;
    float error{0.f};
    float dCotThetaMinusError2{0.f};
    *cut1 = false;
    if (deltaCotTheta2 - error2 > 0) {
        deltaCotTheta = std::fabs(deltaCotTheta);
        error = std::sqrt(error2);
        dCotThetaMinusError2 = deltaCotTheta2 + error2 - 2.f * deltaCotTheta * error;
        if (dCotThetaMinusError2 > scatteringInRegion2) return false;
    }
    *cut1 = true;
    float dU = lt.U - lb.U;
    if (dU == 0.f) return false;
    float A = (lt.V - lb.V) / dU;
    float S2 = 1.f + A * A;
    float B = lb.V - A * lb.U;
    float B2 = B * B;
    if (S2 < B2 * config.minHelixDiameter2) return false;
    float iHelixDiameter2 = B2 / S2;
    float pT2scatter = 4.f * iHelixDiameter2 * config.pT2perRadius;
    float pT = config.pTPerHelixRadius * std::sqrt(S2 / B2) / 2.f;
    if (pT > config.maxPtScattering) {
        float pTscatter = config.highland / config.maxPtScattering;
        pT2scatter = pTscatter * pTscatter;
    }
    float p2scatter = pT2scatter * iSinTheta2;
    if ((deltaCotTheta2 - error2 > 0.f) &&
        (dCotThetaMinusError2 > p2scatter * config.sigmaScattering * config.sigmaScattering)) {
        return false;
    }
    curvature = B / std::sqrt(S2);
    impact_parameter = std::fabs((A - B * spM.radius()) * spM.radius());
    if (impact_parameter > config.impactMax) return false;
    return true;
}

// sp_location — detail/singlet.hpp; triplet — detail/triplet.hpp
struct SpLoc {
    uint32_t bin_idx, sp_idx;
};
struct Triplet {
    SpLoc sp1, sp2, sp3;  // bottom, middle, top
    float curvature, weight, z_vertex;
};

struct Grid {
    Axis phi, z;
    std::vector<std::vector<uint32_t>> bins;  // serialised: phi + n_phi * z
    uint32_t nbins() const { return static_cast<uint32_t>(bins.size()); }
};

}  // namespace

// ===========================================================================
// Result object handed to the tests through a small C API (ctypes friendly)
// ===========================================================================
struct oracle_result {
    int status = 0;
    Axis phi_axis{}, z_axis{};
    std::vector<uint32_t> bin_offsets, bin_entries;
    // doublets of active middles, canonical order (seed_finding.cpp:69-95)
    std::vector<uint32_t> mb_mid, mb_other, mt_mid, mt_other;
    std::vector<float> mb_lc, mt_lc;  // 6 floats each: Zo,cotTheta,iDeltaR,Er,U,V
    // triplets, canonical order; weight = after compatible-seed bonus
    std::vector<uint32_t> tr_b, tr_m, tr_t;
    std::vector<float> tr_curv, tr_weight, tr_zv;
    // seeds, CPU output order
    std::vector<uint32_t> sd_b, sd_m, sd_t;
    std::vector<float> sd_q;
    // bound parameters per seed
    std::vector<b200seed_bound_params> params;
    // work counters
    uint64_t n_valid = 0, pair_tests = 0, stage1_bot = 0, stage1_top = 0, n_mid_bot_all = 0,
             n_mid_top_all = 0, n_active_middles = 0, n_mid_bot = 0, n_mid_top = 0,
             triplet_tests = 0, triplet_cut1 = 0, n_triplets = 0, max_q_middle = 0,
             weight_pairs = 0;
};

namespace {

struct Oracle {
    b200seed_finder_cfg fc;
    b200seed_filter_cfg flc;
    Axis phi_axis, z_axis;
    std::vector<Sp> sps;
    Grid grid;
    int dump = 0;
    uint32_t bin_lo = 0, bin_hi = 0xFFFFFFFFu;  // middle bins to process (bounded samples)
    oracle_result* res = nullptr;

    const Sp& at(const SpLoc& l) const { return sps[grid.bins[l.bin_idx][l.sp_idx]]; }
    uint32_t idx(const SpLoc& l) const { return grid.bins[l.bin_idx][l.sp_idx]; }

    // host::details::spacepoint_binning::operator() — core/src/seeding/spacepoint_binning.cpp:27-53
    void binning() {
        grid.phi = phi_axis;
        grid.z = z_axis;
        grid.bins.assign(static_cast<size_t>(phi_axis.n_bins) * z_axis.n_bins, {});
        for (uint32_t i = 0; i < sps.size(); ++i) {
            const Sp& sp = sps[i];
            if (is_valid_sp(fc, sp)) {
                const uint32_t bin_index =
                    circular_bin(phi_axis, sp.phi()) + phi_axis.n_bins * regular_bin(z_axis, sp.z);
                grid.bins[bin_index].push_back(i);
                ++res->n_valid;
            }
        }
    }

    // host::details::doublet_finding<T>::operator() — core/src/seeding/doublet_finding.hpp:51-110
    void doublet_finding(bool bottom, const SpLoc& middle, std::vector<SpLoc>& others,
                         std::vector<LinCircle>& lcs) {
        others.clear();
        lcs.clear();
        const Sp& middle_sp = at(middle);
        const std::vector<uint32_t> phi_bins =
            circular_zone(phi_axis, middle_sp.phi(), fc.neighbor_scope);
        const std::vector<uint32_t> z_bins = regular_zone(z_axis, middle_sp.z, fc.neighbor_scope);
        for (uint32_t phi_bin : phi_bins) {
            for (uint32_t z_bin : z_bins) {
                const uint32_t bin_idx = phi_bin + z_bin * phi_axis.n_bins;
                const std::vector<uint32_t>& sp_indices = grid.bins[bin_idx];
                uint32_t i = 0;
                for (uint32_t sp_index : sp_indices) {
                    const Sp& other_sp = sps[sp_index];
                    bool s1 = false;
                    if (bottom) ++res->pair_tests;
                    if (doublet_is_compatible(bottom, middle_sp, other_sp, fc, &s1)) {
                        others.push_back({bin_idx, i});
                        lcs.push_back(transform_coordinates(bottom, middle_sp, other_sp));
                    }
                    if (s1) ++(bottom ? res->stage1_bot : res->stage1_top);
                    ++i;
                }
            }
        }
    }

    // host::details::triplet_finding::operator() — core/src/seeding/triplet_finding.hpp:60-183
    void triplet_finding(const SpLoc& middle, const SpLoc& bottom, const LinCircle& mid_bot_lc,
                         const std::vector<SpLoc>& tops, const std::vector<LinCircle>& mid_top_lcs,
                         std::vector<Triplet>& result) {
        result.clear();
        const Sp& spM = at(middle);
        const float iSinTheta2 = 1.f + mid_bot_lc.cotTheta * mid_bot_lc.cotTheta;
        float scatteringInRegion2 = fc.maxScatteringAngle2 * iSinTheta2;
        scatteringInRegion2 *= fc.sigmaScattering * fc.sigmaScattering;
        float curvature, impact_parameter;
        for (size_t i = 0; i < tops.size(); ++i) {
            bool c1 = false;
            ++res->triplet_tests;
            const bool ok = triplet_is_compatible(spM, mid_bot_lc, mid_top_lcs[i], fc, iSinTheta2,
                                                  scatteringInRegion2, curvature,
                                                  impact_parameter, &c1);
            if (c1) ++res->triplet_cut1;
            if (!ok) continue;
            result.push_back({bottom, middle, tops[i], curvature,
                              -impact_parameter * flc.impactWeightFactor, mid_bot_lc.Zo});
        }
        // compatible-seed bonus (:107-179)
        for (size_t i = 0; i < result.size(); ++i) {
            Triplet& current_triplet = result[i];
            const float currentTop_r = at(current_triplet.sp3).radius();
            std::vector<float> compatibleSeedR;
            float lowerLimitCurv = current_triplet.curvature - flc.deltaInvHelixDiameter;
            float upperLimitCurv = current_triplet.curvature + flc.deltaInvHelixDiameter;
            for (size_t j = 0; j < result.size(); ++j) {
                if (i == j) continue;
                ++res->weight_pairs;
                const Triplet& other_triplet = result[j];
                const float otherTop_r = at(other_triplet.sp3).radius();
                const float deltaR = currentTop_r - otherTop_r;
                if (std::abs(deltaR) < flc.deltaRMin) continue;
                if (other_triplet.curvature < lowerLimitCurv) continue;
                if (other_triplet.curvature > upperLimitCurv) continue;
                bool newCompSeed = true;
                for (float previousDiameter : compatibleSeedR) {
                    if (std::abs(previousDiameter - otherTop_r) < flc.deltaRMin) {
                        newCompSeed = false;
                        break;
                    }
                }
                if (newCompSeed) {
                    compatibleSeedR.push_back(otherTop_r);
                    current_triplet.weight += flc.compatSeedWeight;
                }
                if (compatibleSeedR.size() >= flc.compatSeedLimit) break;
            }
        }
    }

    // seed_selecting_helper — core/include/traccc/seeding/seed_selecting_helper.hpp:28-80
    void seed_weight(const Sp& spB, const Sp& spT, float& triplet_weight) const {
        float weight = 0;
        if (spB.radius() > flc.good_spB_min_radius) weight = flc.good_spB_weight_increase;
        if (spT.radius() < flc.good_spT_max_radius) weight = flc.good_spT_weight_increase;
        triplet_weight += weight;
    }
    bool single_seed_cut(const Sp& spB, float triplet_weight) const {
        return !(spB.radius() > flc.good_spB_min_radius &&
                 triplet_weight < flc.good_spB_min_weight);
    }
    bool cut_per_middle_sp(const Sp& spB, const float weight) const {
        return (weight > flc.seed_min_weight || spB.radius() > flc.spB_min_radius);
    }

    // host::details::seed_filtering::operator() — core/src/seeding/seed_filtering.cpp:28-123
    // with traccc::details::triplet_sorter — detail/triplet_sorter.hpp:39-70
    void seed_filtering(std::vector<Triplet>& triplets) {
        std::vector<std::reference_wrapper<const Triplet>> passing;
        passing.reserve(triplets.size());
        for (Triplet& t : triplets) {
            const Sp& spB = at(t.sp1);
            const Sp& spT = at(t.sp3);
            seed_weight(spB, spT, t.weight);
            if (!single_seed_cut(spB, t.weight)) continue;
            passing.push_back(t);
        }
        std::sort(passing.begin(), passing.end(), [this](const Triplet& s1, const Triplet& s2) {
            if (s1.weight != s2.weight) return s1.weight > s2.weight;
            const Sp& spB1 = at(s1.sp1);
            const Sp& spT1 = at(s1.sp3);
            const Sp& spB2 = at(s2.sp1);
            const Sp& spT2 = at(s2.sp3);
            const float seed1_sum =
                spB1.y * spB1.y + spB1.z * spB1.z + spT1.y * spT1.y + spT1.z * spT1.z;
            const float seed2_sum =
                spB2.y * spB2.y + spB2.z * spB2.z + spT2.y * spT2.y + spT2.z * spT2.z;
            return seed1_sum > seed2_sum;
        });
        std::vector<std::reference_wrapper<const Triplet>> final_cuts;
        final_cuts.reserve(passing.size());
        if (passing.size() > 0u) {
            final_cuts.push_back(passing[0]);
            const size_t itLength =
                std::min(passing.size(), static_cast<size_t>(fc.maxSeedsPerSpM));
            for (size_t i = 1; i < itLength; ++i) {
                const Triplet& this_seed = passing[i].get();
                if (cut_per_middle_sp(at(this_seed.sp1), this_seed.weight))
                    final_cuts.push_back(passing[i]);
            }
        }
        size_t i = 0;
        for (const Triplet& t : final_cuts) {
            if (i++ >= fc.maxSeedsPerSpM) break;
            res->sd_b.push_back(idx(t.sp1));
            res->sd_m.push_back(idx(t.sp2));
            res->sd_t.push_back(idx(t.sp3));
            res->sd_q.push_back(static_cast<float>(t.weight));
        }
    }

    // host::details::seed_finding::operator() — core/src/seeding/seed_finding.cpp:58-121
    void seed_finding() {
        std::vector<SpLoc> bots, tops;
        std::vector<LinCircle> bot_lcs, top_lcs;
        std::vector<Triplet> triplets, for_mid_bot;
        for (uint32_t i = bin_lo; i < grid.nbins() && i < bin_hi; ++i) {
            const auto& middle_indices = grid.bins[i];
            for (uint32_t j = 0; j < middle_indices.size(); ++j) {
                SpLoc spM_location{i, j};
                doublet_finding(true, spM_location, bots, bot_lcs);
                res->n_mid_bot_all += bots.size();
                if (bots.empty()) continue;
                doublet_finding(false, spM_location, tops, top_lcs);
                res->n_mid_top_all += tops.size();
                if (tops.empty()) continue;
                ++res->n_active_middles;
                res->n_mid_bot += bots.size();
                res->n_mid_top += tops.size();
                res->max_q_middle =
                    std::max<uint64_t>(res->max_q_middle, uint64_t(bots.size()) * tops.size());
                if (dump) {
                    const uint32_t mi = middle_indices[j];
                    for (size_t k = 0; k < bots.size(); ++k) {
                        res->mb_mid.push_back(mi);
                        res->mb_other.push_back(idx(bots[k]));
                        const LinCircle& l = bot_lcs[k];
                        res->mb_lc.insert(res->mb_lc.end(),
                                          {l.Zo, l.cotTheta, l.iDeltaR, l.Er, l.U, l.V});
                    }
                    for (size_t k = 0; k < tops.size(); ++k) {
                        res->mt_mid.push_back(mi);
                        res->mt_other.push_back(idx(tops[k]));
                        const LinCircle& l = top_lcs[k];
                        res->mt_lc.insert(res->mt_lc.end(),
                                          {l.Zo, l.cotTheta, l.iDeltaR, l.Er, l.U, l.V});
                    }
                }
                triplets.clear();
                for (size_t k = 0; k < bots.size(); ++k) {
                    triplet_finding(spM_location, bots[k], bot_lcs[k], tops, top_lcs, for_mid_bot);
                    triplets.insert(triplets.end(), for_mid_bot.begin(), for_mid_bot.end());
                }
                res->n_triplets += triplets.size();
                if (dump) {
                    for (const Triplet& t : triplets) {
                        res->tr_b.push_back(idx(t.sp1));
                        res->tr_m.push_back(idx(t.sp2));
                        res->tr_t.push_back(idx(t.sp3));
                        res->tr_curv.push_back(t.curvature);
                        res->tr_weight.push_back(t.weight);
                        res->tr_zv.push_back(t.z_vertex);
                    }
                }
                seed_filtering(triplets);
            }
        }
    }
};

// --- algebra-plugins (array/cmath) vector ops, external, restated -----------------
struct V3 {
    float v[3];
};
inline V3 sub(const V3& a, const V3& b) {
    return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}};
}
inline float dot(const V3& a, const V3& b) {
    return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2];
}
inline float norm(const V3& a) {
    return std::sqrt(dot(a, a));
}
inline V3 normalize(const V3& a) {
    const float s = 1.f / norm(a);
    return {{s * a.v[0], s * a.v[1], s * a.v[2]}};
}
inline V3 cross(const V3& a, const V3& b) {
    return {{a.v[1] * b.v[2] - b.v[1] * a.v[2], a.v[2] * b.v[0] - b.v[2] * a.v[0],
             a.v[0] * b.v[1] - b.v[0] * a.v[1]}};
}
inline float perp2(float x, float y) {
    return std::sqrt(x * x + y * y);
}

// 4x4 determinant / inverse by cofactor expansion, m[col][row] — the "hard-coded" 4x4 forms of the
// array plugin's transform3 (restated from the published implementation; the library is absent).
inline float det44(const float (&m)[4][4]) {
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
inline void inverse44(const float (&m)[4][4], float (&i)[4][4]) {
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

// seed_to_bound_param_vector — core/include/traccc/seeding/track_params_estimation_helper.hpp:47-129
// + covariance of host::track_params_estimation::operator() —
// core/src/seeding/track_params_estimation.cpp:64-86 (== device/.../impl/estimate_track_params.ipp:60-87).
void estimate_params(const b200seed_tpe_cfg& cfg, const Sp& spB, const Sp& spM, const Sp& spT,
                     float loc0, float loc1, uint64_t surface, const float bf[3],
                     b200seed_bound_params& out) {
    std::memset(&out, 0, sizeof(out));
    const V3 p0{{spB.x, spB.y, spB.z}}, p1{{spM.x, spM.y, spM.z}}, p2{{spT.x, spT.y, spT.z}};
    const V3 bfield{{bf[0], bf[1], bf[2]}};
    V3 relVec = sub(p1, p0);
    V3 newZAxis = normalize(bfield);
    V3 newYAxis = normalize(cross(newZAxis, relVec));
    V3 newXAxis = cross(newYAxis, newZAxis);
    // transform3 trans(translation, x, y, z): columns x, y, z, t of a 4x4 matrix and its inverse by
    // cofactor expansion; point_to_local(p) = rotate(inverse, p) + translation column of the inverse
    // (:81-85). Rounds differently from R^T (p - t): q/p of a stiff track is ill-conditioned in the
    // local coordinates (B = v2 - A u2 cancels ~300-fold), the two forms differ by up to 4e-4 there.
    float tm[4][4], ti[4][4];
    for (int r = 0; r < 3; ++r) {
        tm[0][r] = newXAxis.v[r];
        tm[1][r] = newYAxis.v[r];
        tm[2][r] = newZAxis.v[r];
        tm[3][r] = p0.v[r];
    }
    tm[0][3] = tm[1][3] = tm[2][3] = 0.f;
    tm[3][3] = 1.f;
    inverse44(tm, ti);
    auto rotate = [](const float (&m)[4][4], const V3& v) {
        return V3{{m[0][0] * v.v[0] + m[1][0] * v.v[1] + m[2][0] * v.v[2],
                   m[0][1] * v.v[0] + m[1][1] * v.v[1] + m[2][1] * v.v[2],
                   m[0][2] * v.v[0] + m[1][2] * v.v[1] + m[2][2] * v.v[2]}};
    };
    auto to_local = [&](const V3& p) {
        const V3 rg = rotate(ti, p);
        return V3{{rg.v[0] + ti[3][0], rg.v[1] + ti[3][1], rg.v[2] + ti[3][2]}};
    };
    const V3 local1 = to_local(p1);
    const V3 local2 = to_local(p2);
    // uv_transform (:29-36)
    auto uv = [](float x, float y, float o[2]) {
        float denominator = x * x + y * y;
        o[0] = x / denominator;
        o[1] = y / denominator;
    };
    float uv1[2], uv2[2];
    uv(local1.v[0], local1.v[1], uv1);
    uv(local2.v[0], local2.v[1], uv2);
    float A = (uv2[1] - uv1[1]) / (uv2[0] - uv1[0]);
    float B = uv2[1] - A * uv2[0];
    float R = -perp2(1.f, A) / (2.f * B);
    float invTanTheta =
        local2.v[2] / (2.f * R * std::asin(perp2(local2.v[0], local2.v[1]) / (2.f * R)));
    V3 transDirection{{1.f, A, perp2(1.f, A) * invTanTheta}};
    const V3 nd = normalize(transDirection);
    // transform3::rotate(trans._data, v)
    const V3 direction = rotate(tm, nd);
    const float phi = std::atan2(direction.v[1], direction.v[0]);
    const float theta = std::atan2(perp2(direction.v[0], direction.v[1]), direction.v[2]);
    float qOverPt = 1.f / (R * norm(bfield));
    const float qop = qOverPt / perp2(1.f, invTanTheta);
    out.surface_link = surface;
    out.vec[0] = loc0;
    out.vec[1] = loc1;
    out.vec[2] = phi;
    out.vec[3] = theta;
    out.vec[4] = qop;
    out.vec[5] = 0.f;
    for (size_t j = 0; j < 6; ++j) {
        float var = cfg.initial_sigma[j] * cfg.initial_sigma[j];
        if (j == 4) {
            float var_theta = out.cov[3 * 6 + 3];
            var += std::pow(cfg.initial_sigma_qopt * std::sin(theta), 2.f);
            var += std::pow(cfg.initial_sigma_pt_rel * qop, 2.f);
            var += var_theta * std::pow(qop / std::tan(theta), 2.f);
        }
        var *= cfg.initial_inflation[j];
        out.cov[j * 6 + j] = var;
    }
}

}  // namespace

// ===========================================================================
// C API
// ===========================================================================
extern "C" {

// seedfinder_config defaults + setup() — seeding_config.hpp:17-139
void oracle_finder_cfg_setup(b200seed_finder_cfg* c) {
    c->highland = 13.6f * unit_MeV * std::sqrt(c->radLengthPerSeed) *
                  (1.f + 0.038f * std::log(c->radLengthPerSeed));
    float maxScatteringAngle = c->highland / c->minPt;
    c->maxScatteringAngle2 = maxScatteringAngle * maxScatteringAngle;
    c->pTPerHelixRadius = c->bFieldInZ;
    c->minHelixDiameter2 = std::pow(c->minPt * 2.f / c->pTPerHelixRadius, 2.f);
    c->minHelixRadius = std::sqrt(c->minHelixDiameter2) / 2.f;
    c->pT2perRadius = std::pow(c->highland / c->pTPerHelixRadius, 2.f);
}
void oracle_finder_cfg_defaults(b200seed_finder_cfg* c) {
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
    oracle_finder_cfg_setup(c);
}
// spacepoint_grid_config(const seedfinder_config&) — seeding_config.hpp:145-156
void oracle_grid_cfg_from_finder(const b200seed_finder_cfg* f, b200seed_grid_cfg* g) {
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
void oracle_filter_cfg_defaults(b200seed_filter_cfg* c) {
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
void oracle_tpe_cfg_defaults(b200seed_tpe_cfg* c) {
    const float s[6] = {1.f * unit_mm,     1.f * unit_mm,           1.f * unit_degree,
                        1.f * unit_degree, 0.f * 1.f / unit_GeV, 1.f * unit_ns};
    const float infl[6] = {1.f, 1.f, 1.f, 1.f, 1.f, 100.f};
    for (int i = 0; i < 6; ++i) {
        c->initial_sigma[i] = s[i];
        c->initial_inflation[i] = infl[i];
    }
    c->initial_sigma_qopt = 0.1f * 1.f / unit_GeV;
    c->initial_sigma_pt_rel = 0.1f;
}

void oracle_use_libm_atan2(int on) {
    g_use_libm_atan2 = (on != 0);
}
float oracle_atan2f_fdlibm(float y, float x) {
    return fd_atan2f(y, x);
}
// Number of mismatches between the fdlibm restatement and this box's libm atan2f on n
// pseudo-random points of [-scale, scale]^2 (xorshift64*; deterministic).
uint64_t oracle_selftest_atan2f(uint64_t n, float scale, uint64_t seed) {
    uint64_t s = seed ? seed : 0x9E3779B97F4A7C15ull, bad = 0;
    auto next = [&]() {
        s ^= s >> 12;
        s ^= s << 25;
        s ^= s >> 27;
        return s * 0x2545F4914F6CDD1Dull;
    };
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t r = next();
        const float x = (static_cast<float>(static_cast<uint32_t>(r)) / 4294967296.f * 2.f - 1.f) *
                        scale;
        const float y =
            (static_cast<float>(static_cast<uint32_t>(r >> 32)) / 4294967296.f * 2.f - 1.f) * scale;
        if (f2u(fd_atan2f(y, x)) != f2u(std::atan2(y, x))) ++bad;
    }
    return bad;
}

// Axis known-answer access (tests/cpu/test_axis.cpp)
uint32_t oracle_axis_regular_bin(uint32_t n, float mn, float mx, float v) {
    return regular_bin(Axis{n, mn, mx}, v);
}
uint32_t oracle_axis_circular_bin(uint32_t n, float mn, float mx, float v) {
    return circular_bin(Axis{n, mn, mx}, v);
}
uint32_t oracle_axis_circular_remap(uint32_t n, float mn, float mx, uint32_t ibin, int shood) {
    return circular_remap(Axis{n, mn, mx}, ibin, shood);
}
void oracle_axis_regular_range(uint32_t n, float mn, float mx, float v, uint32_t n0, uint32_t n1,
                               uint32_t out[2]) {
    const uint32_t nh[2] = {n0, n1};
    regular_range(Axis{n, mn, mx}, v, nh, out);
}
void oracle_axis_circular_range(uint32_t n, float mn, float mx, float v, uint32_t n0, uint32_t n1,
                                uint32_t out[2]) {
    const uint32_t nh[2] = {n0, n1};
    circular_range(Axis{n, mn, mx}, v, nh, out);
}
// zone: writes up to cap entries, returns the sequence length
uint32_t oracle_axis_zone(int circular, uint32_t n, float mn, float mx, float v, uint32_t n0,
                          uint32_t n1, uint32_t* out, uint32_t cap) {
    const uint32_t nh[2] = {n0, n1};
    const std::vector<uint32_t> z =
        circular ? circular_zone(Axis{n, mn, mx}, v, nh) : regular_zone(Axis{n, mn, mx}, v, nh);
    for (uint32_t i = 0; i < z.size() && i < cap; ++i) out[i] = z[i];
    return static_cast<uint32_t>(z.size());
}
int oracle_get_axes(const b200seed_grid_cfg* g, uint32_t* n_phi, float* phi_min, float* phi_max,
                    uint32_t* n_z, float* z_min, float* z_max) {
    Axis p, z;
    if (get_axes(*g, p, z) != 0) return -1;
    *n_phi = p.n_bins;
    *phi_min = p.min;
    *phi_max = p.max;
    *n_z = z.n_bins;
    *z_min = z.min;
    *z_max = z.max;
    return 0;
}

// Pair/triplet cut probes (used to cross-check against oracle/_ref)
int oracle_doublet_is_compatible(int bottom, const float m[5], const float o[5],
                                 const b200seed_finder_cfg* c) {
    bool s1;
    return doublet_is_compatible(bottom != 0, Sp{m[0], m[1], m[2], m[3], m[4]},
                                 Sp{o[0], o[1], o[2], o[3], o[4]}, *c, &s1)
               ? 1
               : 0;
}
void oracle_transform_coordinates(int bottom, const float m[5], const float o[5], float lc[6]) {
    const LinCircle l = transform_coordinates(bottom != 0, Sp{m[0], m[1], m[2], m[3], m[4]},
                                              Sp{o[0], o[1], o[2], o[3], o[4]});
    lc[0] = l.Zo;
    lc[1] = l.cotTheta;
    lc[2] = l.iDeltaR;
    lc[3] = l.Er;
    lc[4] = l.U;
    lc[5] = l.V;
}
int oracle_triplet_is_compatible(const float m[5], const float lb[6], const float lt[6],
                                 const b200seed_finder_cfg* c, float out[2]) {
    const LinCircle b{lb[0], lb[1], lb[2], lb[3], lb[4], lb[5]};
    const LinCircle t{lt[0], lt[1], lt[2], lt[3], lt[4], lt[5]};
    const float iSinTheta2 = 1.f + b.cotTheta * b.cotTheta;
    float scatteringInRegion2 = c->maxScatteringAngle2 * iSinTheta2;
    scatteringInRegion2 *= c->sigmaScattering * c->sigmaScattering;
    bool c1;
    out[0] = out[1] = 0.f;
    return triplet_is_compatible(Sp{m[0], m[1], m[2], m[3], m[4]}, b, t, *c, iSinTheta2,
                                 scatteringInRegion2, out[0], out[1], &c1)
               ? 1
               : 0;
}

// seed_selecting_helper probe: returns the updated weight; flags = {single_seed_cut, cut_per_middle_sp}
float oracle_seed_select(const b200seed_filter_cfg* c, const float b[5], const float t[5],
                         float weight, int flags[2]) {
    Oracle o;
    o.flc = *c;
    const Sp spB{b[0], b[1], b[2], b[3], b[4]}, spT{t[0], t[1], t[2], t[3], t[4]};
    float w = weight;
    o.seed_weight(spB, spT, w);
    flags[0] = o.single_seed_cut(spB, w) ? 1 : 0;
    flags[1] = o.cut_per_middle_sp(spB, w) ? 1 : 0;
    return w;
}
float oracle_sp_radius(const float p[5]) {
    return Sp{p[0], p[1], p[2], p[3], p[4]}.radius();
}
float oracle_sp_phi(const float p[5]) {
    return Sp{p[0], p[1], p[2], p[3], p[4]}.phi();
}

// host::seeding_algorithm::operator() — core/src/seeding/seeding_algorithm.cpp:24-28.
// dump != 0 additionally records grid, doublets and triplets.
// oracle_run_bins restricts the middle-spacepoint loop (seed_finding.cpp:69) to the bins
// [bin_lo, bin_hi): a bounded, exactly proportional sample of an event for CPU timing.
oracle_result* oracle_run_bins(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                               const b200seed_filter_cfg* filter, uint32_t n_sp, const float* xyz,
                               const float* var_z, const float* var_r, int dump, uint32_t bin_lo,
                               uint32_t bin_hi);
oracle_result* oracle_run(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                          const b200seed_filter_cfg* filter, uint32_t n_sp, const float* xyz,
                          const float* var_z, const float* var_r, int dump) {
    return oracle_run_bins(finder, grid, filter, n_sp, xyz, var_z, var_r, dump, 0, 0xFFFFFFFFu);
}
oracle_result* oracle_run_bins(const b200seed_finder_cfg* finder, const b200seed_grid_cfg* grid,
                               const b200seed_filter_cfg* filter, uint32_t n_sp, const float* xyz,
                               const float* var_z, const float* var_r, int dump, uint32_t bin_lo,
                               uint32_t bin_hi) {
    oracle_result* res = new oracle_result();
    Oracle o;
    o.bin_lo = bin_lo;
    o.bin_hi = bin_hi;
    o.fc = *finder;
    o.flc = *filter;
    o.dump = dump;
    o.res = res;
    if (get_axes(*grid, o.phi_axis, o.z_axis) != 0) {
        res->status = -1;
        return res;
    }
    res->phi_axis = o.phi_axis;
    res->z_axis = o.z_axis;
    o.sps.resize(n_sp);
    for (uint32_t i = 0; i < n_sp; ++i)
        o.sps[i] = Sp{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], var_z ? var_z[i] : 0.f,
                      var_r ? var_r[i] : 0.f};
    o.binning();
    if (dump) {
        res->bin_offsets.push_back(0);
        for (const auto& b : o.grid.bins) {
            res->bin_entries.insert(res->bin_entries.end(), b.begin(), b.end());
            res->bin_offsets.push_back(static_cast<uint32_t>(res->bin_entries.size()));
        }
    }
    o.seed_finding();
    return res;
}

// host::track_params_estimation::operator() on the seeds held by `res`.
void oracle_estimate_params(oracle_result* res, const b200seed_tpe_cfg* cfg, const float* xyz,
                            const uint32_t* sp_meas_index_1, const float* meas_local,
                            const uint64_t* meas_surface, const float bfield[3]) {
    const size_t n = res->sd_b.size();
    res->params.resize(n);
    auto sp = [&](uint32_t i) { return Sp{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f, 0.f}; };
    for (size_t i = 0; i < n; ++i) {
        const uint32_t b = res->sd_b[i];
        const uint32_t mi = sp_meas_index_1 ? sp_meas_index_1[b] : b;
        estimate_params(*cfg, sp(b), sp(res->sd_m[i]), sp(res->sd_t[i]),
                        meas_local ? meas_local[2 * mi] : 0.f,
                        meas_local ? meas_local[2 * mi + 1] : 0.f,
                        meas_surface ? meas_surface[mi] : 0, bfield, res->params[i]);
    }
}
// Field vector of an inhomogeneous field at a point: the reference's inhom_bfield_backend_t =
// covfie affine<linear<clamp<strided<array<float3>>>>> (core/include/traccc/bfield/
// magnetic_field_types.hpp:44-48). covfie 0.15.4 is third-party and absent: restated from its
// published semantics ("parity unpinned" at the last ulp, like the detray arithmetic):
//   affine  : c_i = ((A_i0 x + A_i1 y) + A_i2 z) + A_i3
//   linear  : i = floor(c), a = c - i, sum over the 8 corners n = 4 dx + 2 dy + dz of
//             ((wx * wy) * wz) * f(corner), accumulated in that order
//   clamp   : corner indices clamped into [0, size - 1]
//   strided : row-major, point (i, j, k) at (i * size[1] + j) * size[2] + k
void oracle_field_at(const b200seed_field_grid* fg, const float p[3], float out[3]) {
    float c[3];
    for (int i = 0; i < 3; ++i)
        c[i] = ((fg->affine[4 * i] * p[0] + fg->affine[4 * i + 1] * p[1]) +
                fg->affine[4 * i + 2] * p[2]) +
               fg->affine[4 * i + 3];
    long i0[3], i1[3];
    float w0[3], w1[3];
    for (int k = 0; k < 3; ++k) {
        const float fl = std::floor(c[k]);
        w1[k] = c[k] - fl;
        w0[k] = 1.f - w1[k];
        const float hi = static_cast<float>(fg->size[k] - 1u);
        auto clampf = [&](float v) { return (v >= 0.f) ? ((v <= hi) ? v : hi) : 0.f; };
        i0[k] = static_cast<long>(clampf(fl));
        i1[k] = static_cast<long>(clampf(fl + 1.f));
    }
    float r[3] = {0.f, 0.f, 0.f};
    for (int n = 0; n < 8; ++n) {
        const long ix = (n & 4) ? i1[0] : i0[0], iy = (n & 2) ? i1[1] : i0[1],
                   iz = (n & 1) ? i1[2] : i0[2];
        const float w = (((n & 4) ? w1[0] : w0[0]) * ((n & 2) ? w1[1] : w0[1])) *
                        ((n & 1) ? w1[2] : w0[2]);
        const float* f = fg->data + 3 * ((size_t(ix) * fg->size[1] + size_t(iy)) * fg->size[2] + size_t(iz));
        for (int q = 0; q < 3; ++q) r[q] = r[q] + w * f[q];
    }
    out[0] = r[0], out[1] = r[1], out[2] = r[2];
}

// device::estimate_track_params with an inhomogeneous field: the field is sampled at the
// bottom spacepoint of every seed (device/common/.../impl/estimate_track_params.ipp:45-50).
// fg->data is HOST memory here.
void oracle_estimate_params_inhom(const b200seed_tpe_cfg* cfg, uint32_t n_seeds, const uint32_t* b,
                                  const uint32_t* m, const uint32_t* t, const float* xyz,
                                  const uint32_t* sp_meas_index_1, const float* meas_local,
                                  const uint64_t* meas_surface, const b200seed_field_grid* fg,
                                  b200seed_bound_params* out) {
    auto sp = [&](uint32_t i) { return Sp{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f, 0.f}; };
    for (uint32_t i = 0; i < n_seeds; ++i) {
        float bf[3];
        oracle_field_at(fg, xyz + 3 * size_t(b[i]), bf);
        const uint32_t mi = sp_meas_index_1 ? sp_meas_index_1[b[i]] : b[i];
        estimate_params(*cfg, sp(b[i]), sp(m[i]), sp(t[i]), meas_local ? meas_local[2 * mi] : 0.f,
                        meas_local ? meas_local[2 * mi + 1] : 0.f,
                        meas_surface ? meas_surface[mi] : 0, bf, out[i]);
    }
}

// Stand-alone parameter estimation for explicit seeds (test_track_params_estimation.cpp)
void oracle_estimate_params_for(const b200seed_tpe_cfg* cfg, uint32_t n_seeds, const uint32_t* b,
                                const uint32_t* m, const uint32_t* t, const float* xyz,
                                const uint32_t* sp_meas_index_1, const float* meas_local,
                                const uint64_t* meas_surface, const float bfield[3],
                                b200seed_bound_params* out) {
    oracle_result tmp;
    tmp.sd_b.assign(b, b + n_seeds);
    tmp.sd_m.assign(m, m + n_seeds);
    tmp.sd_t.assign(t, t + n_seeds);
    oracle_estimate_params(&tmp, cfg, xyz, sp_meas_index_1, meas_local, meas_surface, bfield);
    for (uint32_t i = 0; i < n_seeds; ++i) out[i] = tmp.params[i];
}

int oracle_status(const oracle_result* r) {
    return r->status;
}
void oracle_free(oracle_result* r) {
    delete r;
}
// sizes: [n_bins+1, n_entries, n_mb, n_mt, n_triplets_dumped, n_seeds, n_params]
void oracle_sizes(const oracle_result* r, uint64_t out[7]) {
    out[0] = r->bin_offsets.size();
    out[1] = r->bin_entries.size();
    out[2] = r->mb_mid.size();
    out[3] = r->mt_mid.size();
    out[4] = r->tr_b.size();
    out[5] = r->sd_b.size();
    out[6] = r->params.size();
}
// counters: see oracle_result
void oracle_counters(const oracle_result* r, uint64_t out[14]) {
    out[0] = r->n_valid;
    out[1] = r->pair_tests;
    out[2] = r->stage1_bot;
    out[3] = r->stage1_top;
    out[4] = r->n_mid_bot_all;
    out[5] = r->n_mid_top_all;
    out[6] = r->n_active_middles;
    out[7] = r->n_mid_bot;
    out[8] = r->n_mid_top;
    out[9] = r->triplet_tests;
    out[10] = r->triplet_cut1;
    out[11] = r->n_triplets;
    out[12] = r->max_q_middle;
    out[13] = r->weight_pairs;
}
void oracle_axes(const oracle_result* r, uint32_t* n_phi, uint32_t* n_z) {
    *n_phi = r->phi_axis.n_bins;
    *n_z = r->z_axis.n_bins;
}
#define COPY(vec, dst) \
    if (dst) std::memcpy(dst, (vec).data(), (vec).size() * sizeof((vec)[0]))
void oracle_copy_grid(const oracle_result* r, uint32_t* offsets, uint32_t* entries) {
    COPY(r->bin_offsets, offsets);
    COPY(r->bin_entries, entries);
}
void oracle_copy_doublets(const oracle_result* r, uint32_t* mb_mid, uint32_t* mb_other,
                          float* mb_lc, uint32_t* mt_mid, uint32_t* mt_other, float* mt_lc) {
    COPY(r->mb_mid, mb_mid);
    COPY(r->mb_other, mb_other);
    COPY(r->mb_lc, mb_lc);
    COPY(r->mt_mid, mt_mid);
    COPY(r->mt_other, mt_other);
    COPY(r->mt_lc, mt_lc);
}
void oracle_copy_triplets(const oracle_result* r, uint32_t* b, uint32_t* m, uint32_t* t,
                          float* curv, float* weight, float* zv) {
    COPY(r->tr_b, b);
    COPY(r->tr_m, m);
    COPY(r->tr_t, t);
    COPY(r->tr_curv, curv);
    COPY(r->tr_weight, weight);
    COPY(r->tr_zv, zv);
}
void oracle_copy_seeds(const oracle_result* r, uint32_t* b, uint32_t* m, uint32_t* t, float* q) {
    COPY(r->sd_b, b);
    COPY(r->sd_m, m);
    COPY(r->sd_t, t);
    COPY(r->sd_q, q);
}
void oracle_copy_params(const oracle_result* r, b200seed_bound_params* p) {
    COPY(r->params, p);
}

// Spacepoint formation — the host loop of
// core/src/seeding/silicon_pixel_spacepoint_formation.hpp:33-62 with
// details::is_valid_measurement / fill_pixel_spacepoint
// (core/include/traccc/seeding/impl/spacepoint_formation.ipp:17-47): one spacepoint per 2D
// measurement, in measurement order, global = surface.local_to_global(local), zero variances,
// measurement_index_1 = i, measurement_index_2 = INVALID. detray's local_to_global for a planar
// surface is transform3::point_to_global({l0, l1, 0}) = rotation * p + translation (third-party
// code, absent here: restated as (x_axis * l0 + y_axis * l1) + translation per component —
// "parity unpinned" at the last ulp, see the file header). dim == NULL: all 2D. A surface index
// outside the table skips the measurement (the reference would read out of bounds).
// Returns the number of spacepoints.
uint32_t oracle_form_spacepoints(uint32_t n_meas, const float* local, const uint32_t* dim,
                                 const uint32_t* surface_index, const b200seed_surface* surfaces,
                                 uint32_t n_surfaces, float* xyz, float* var_z, float* var_r,
                                 uint32_t* mi1, uint32_t* mi2) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < n_meas; ++i) {
        if (dim && dim[i] != 2u) continue;
        if (surface_index[i] >= n_surfaces) continue;
        const b200seed_surface& S = surfaces[surface_index[i]];
        const float l0 = local[2 * size_t(i)], l1 = local[2 * size_t(i) + 1];
        for (int k = 0; k < 3; ++k)
            xyz[3 * size_t(n) + k] = (S.x_axis[k] * l0 + S.y_axis[k] * l1) + S.translation[k];
        if (var_z) var_z[n] = 0.f;
        if (var_r) var_r[n] = 0.f;
        if (mi1) mi1[n] = i;
        if (mi2) mi2[n] = 0xFFFFFFFFu;
        ++n;
    }
    return n;
}

}  // extern "C"
