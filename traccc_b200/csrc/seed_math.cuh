// seed_math.cuh — the cut arithmetic of the seeding path as host/device inline functions.
//
// Everything here must round exactly like the reference's *CPU* build (x86-64-v2: SSE
// scalar float, no FMA, IEEE div/sqrt), so this translation unit is compiled with
// -fmad=false and default -prec-div/-prec-sqrt/-ftz=false; expressions keep the
// reference's operand order and association. The same functions are also compiled for
// the host (b200seed_host_probe_* in b200seed_api.cu) so that the CPU test-suite can
// compare them with the oracle without a GPU.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200seed {

// Flattened device copy of the three reference configs + derived constants. Products
// of configuration values that the reference evaluates at every use are evaluated once
// on the host, in float, with the same association.
struct DevCfg {
    // is_valid_sp (spacepoint_binning_helper.hpp:112-127)
    float zMin, zMax, phiMin, phiMax, beamX, beamY;
    unsigned long long numRBins;  // size_t(rMax + |beamPos|), seeding_config.hpp:110-113
    // axes (get_axes, spacepoint_binning_helper.hpp:22-110)
    uint32_t nPhi, nZ;
    float phiAxisMin, phiAxisMax, zAxisMin, zAxisMax;
    uint32_t scope0, scope1;  // neighbor_scope
    // doublet cuts (doublet_finding_helper.hpp:77-213)
    float deltaRMin, deltaRMax, cotThetaMax, collisionRegionMin, collisionRegionMax, deltaZMax;
    float minHelixRadius2;      // minHelixRadius * minHelixRadius
    float helixImpactMargin2;   // (minHelixRadius - impactMax)^2
    // triplet cuts (triplet_finding.hpp:77-82, triplet_finding_helper.hpp:51-132)
    float maxScatteringAngle2, sigmaScattering, sigmaScattering2, minHelixDiameter2, pT2perRadius,
        pTPerHelixRadius, maxPtScattering, pT2scatterMax /* (highland/maxPtScattering)^2 */,
        impactMax;
    // filter (seed_filtering.cpp, seed_selecting_helper.hpp, triplet_finding.hpp:107-179)
    float impactWeightFactor, deltaInvHelixDiameter, compatSeedWeight, filterDeltaRMin;
    uint32_t compatSeedLimit, maxSeedsPerSpM;
    float good_spB_min_radius, good_spB_weight_increase, good_spT_max_radius,
        good_spT_weight_increase, good_spB_min_weight, seed_min_weight, spB_min_radius;
    // 1 if every valid spacepoint is so far inside the minimum helix radius that the magnitude
    // guards of doublet_stage2_fast can never fire (set on the host, see fill_devcfg)
    uint32_t fast_bounded;
};

B200_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union {
        float f;
        uint32_t u;
    } c;
    c.f = f;
    return c.u;
#endif
}
B200_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union {
        float f;
        uint32_t u;
    } c;
    c.u = u;
    return c.f;
#endif
}
B200_HD float absf(float x) {
    return u2f(f2u(x) & 0x7fffffffu);
}

// ---------------------------------------------------------------------------
// atan2f exactly as glibc 2.39 / Sun fdlibm compute it (e_atan2f.c, s_atanf.c):
// CUDA's atan2f differs in the last ulp, which would move ~3e-6 of the spacepoints
// into the neighbouring phi bin. Pure +,-,*,/ on floats.
// ---------------------------------------------------------------------------
B200_HD float fd_atanf(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f,
                             1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f,
                             7.5497894159e-08f};
    const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f,
                aT3 = -1.1111110449e-01f, aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f,
                aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f, aT8 = 4.9768779427e-02f,
                aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
    const int32_t hx = static_cast<int32_t>(f2u(x));
    const int32_t ix = hx & 0x7fffffff;
    int id;
    float hi = 0.f, lo = 0.f;
    if (ix >= 0x4c800000) {
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return atanhi[3] + atanlo[3];
        return -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = absf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) {
                id = 0;
                hi = atanhi[0];
                lo = atanlo[0];
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                id = 1;
                hi = atanhi[1];
                lo = atanlo[1];
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000) {
                id = 2;
                hi = atanhi[2];
                lo = atanlo[2];
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else {
                id = 3;
                hi = atanhi[3];
                lo = atanlo[3];
                x = -1.0f / x;
            }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return (hx < 0) ? -r : r;
}

B200_HD float fd_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = 7.8539818525e-01f;
    const float pi_o_2 = 1.5707963705e+00f;
    const float pi = 3.1415927410e+00f;
    const float pi_lo = -8.7422776573e-08f;
    const int32_t hx = static_cast<int32_t>(f2u(x));
    const int32_t ix = hx & 0x7fffffff;
    const int32_t hy = static_cast<int32_t>(f2u(y));
    const int32_t iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return fd_atanf(y);
    const int32_t m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        if (m < 2) return y;
        return (m == 2) ? pi + tiny : -pi - tiny;
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            if (m == 0) return pi_o_4 + tiny;
            if (m == 1) return -pi_o_4 - tiny;
            if (m == 2) return 3.0f * pi_o_4 + tiny;
            return -3.0f * pi_o_4 - tiny;
        }
        if (m == 0) return 0.0f;
        if (m == 1) return -0.0f;
        if (m == 2) return pi + tiny;
        return -pi - tiny;
    }
    if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60)
        z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60)
        z = 0.0f;
    else
        z = fd_atanf(absf(y / x));
    if (m == 0) return z;
    if (m == 1) return u2f(f2u(z) ^ 0x80000000u);
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

// ---------------------------------------------------------------------------
// Axes (core/include/traccc/seeding/grids/axis.hpp)
// ---------------------------------------------------------------------------
// raw bin of regular::bin / circular::bin / regular::range (:88-89, :110-111, :278-279)
B200_HD int axis_ibin(float mn, float mx, uint32_t n, float v) {
    return static_cast<int>((v - mn) / (mx - mn) * static_cast<float>(n));
}
// axis2::circular::bin (:277-289)
B200_HD uint32_t circular_bin(float mn, float mx, uint32_t n, float v) {
    const int ibin = axis_ibin(mn, mx, n, v);
    if (ibin >= 0 && ibin < static_cast<int>(n)) return static_cast<uint32_t>(ibin);
    if (ibin < 0) return n + static_cast<uint32_t>(ibin);
    return static_cast<uint32_t>(ibin) - n;
}
// axis2::regular::bin (:87-99)
B200_HD uint32_t regular_bin(float mn, float mx, uint32_t n, float v) {
    const int ibin = axis_ibin(mn, mx, n, v);
    if (ibin >= 0 && ibin < static_cast<int>(n)) return static_cast<uint32_t>(ibin);
    if (ibin < 0) return 0u;
    return n - 1u;
}
// axis2::circular::remap (:394-404)
B200_HD uint32_t circular_remap(uint32_t n, uint32_t ibin, int shood) {
    const int opt_bin = static_cast<int>(ibin) + shood;
    if (opt_bin >= 0 && opt_bin < static_cast<int>(n)) return static_cast<uint32_t>(opt_bin);
    if (opt_bin < 0) return static_cast<uint32_t>(static_cast<int>(n) + opt_bin);
    return static_cast<uint32_t>(opt_bin) - n;
}

// spacepoint radius (edm/impl/spacepoint_collection.ipp:51-57)
B200_HD float sp_radius(float x, float y) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x * x + y * y);
#else
    return __builtin_sqrtf(x * x + y * y);
#endif
}
B200_HD float sqrt_rn(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}

// is_valid_sp (spacepoint_binning_helper.hpp:112-127) + bin index
// (core/src/seeding/spacepoint_binning.cpp:47-48). Returns the serialised bin
// phi + nPhi * z, or 0xFFFFFFFF for a rejected spacepoint.
B200_HD uint32_t sp_bin(const DevCfg& c, float x, float y, float z) {
    if (z > c.zMax || z < c.zMin) return 0xFFFFFFFFu;
    const float phi = fd_atan2f(y, x);
    if (phi > c.phiMax || phi < c.phiMin) return 0xFFFFFFFFu;
    const float px = x - c.beamX, py = y - c.beamY;
    const float perp = sqrt_rn(px * px + py * py);
    if (!(static_cast<unsigned long long>(perp) < c.numRBins)) return 0xFFFFFFFFu;
    return circular_bin(c.phiAxisMin, c.phiAxisMax, c.nPhi, phi) +
           c.nPhi * regular_bin(c.zAxisMin, c.zAxisMax, c.nZ, z);
}

// ---------------------------------------------------------------------------
// Doublet cuts (core/include/traccc/seeding/doublet_finding_helper.hpp:51-216)
// ---------------------------------------------------------------------------
// First block of cuts (:59-84) for BOTH directions at once. With
//   dR = rM - r2, dz = zM - z2, zo = zM*dR - rM*dz        (the <bottom> operands)
// the <top> operands are exactly -dR, -dz, -zo (IEEE negation is exact), and since
// deltaR must exceed deltaRMin > 0 at most one direction can pass.
// Returns 0 = neither, 1 = bottom doublet, 2 = top doublet.
B200_HD int doublet_stage1(const DevCfg& c, float rM, float zM, float r2, float z2) {
    const float dR = rM - r2;
    const float dz = zM - z2;
    const float zo = zM * dR - rM * dz;
    const bool top = dR < 0.f;
    const float deltaR = top ? -dR : dR;
    const float zOrigin = top ? -zo : zo;
    const float acot = absf(dz);
    if ((deltaR >= c.deltaRMax) || (deltaR <= c.deltaRMin) || (acot >= c.cotThetaMax * deltaR) ||
        (zOrigin <= c.collisionRegionMin * deltaR) || (zOrigin >= c.collisionRegionMax * deltaR) ||
        (acot >= c.deltaZMax)) {
        return 0;
    }
    return top ? 2 : 1;
}

// Minimum-helix-radius cut (:120-213); direction independent. NaNs (deltaX == 0 or
// deltaY == 0) must behave as on the host, where math::min is std::min: (b < a) ? b : a.
B200_HD bool doublet_stage2(const DevCfg& c, float x1, float y1, float x2, float y2) {
    const float midX = 0.5f * (x1 + x2);
    const float midY = 0.5f * (y1 + y2);
    const float slope = (y2 - y1) / (x2 - x1);
    const float deltaX = x2 - x1;
    const float deltaY = y2 - y1;
    const float deltaXY2 = deltaX * deltaX + deltaY * deltaY;
    const float sagittaLength = sqrt_rn(c.minHelixRadius2 - deltaXY2 / 4.f);
    const float denom = sqrt_rn((slope * slope + 1) / (slope * slope));
    const float cosCentralAngle = 1.f / denom;
    const float sinCentralAngle = -1.f / (slope * denom);
    const float mpDeltaX = sagittaLength * cosCentralAngle;
    const float mpDeltaY = sagittaLength * sinCentralAngle;
    const float mp1X = midX + mpDeltaX;
    const float mp2X = midX - mpDeltaX;
    const float mp1Y = midY + mpDeltaY;
    const float mp2Y = midY - mpDeltaY;
    const float mp1R2 = mp1X * mp1X + mp1Y * mp1Y;
    const float mp2R2 = mp2X * mp2X + mp2Y * mp2Y;
    const float mn = (mp2R2 < mp1R2) ? mp2R2 : mp1R2;  // std::min(mp1R2, mp2R2)
    if (mn <= c.helixImpactMargin2) return false;
    return true;
}

// Pre-decision of the minimum-helix-radius cut without divisions or square roots.
// In exact arithmetic the two circle centres of doublet_stage2 are Mid +- q n (Mid the chord's
// midpoint, n its unit normal, q^2 = R^2 - c^2/4, c the chord length), so with
//   dot = x1 x2 + y1 y2,  cross = x1 y2 - x2 y1,  L = dot + (R^2 - margin^2)
// min(mp1R2, mp2R2) - margin^2 = L - Rt,  Rt = |cross| sqrt(4 R^2 - c^2) / c >= 0,
// i.e. the cut passes iff L > Rt. The reference's float chain and the polynomials below both
// carry rounding errors far below delta = 2e-5 R^2 (about 14 mm^2 of the ~7e5 mm^2 compared:
// the chain's own error is ~0.5 mm^2), so outside the band |L - Rt| <= delta the exact chain
// must give the same answer. Returns 1 = passes, 0 = fails, 2 = undecided (inside the band,
// axis-parallel or degenerate chords, non-finite values): the caller then runs doublet_stage2.
B200_HD int doublet_stage2_fast(const DevCfg& c, float x1, float y1, float x2, float y2) {
    const float dx = x2 - x1, dy = y2 - y1;
    const float c2 = dx * dx + dy * dy;
    const float fourR2 = 4.f * c.minHelixRadius2;
    const float delta = 2e-5f * c.minHelixRadius2;
    const float w = fourR2 - c2;  // 4 q^2
    const float adx = absf(dx), ady = absf(dy);
    // slope / 1/slope beyond 2^20: the reference chain saturates (s*s + 1 == s*s, NaNs at 0)
    if (!(adx > 1e-6f * ady) || !(ady > 1e-6f * adx) || !(w > 0.01f * fourR2) || !(c2 < 1e30f))
        return 2;
    const float dot = x1 * x2 + y1 * y2;
    const float cross = x1 * y2 - x2 * y1;
    const float L = dot + (c.minHelixRadius2 - c.helixImpactMargin2);
    const float rhs = cross * cross * w;  // Rt^2 c^2
    if (!(absf(L) < 1e15f) || !(rhs < 1e30f)) return 2;
    const float lo = L - delta, hi = L + delta;
    if (lo > 0.f && lo * lo * c2 > rhs) return 1;          // L - delta > Rt
    if (hi < 0.f || hi * hi * c2 < rhs) return 0;          // L + delta < Rt
    return 2;
}

// The same decision without the magnitude guards, for configurations where they cannot fire
// (DevCfg::fast_bounded: every valid spacepoint has |x|, |y| <= rb with rb < 0.99 R, rb < 1e5,
// R^2 < 1e12, so w > 0.01 * 4 R^2, c2 < 1e30, |L| < 1e15 and rhs < 1e30 hold for finite inputs;
// NaN inputs fail every comparison below and come out undecided, like in the guarded version).
B200_HD int doublet_stage2_fast_bounded(const DevCfg& c, float x1, float y1, float x2, float y2) {
    const float dx = x2 - x1, dy = y2 - y1;
    const float c2 = dx * dx + dy * dy;
    const float adx = absf(dx), ady = absf(dy);
    if (!(adx > 1e-6f * ady) || !(ady > 1e-6f * adx)) return 2;
    const float w = 4.f * c.minHelixRadius2 - c2;  // 4 q^2
    const float dot = x1 * x2 + y1 * y2;
    const float cross = x1 * y2 - x2 * y1;
    const float L = dot + (c.minHelixRadius2 - c.helixImpactMargin2);
    const float rhs = cross * cross * w;  // Rt^2 c^2
    const float delta = 2e-5f * c.minHelixRadius2;
    const float lo = L - delta, hi = L + delta;
    if (lo > 0.f && lo * lo * c2 > rhs) return 1;  // L - delta > Rt
    if (hi < 0.f || hi * hi * c2 < rhs) return 0;  // L + delta < Rt
    return 2;
}

// ---------------------------------------------------------------------------
// Fine (r, z) cells inside every reference grid bin — a pruning index that is NOT part of
// the reference: the reference tests every spacepoint of the (2*scope+1)^2 neighbour bins
// against each middle (doublet_finding.hpp:74-104). Here each bin's content is additionally
// stored sorted by (r row, z cell), and a middle only visits the cells that can hold a
// spacepoint passing the first block of doublet cuts. The windows below are conservative
// (margins far above the float rounding of the exact cuts), the exact cut still runs on
// every visited candidate, so the accepted set is identical to the full scan.
// ---------------------------------------------------------------------------
struct CellGrid {
    uint32_t NR;    // r rows per bin
    uint32_t NZc;   // z cells per bin and row
    uint32_t CPB;   // cells per bin = NR * NZc
    uint32_t NZg;   // z cells over the whole z axis = nZ * NZc
    float invRw;    // rows per mm
    float rw;       // mm per row
    float zMin;     // start of the z axis
    float invZw;    // z cells per mm
};

// Row of a radius; monotone non-decreasing in r, last row open-ended.
B200_HD uint32_t cell_row(const CellGrid& g, float r) {
    const float t = r * g.invRw;
    if (!(t >= 0.f)) return 0u;
    if (t >= static_cast<float>(g.NR)) return g.NR - 1u;
    return static_cast<uint32_t>(static_cast<int>(t));
}
// z cell inside reference z bin zb; monotone non-decreasing in z for a fixed zb, the first
// and last cells of a bin are open-ended.
B200_HD uint32_t cell_z(const CellGrid& g, uint32_t zb, float z) {
    const float t = (z - g.zMin) * g.invZw;
    long long zg;
    if (!(t >= 0.f))
        zg = 0;
    else if (t >= static_cast<float>(g.NZg))
        zg = static_cast<long long>(g.NZg) - 1;
    else
        zg = static_cast<long long>(static_cast<int>(t));
    long long l = zg - static_cast<long long>(zb) * g.NZc;
    if (l < 0) l = 0;
    if (l > static_cast<long long>(g.NZc) - 1) l = static_cast<long long>(g.NZc) - 1;
    return static_cast<uint32_t>(l);
}

// The same in two steps, for windows shared by several middles (k_doublets_tile): cell_zg is
// the cell over the whole z axis (monotone in z), cell_z_of turns it into the cell inside
// reference z bin zb; cell_z_of(g, zb, cell_zg(g, z)) == cell_z(g, zb, z).
B200_HD int cell_zg(const CellGrid& g, float z) {
    const float t = (z - g.zMin) * g.invZw;
    if (!(t >= 0.f)) return 0;
    if (t >= static_cast<float>(g.NZg)) return static_cast<int>(g.NZg) - 1;
    return static_cast<int>(t);
}
B200_HD uint32_t cell_z_of(const CellGrid& g, uint32_t zb, int zg) {
    long long l = static_cast<long long>(zg) - static_cast<long long>(zb) * g.NZc;
    if (l < 0) l = 0;
    if (l > static_cast<long long>(g.NZc) - 1) l = static_cast<long long>(g.NZc) - 1;
    return static_cast<uint32_t>(l);
}

// Serialised cell of a valid spacepoint: (reference bin, r row, z cell).
B200_HD uint32_t sp_cell(const DevCfg& c, const CellGrid& g, uint32_t bin, float r, float z) {
    const uint32_t zb = bin / c.nPhi;
    return bin * g.CPB + cell_row(g, r) * g.NZc + cell_z(g, zb, z);
}

B200_HD float fmin_nan_lo(float a, float b) { return (b < a) ? b : a; }
B200_HD float fmax_nan_hi(float a, float b) { return (b > a) ? b : a; }

// z interval [L, U] outside of which no spacepoint of row `row` can pass doublet_stage1
// against the middle (rM, zM); returns false if the whole row is excluded. NaN bounds mean
// "unbounded" to the caller (cell_z_lo / cell_z_hi).
B200_HD bool cell_row_window(const DevCfg& c, const CellGrid& g, float rM, float zM, uint32_t row,
                             float& L, float& U) {
    const float inf = u2f(0x7f800000u);
    const float er = 1e-3f + 1e-5f * (absf(rM) + static_cast<float>(row + 1u) * g.rw);
    const float rlo = static_cast<float>(row) * g.rw - er;
    const float rhi = (row + 1u >= g.NR) ? inf : static_cast<float>(row + 1u) * g.rw + er;
    const bool zo_ok = rM > 1e-3f;
    const float inv = zo_ok ? 1.f / rM : 0.f;
    const float a = (zM - c.collisionRegionMin) * inv;
    const float b = (zM - c.collisionRegionMax) * inv;
    const float cmag = (absf(zM) + absf(c.collisionRegionMin) + absf(c.collisionRegionMax)) * inv;
    bool any = false;
    L = inf;
    U = -inf;
    for (int dir = 0; dir < 2; ++dir) {
        // dir 0: bottom, deltaR = rM - r2; dir 1: top, deltaR = r2 - rM
        float dlo = (dir == 0) ? (rM - rhi) : (rlo - rM);
        float dhi = (dir == 0) ? (rM - rlo) : (rhi - rM);
        if (dlo < c.deltaRMin - er) dlo = c.deltaRMin - er;
        if (dhi > c.deltaRMax + er) dhi = c.deltaRMax + er;
        if (dlo > dhi) continue;
        if (dlo < 0.f) dlo = 0.f;
        float l = -inf, u = inf;
        if (zo_ok) {
            // bottom: z2 in (zM - a dR, zM - b dR); top: z2 in (zM + b dR, zM + a dR); the
            // bounds are linear in dR, so their extremes sit at dlo / dhi.
            const float sg = (dir == 0) ? -1.f : 1.f;
            const float v0 = zM + sg * (a * dlo), v1 = zM + sg * (a * dhi);
            const float v2 = zM + sg * (b * dlo), v3 = zM + sg * (b * dhi);
            const float m = 0.05f + 1e-4f * (absf(zM) + cmag * dhi);
            if ((v0 == v0) && (v1 == v1) && (v2 == v2) && (v3 == v3) && (m == m)) {
                l = fmin_nan_lo(fmin_nan_lo(v0, v1), fmin_nan_lo(v2, v3)) - m;
                u = fmax_nan_hi(fmax_nan_hi(v0, v1), fmax_nan_hi(v2, v3)) + m;
                if (!(l == l)) l = -inf;
                if (!(u == u)) u = inf;
            }
        }
        // |deltaZ| < cotThetaMax * deltaR and |deltaZ| < deltaZMax
        const float m2 = 0.05f + 1e-4f * absf(zM);
        float w = c.cotThetaMax * dhi;
        w = fmin_nan_lo(w, c.deltaZMax);
        w = w + m2 + 1e-4f * absf(w);
        if (w == w) {
            l = fmax_nan_hi(l, zM - w);
            u = fmin_nan_lo(u, zM + w);
        }
        if (l > u) continue;
        any = true;
        L = fmin_nan_lo(L, l);
        U = fmax_nan_hi(U, u);
    }
    return any;
}

// lin_circle (core/include/traccc/seeding/detail/lin_circle.hpp)
struct LinCircle {
    float Zo, cotTheta, iDeltaR, Er, U, V;
};

// doublet_finding_helper::transform_coordinates (:218-272). cosPhiM = xM / rM and sinPhiM = yM / rM
// depend on the middle only: callers that process many doublets of one middle pass them in.
B200_HD LinCircle transform_coordinates_cs(bool bottom, float cosPhiM, float sinPhiM, float xM, float yM,
                                           float zM, float rM, float varianceZM, float varianceRM,
                                           float x2, float y2, float z2, float varZ2, float varR2) {
    const float deltaX = x2 - xM;
    const float deltaY = y2 - yM;
    const float deltaZ = z2 - zM;
    const float x = deltaX * cosPhiM + deltaY * sinPhiM;
    const float y = deltaY * cosPhiM - deltaX * sinPhiM;
    const float iDeltaR2 = 1.f / (deltaX * deltaX + deltaY * deltaY);
    const float iDeltaR = sqrt_rn(iDeltaR2);
    float cot_theta = deltaZ * iDeltaR;
    if (bottom) cot_theta = -cot_theta;
    LinCircle l;
    l.cotTheta = cot_theta;
    l.Zo = zM - rM * cot_theta;
    l.iDeltaR = iDeltaR;
    l.U = x * iDeltaR2;
    l.V = y * iDeltaR2;
    l.Er = ((varianceZM + varZ2) + (cot_theta * cot_theta) * (varianceRM + varR2)) * iDeltaR2;
    return l;
}
B200_HD LinCircle transform_coordinates(bool bottom, float xM, float yM, float zM, float rM,
                                        float varianceZM, float varianceRM, float x2, float y2,
                                        float z2, float varZ2, float varR2) {
    return transform_coordinates_cs(bottom, xM / rM, yM / rM, xM, yM, zM, rM, varianceZM, varianceRM,
                                    x2, y2, z2, varZ2, varR2);
}

// doublet_stage1 for a candidate known to lie in a LOWER r row than the middle (so r2 < rM:
// it can only be a bottom doublet) resp. a HIGHER one (only a top doublet): the same
// comparisons as doublet_stage1 without the direction selects. true = passes.
B200_HD bool doublet_stage1_bottom(const DevCfg& c, float rM, float zM, float r2, float z2) {
    const float dR = rM - r2;
    const float dz = zM - z2;
    const float zo = zM * dR - rM * dz;
    const float acot = absf(dz);
    return !((dR >= c.deltaRMax) || (dR <= c.deltaRMin) || (acot >= c.cotThetaMax * dR) ||
             (zo <= c.collisionRegionMin * dR) || (zo >= c.collisionRegionMax * dR) ||
             (acot >= c.deltaZMax));
}
B200_HD bool doublet_stage1_top(const DevCfg& c, float rM, float zM, float r2, float z2) {
    const float dR = rM - r2;
    const float dz = zM - z2;
    const float zo = zM * dR - rM * dz;
    const float deltaR = -dR, zOrigin = -zo;  // exact negations of the <bottom> operands
    const float acot = absf(dz);
    return !((deltaR >= c.deltaRMax) || (deltaR <= c.deltaRMin) || (acot >= c.cotThetaMax * deltaR) ||
             (zOrigin <= c.collisionRegionMin * deltaR) || (zOrigin >= c.collisionRegionMax * deltaR) ||
             (acot >= c.deltaZMax));
}

// ---------------------------------------------------------------------------
// Triplet cuts (core/include/traccc/seeding/triplet_finding_helper.hpp:42-133)
// ---------------------------------------------------------------------------
// Per mid-bottom doublet constants (core/src/seeding/triplet_finding.hpp:77-82):
// note "(a*b)" then "*= (s*s)", NOT the device reference's a*b*s*s.
B200_HD void triplet_row_constants(const DevCfg& c, float cotThetaB, float& iSinTheta2,
                                   float& scatteringInRegion2) {
    iSinTheta2 = 1.f + cotThetaB * cotThetaB;
    scatteringInRegion2 = c.maxScatteringAngle2 * iSinTheta2;
    scatteringInRegion2 *= c.sigmaScattering2;
}

// First scattering cut only (:51-78): true = the pair survives it. sqrt(0) == 0
// exactly, so the IEEE square root is skipped for error-free spacepoints.
B200_HD bool triplet_cut1(float lbCot, float lbIDeltaR, float lbEr, float ltCot, float ltIDeltaR,
                          float ltEr, float varRM, float varZM, float scatteringInRegion2) {
    const float error2 =
        ltEr + lbEr + 2.f * (lbCot * ltCot * varRM + varZM) * lbIDeltaR * ltIDeltaR;
    float deltaCotTheta = lbCot - ltCot;
    const float deltaCotTheta2 = deltaCotTheta * deltaCotTheta;
    if (deltaCotTheta2 - error2 > 0) {
        deltaCotTheta = absf(deltaCotTheta);
        const float error = (error2 == 0.f) ? error2 : sqrt_rn(error2);
        const float dCotThetaMinusError2 = deltaCotTheta2 + error2 - 2.f * deltaCotTheta * error;
        if (dCotThetaMinusError2 > scatteringInRegion2) return false;
    }
    return true;
}

// Full triplet_finding_helper::isCompatible (:42-133).
B200_HD bool triplet_is_compatible(const DevCfg& c, float rM, float varRM, float varZM,
                                   const LinCircle& lb, const LinCircle& lt, float iSinTheta2,
                                   float scatteringInRegion2, float& curvature,
                                   float& impact_parameter) {
    const float error2 = lt.Er + lb.Er +
                         2.f * (lb.cotTheta * lt.cotTheta * varRM + varZM) * lb.iDeltaR *
                             lt.iDeltaR;
    float deltaCotTheta = lb.cotTheta - lt.cotTheta;
    const float deltaCotTheta2 = deltaCotTheta * deltaCotTheta;
    float error = 0.f;
    float dCotThetaMinusError2 = 0.f;
    if (deltaCotTheta2 - error2 > 0) {
        deltaCotTheta = absf(deltaCotTheta);
        error = sqrt_rn(error2);
        dCotThetaMinusError2 = deltaCotTheta2 + error2 - 2.f * deltaCotTheta * error;
        if (dCotThetaMinusError2 > scatteringInRegion2) return false;
    }
    const float dU = lt.U - lb.U;
    if (dU == 0.f) return false;
    const float A = (lt.V - lb.V) / dU;
    const float S2 = 1.f + A * A;
    const float B = lb.V - A * lb.U;
    const float B2 = B * B;
    if (S2 < B2 * c.minHelixDiameter2) return false;
    const float iHelixDiameter2 = B2 / S2;
    float pT2scatter = 4.f * iHelixDiameter2 * c.pT2perRadius;
    const float pT = c.pTPerHelixRadius * sqrt_rn(S2 / B2) / 2.f;
    if (pT > c.maxPtScattering) pT2scatter = c.pT2scatterMax;
    const float p2scatter = pT2scatter * iSinTheta2;
    if ((deltaCotTheta2 - error2 > 0.f) &&
        (dCotThetaMinusError2 > p2scatter * c.sigmaScattering * c.sigmaScattering)) {
        return false;
    }
    curvature = B / sqrt_rn(S2);
    impact_parameter = absf((A - B * rM) * rM);
    if (impact_parameter > c.impactMax) return false;
    return true;
}

// Division-free pre-filter of triplet_is_compatible: true = the (mid-bottom, mid-top) pair is
// CERTAINLY rejected by the helix-diameter cut (:95-103) or the impact-parameter cut (:127-131),
// false = undecided (run the exact chain). With dU = Ut - Ub, dV = Vt - Vb (the reference's own
// first operations, so identical floats) and num = Vb dU - dV Ub:
//   A = dV / dU,  B = num / dU,  S2 = (dU^2 + dV^2) / dU^2,  B2 = num^2 / dU^2
//   helix : S2 < B2 D                      <=>  dU^2 + dV^2 < num^2 D
//   impact: |(A - B rM) rM| > impactMax    <=>  |dV - num rM| rM > impactMax |dU|
// The reference evaluates B = Vb - A Ub with cancellation: its absolute error is bounded by
// eB = 2^-22 (|Vb| + |A Ub|); eN = eB |dU| is that bound on num. The polynomials carry a few
// 2^-24 relative roundings of their own (covered by the 1e-5 factors). A pair is only declared
// rejected when the inequality holds with those error bounds applied against it. All cuts of
// the reference are pure rejections, so testing these two first does not change the result.
// Used by k_triplets<DENSE> only: on ordinary events a 32-row block holds ~23 pairs, the exact
// evaluation runs once per block whatever it contains, and the filter costs more than it saves.
B200_HD bool triplet_certainly_rejected(const DevCfg& c, float rM, float Ub, float Vb, float Ut,
                                        float Vt) {
    const float dU = Ut - Ub;
    if (dU == 0.f) return true;  // :88-90
    const float dV = Vt - Vb;
    const float adU = absf(dU);
    const float t1 = Vb * dU, t2 = dV * Ub;
    const float num = t1 - t2;
    const float eN = 4.8e-7f * (absf(t1) + absf(t2));
    const float anum = absf(num);
    const float s2 = dU * dU + dV * dV;
    // helix: surely S2 < B2 D
    const float nlo = anum - eN;
    if (nlo > 0.f && s2 * 1.00001f < nlo * nlo * c.minHelixDiameter2) return true;
    // impact: surely |(A - B rM) rM| > impactMax
    const float nr = num * rM;
    const float g = absf(dV - nr);
    const float eG = eN * rM + 2.4e-7f * (absf(dV) + absf(nr));
    if ((g - eG) * rM > c.impactMax * adU * 1.00001f) return true;
    return false;
}

// ---------------------------------------------------------------------------
// Seed selection (core/include/traccc/seeding/seed_selecting_helper.hpp:28-80)
// ---------------------------------------------------------------------------
B200_HD float seed_weight_increase(const DevCfg& c, float rB, float rT) {
    float weight = 0.f;
    if (rB > c.good_spB_min_radius) weight = c.good_spB_weight_increase;
    if (rT < c.good_spT_max_radius) weight = c.good_spT_weight_increase;  // overwrites (:35-40)
    return weight;
}
B200_HD bool single_seed_cut(const DevCfg& c, float rB, float w) {
    return !(rB > c.good_spB_min_radius && w < c.good_spB_min_weight);
}
B200_HD bool cut_per_middle_sp(const DevCfg& c, float rB, float w) {
    return (w > c.seed_min_weight || rB > c.spB_min_radius);
}
// triplet_sorter (detail/triplet_sorter.hpp:39-70): true if seed 1 sorts before seed 2.
B200_HD bool seed_before(float w1, float s1, float w2, float s2) {
    if (w1 != w2) return w1 > w2;
    return s1 > s2;
}
// the tie-break sum of triplet_sorter (:61-64)
B200_HD float sorter_sum(float yB, float zB, float yT, float zT) {
    return yB * yB + zB * zB + yT * yT + zT * zT;
}

}  // namespace b200seed
