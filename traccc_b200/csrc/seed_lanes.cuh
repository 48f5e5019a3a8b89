// b200seed — triplet search for the LIGHT middles, one middle per lane. EXPERIMENTAL
// (B200SEED_TRIPLETS=lanes; bit-identical to k_triplets, test_parity_lane_triplet_kernel): the
// pattern that pays in k_doublets<3> — 32 short latency chains per warp instead of one — does not
// pay here: a lane's program over up to 31 mid-bottom rows is ~450 us of dependent loads, and an
// event has only ~300 such warps, so the launch is one long tail (triplet stage 110 -> 546 us on the
// 10k-particle event; with the 32+-row light class included: 948 us).
//
// k_triplets gives every middle a warp. For the light middles (few mid-bottom rows, at most 64
// mid-tops, fewer than 32 rows: work classes LANES_FIRST_CLASS .. WORK_CLASSES-1) that is one long chain of dependent
// loads and warp-wide phases over a handful of elements — 8.5 us even for a middle with six
// bottoms and four tops (profiles/r02_phases_and_tails.md). Here a warp draws 32 of them and
// every LANE runs the reference's loop nest for its own middle, serially
// (triplet_finding.hpp:60-183, seed_filtering.cpp:28-123): 32 chains in flight per warp instead of
// one. Same arithmetic, same windows (the cotTheta window of every mid-bottom row, searched in the
// lane's column of a shared-memory copy of its mid-tops' cotTheta), same orders:
//   * the triplets of one mid-bottom row are kept in the reference's order (canon_key of the top)
//     for the compatible-seed bonus,
//   * the per-middle top-N is kept sorted under triplet_sorter's order with the reference's order
//     of discovery for full ties (a strict total order: the result does not depend on the order
//     in which candidates arrive).
// A row with more accepted triplets than the lane's buffer holds hands the middle to the slow path
// (slow_list, as k_triplets does). Launched behind k_triplets<.> (heavy_only) as a programmatic
// dependent: no data is shared, so its CTAs fill the slots that launch frees at its end; it
// executes griddepcontrol.wait before it completes (see k_doublets<3>).
#pragma once

namespace b200seed {

constexpr int LANES_WARPS = 4;          // warps per CTA
constexpr uint32_t LANES_ROW_CAP = 12;  // accepted triplets of one mid-bottom row kept per lane

__host__ __device__ inline size_t lanes_smem_bytes() {
    return size_t(LANES_WARPS) * POOL_NT * 32 * sizeof(float);  // cotTheta columns
}

struct LaneTop {
    float w, s, rb;
    uint32_t b, t;
};

__global__ void __launch_bounds__(LANES_WARPS * 32)
k_triplets_lanes(const __grid_constant__ DevCfg cfg, const __grid_constant__ TripletArgs a) {
    extern __shared__ __align__(16) float s_cot_all[];
    __shared__ uint32_t s_pre[WORK_CLASSES + 1];
    __shared__ uint32_t s_ntrip;
    __shared__ unsigned long long s_tests, s_visited;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* cot_s = s_cot_all + size_t(warp) * POOL_NT * 32 + lane;  // element t at cot_s[t * 32]
    if (threadIdx.x == 0) {
        s_ntrip = 0;
        s_tests = s_visited = 0ull;
    }
    work_prefix(a.ctrl->n_cls, s_pre);
    __syncthreads();
    const uint32_t n_valid = a.ctrl->n_valid;
    const uint32_t K = cfg.maxSeedsPerSpM;
    const uint32_t first = s_pre[LANES_FIRST_CLASS], n_items = s_pre[WORK_CLASSES] - first;
    uint32_t acc_trip = 0;
    unsigned long long acc_tests = 0ull, acc_visited = 0ull;

    while (true) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&a.ctrl->ticket_p, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_items) break;
        if (base + lane >= n_items) continue;  // (no warp-wide operation below)
        const uint32_t m = work_item(a.active_list, a.n_sp, s_pre, first + base + lane);
        const uint32_t nb = a.cnt_b[m], nt = a.cnt_t[m];  // nt <= POOL_NT (pool_is_light)
        acc_tests += (unsigned long long)nb * nt;
        const DoubletRec* LB = a.arena_b + a.off_b[m];
        const DoubletRec* LT = a.arena_t + a.off_t[m];
        const float4 M = __ldg(a.sp4 + m);
        const float2 VM = __ldg(a.var2 + m);  // {varZ, varR}
        const float rM = M.w, varZM = VM.x, varRM = VM.y;
        const uint32_t walk_r0 =
            circular_remap(cfg.nPhi, __ldg(a.sorted_bin + m) % cfg.nPhi, -int(cfg.scope0));
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

        // cotTheta column of the mid-tops (they are sorted by it) + the bounds of the window width
        float maxEr = 0.f, minEr = 0.f, maxIDR = 0.f, maxAbsCot = 0.f;
        for (uint32_t t = 0; t < nt; ++t) {
            const float4 ta = __ldg(&LT[t].a);
            cot_s[t * 32] = ta.x;
            maxEr = fmaxf(maxEr, ta.z);
            minEr = fminf(minEr, ta.z);
            maxIDR = fmaxf(maxIDR, ta.y);
            maxAbsCot = fmaxf(maxAbsCot, absf(ta.x));
        }
        const bool sane = (varRM >= 0.f) && (varZM >= 0.f) && (minEr >= 0.f) && (maxEr < 1e30f) &&
                          (maxIDR < 1e30f) && (maxAbsCot < 1e30f);
        // number of mid-tops with cotTheta < v (strict == true) resp. <= v
        auto bound = [&](float v, bool strict) -> uint32_t {
            uint32_t b0 = 0, n = nt;
            while (n > 1u) {
                const uint32_t half = n >> 1;
                const float c = cot_s[(b0 + half - 1u) * 32];
                b0 += (strict ? (c < v) : (c <= v)) ? half : 0u;
                n -= half;
            }
            const float c = cot_s[b0 * 32];
            return b0 + ((strict ? (c < v) : (c <= v)) ? 1u : 0u);
        };

        LaneTop top[MAX_TOPK];
        uint32_t ntop = 0;
        bool handed_over = false;
        for (uint32_t row = 0; row < nb && !handed_over; ++row) {
            const float4 ba = __ldg(&LB[row].a);
            const float4 bb = __ldg(&LB[row].b);
            LinCircle lb;
            lb.cotTheta = ba.x, lb.iDeltaR = ba.y, lb.Er = ba.z, lb.U = ba.w;
            lb.V = bb.x, lb.Zo = bb.y;
            float is2, sir2;
            triplet_row_constants(cfg, ba.x, is2, sir2);
            // the window of k_triplets (conservative; the exact cut still runs on what is inside)
            const float e2max = ba.z + maxEr + 2.f * (absf(ba.x) * maxAbsCot * varRM + varZM) * ba.y * maxIDR;
            const float W = 1.004f * sqrt_rn(e2max) + 1.002f * sqrt_rn(sir2) +
                            4e-6f * (absf(ba.x) + maxAbsCot) + 1e-30f;
            const bool prune = sane && (ba.z >= 0.f) && (W < 1e30f) && (sir2 >= 0.f);
            uint32_t lo = prune ? bound(ba.x - W, true) : 0u;
            uint32_t hi = prune ? bound(ba.x + W, false) : nt;
            if (hi < lo) hi = lo;
            acc_visited += hi - lo;

            // accepted triplets of this row, in the reference's order (canon_key of the mid-top)
            uint32_t rk[LANES_ROW_CAP], rp[LANES_ROW_CAP];
            float rc[LANES_ROW_CAP], rw[LANES_ROW_CAP], rr[LANES_ROW_CAP];
            uint32_t n = 0;
            for (uint32_t tt = lo; tt < hi; ++tt) {
                const float4 ta = __ldg(&LT[tt].a);
                const float4 tb = __ldg(&LT[tt].b);
                LinCircle lt;
                lt.cotTheta = ta.x, lt.iDeltaR = ta.y, lt.Er = ta.z, lt.U = ta.w;
                lt.V = tb.x, lt.Zo = 0.f;
                float curvature = 0.f, impact = 0.f;
                if (!triplet_is_compatible(cfg, rM, varRM, varZM, lb, lt, is2, sir2, curvature, impact)) continue;
                if (n == LANES_ROW_CAP) {
                    handed_over = true;
                    break;
                }
                const uint32_t key = __float_as_uint(tb.y);
                uint32_t p = n;
                while (p > 0 && rk[p - 1] > key) {
                    rk[p] = rk[p - 1], rp[p] = rp[p - 1], rc[p] = rc[p - 1], rw[p] = rw[p - 1], rr[p] = rr[p - 1];
                    --p;
                }
                rk[p] = key, rp[p] = __float_as_uint(tb.w), rc[p] = curvature;
                rw[p] = -impact * cfg.impactWeightFactor, rr[p] = tb.z;
                ++n;
            }
            if (handed_over) {
                // more accepted triplets in this row than the buffer holds: the slow path redoes the
                // middle; rows before this one were counted and dumped here
                const uint32_t e = atomicAdd(&a.ctrl->n_slow, 1u);
                a.slow_list[2 * e] = m;
                a.slow_list[2 * e + 1] = row;
                break;
            }
            if (n == 0) continue;
            acc_trip += n;
            const uint32_t pos_b = __float_as_uint(bb.w);
            const float rB = bb.z;
            const float4 PB = __ldg(a.sp4 + pos_b);
            for (uint32_t i = 0; i < n; ++i) {
                // compatible-seed bonus (triplet_finding.hpp:107-179): the other triplets of this
                // mid-bottom doublet, in the reference's order
                const float lower = rc[i] - cfg.deltaInvHelixDiameter;
                const float upper = rc[i] + cfg.deltaInvHelixDiameter;
                float compat[MAX_COMPAT];
                uint32_t ncompat = 0;
                for (uint32_t q = 0; q < n; ++q) {
                    if (q == i) continue;
                    const float deltaR = rr[i] - rr[q];
                    if (absf(deltaR) < cfg.filterDeltaRMin) continue;
                    if (rc[q] < lower) continue;
                    if (rc[q] > upper) continue;
                    bool newCompSeed = true;
                    for (uint32_t c = 0; c < ncompat; ++c)
                        if (absf(compat[c] - rr[q]) < cfg.filterDeltaRMin) newCompSeed = false;
                    if (newCompSeed) {
                        if (ncompat < MAX_COMPAT) compat[ncompat] = rr[q];
                        ++ncompat;
                    }
                    if (ncompat >= cfg.compatSeedLimit) break;
                }
                float w = rw[i];  // the reference adds compatSeedWeight one at a time (:171)
                for (uint32_t q = ncompat; q > 0; --q) w += cfg.compatSeedWeight;
                if (a.dump) {
                    const uint32_t d = atomicAdd(&a.ctrl->dump_cursor, 1u);
                    if (d < a.max_dump) {
                        TripletDumpRec r;
                        r.pos_b = pos_b, r.pos_m = m, r.pos_t = rp[i], r.mb_idx = row;
                        r.mt_idx = rk[i], r.curvature = rc[i], r.weight = w;
                        r.z_vertex = bb.y;
                        a.dump[d] = r;
                    } else {
                        atomicOr(&a.ctrl->overflow, B200SEED_OVF_DUMP);
                    }
                }
                w += seed_weight_increase(cfg, rB, rr[i]);
                if (!single_seed_cut(cfg, rB, w)) continue;
                const float4 PT = __ldg(a.sp4 + rp[i]);
                const float s = sorter_sum(PB.y, PB.z, PT.y, PT.z);
                // position in the per-middle top-N
                uint32_t p = 0;
                while (p < ntop && before(top[p].w, top[p].s, top[p].b, top[p].t, w, s, pos_b, rp[i])) ++p;
                if (p >= K) continue;
                const uint32_t last = (ntop < K) ? ntop : K - 1u;
                for (uint32_t q = last; q > p; --q) top[q] = top[q - 1];
                top[p] = LaneTop{w, s, rB, pos_b, rp[i]};
                if (ntop < K) ++ntop;
            }
        }
        if (handed_over) continue;
        // final per-middle selection (seed_filtering.cpp:84-122)
        uint32_t o = 0;
        for (uint32_t i = 0; i < ntop; ++i) {
            if (i != 0 && !cut_per_middle_sp(cfg, top[i].rb, top[i].w)) continue;
            a.seed_b[size_t(m) * K + o] = top[i].b;
            a.seed_t[size_t(m) * K + o] = top[i].t;
            a.seed_w[size_t(m) * K + o] = top[i].w;
            ++o;
        }
        a.seed_cnt[m] = o;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc_trip += __shfl_xor_sync(0xffffffffu, acc_trip, o);
        acc_tests += __shfl_xor_sync(0xffffffffu, acc_tests, o);
        acc_visited += __shfl_xor_sync(0xffffffffu, acc_visited, o);
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
    // launched as a programmatic dependent of the warp-per-middle launch: not to complete before it
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace b200seed
