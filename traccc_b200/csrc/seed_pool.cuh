// seed_pool.cuh — k_triplets_pool: triplet search, compatible-seed bonus and per-middle top-N for
// the LIGHT middles, several middles per warp.
//
// k_triplets gives every active middle a warp of its own. On an ordinary event most middles are
// light — about 100 mid-bottom rows with ~0.7 mid-tops inside each row's cotTheta window — so a
// warp spends most of its instructions on per-middle and per-32-row-block fixed costs, evaluates
// ~20 (row, mid-top) pairs per block and merges ~10 triplets per flush at a third of its lanes.
// Here a warp takes POOL_G light middles at once: the rows of all members form one row space, the
// (row, mid-top) pairs of all members go through one queue and are evaluated 32 at a time at full
// lane occupancy, the accepted triplets of all members share one list that is flushed (reference
// order inside a row, compatible-seed bonus, weights, merge into the per-member top-N) when it is
// nearly full. Arithmetic and ordering are those of k_triplets (see there), so the seeds are the
// same bit for bit. Heavy middles (many combinations or long lists; see pool_is_light) stay with
// k_triplets, which has enough pairs per block to fill its lanes.
//
// Reference: core/src/seeding/triplet_finding.hpp:60-183, seed_filtering.cpp:28-123 (what is
// computed); device/common/.../impl/find_triplets.ipp:71-121 (what it replaces on the device).
#pragma once

#include "seed_kernels.cuh"

namespace b200seed {

#ifndef B200_POOL_G
#define B200_POOL_G 8
#endif
#ifndef B200_POOL_WARPS
#define B200_POOL_WARPS 8
#endif
#ifndef B200_POOL_MIN_CTAS
#define B200_POOL_MIN_CTAS 3
#endif
constexpr uint32_t POOL_G = B200_POOL_G;        // members per work item
constexpr uint32_t POOL_TC = POOL_G * POOL_NT;  // staged cotTheta values per item
constexpr uint32_t POOL_PQ = 256;               // queued (row, mid-top) pairs
constexpr uint32_t POOL_LC = 128;               // accepted triplets waiting for their flush
constexpr int POOL_WARPS = B200_POOL_WARPS;
static_assert(POOL_G <= 8 && POOL_LC >= 2 * POOL_NT, "packing of (member, row, top); one row must fit the list");

// Shared memory of one warp: fixed part (K-dependent top-N arrays follow, see pool_smem_per_warp).
struct __align__(16) PoolWarp {
    float4 M[POOL_G];                  // members {x,y,z,r}
    float2 VM[POOL_G];                 // {varZ, varR}
    float cot[POOL_TC];                // cotTheta of the members' mid-tops, concatenated (each sorted)
    BlockTriplet list[POOL_LC];        // accepted triplets (complete rows are flushed)
    uint32_t lpos[POOL_LC];            // grid position of the top spacepoint
    uint32_t lrow[POOL_LC];            // (member << 16) | mid-bottom row
    uint32_t pq[POOL_PQ];              // pairs: member << 28 | row << 14 | mid-top
    uint16_t ord[POOL_LC];
    uint8_t aux[POOL_LC];
    uint32_t mpos[POOL_G], nb[POOL_G], nt[POOL_G], offB[POOL_G], offT[POOL_G];
    uint32_t rowbase[POOL_G + 1], topbase[POOL_G + 1], walk_r0[POOL_G], ntop[POOL_G];
    uint32_t segb[POOL_G], sege[POOL_G];   // list segment of every member during a flush
    float maxEr[POOL_G], minEr[POOL_G], maxIDR[POOL_G];
};
__host__ __device__ inline size_t pool_smem_per_warp(uint32_t K) {
    return sizeof(PoolWarp) + size_t(POOL_G) * K * 5 * 4;
}

__global__ void __launch_bounds__(POOL_WARPS * 32, B200_POOL_MIN_CTAS)
k_triplets_pool(const DevCfg cfg, const TripletArgs a) {
    extern __shared__ __align__(16) unsigned char s_pool_raw[];
    __shared__ uint32_t s_ntrip;
    __shared__ unsigned long long s_tests;
    __shared__ uint32_t s_pre[WORK_CLASSES + 1];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltmask = lanemask_lt();
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t K = cfg.maxSeedsPerSpM;
    unsigned char* base = s_pool_raw + pool_smem_per_warp(K) * warp;
    PoolWarp& W = *reinterpret_cast<PoolWarp*>(base);
    float* top_w = reinterpret_cast<float*>(base + sizeof(PoolWarp));   // [POOL_G][K]
    float* top_s = top_w + POOL_G * K;
    float* top_rb = top_s + POOL_G * K;
    uint32_t* top_b = reinterpret_cast<uint32_t*>(top_rb + POOL_G * K);
    uint32_t* top_t = top_b + POOL_G * K;
    if (threadIdx.x == 0) {
        s_ntrip = 0;
        s_tests = 0ull;
    }
    work_prefix(a.ctrl->n_cls, s_pre);
    __syncthreads();
    const uint32_t n_valid = a.ctrl->n_valid;
    const bool has_var = a.ctrl->has_variance != 0u;
    uint32_t acc_trip = 0;
    unsigned long long acc_tests = 0ull, acc_visited = 0ull;
    const uint32_t n_heavy = s_pre[WORK_HEAVY_CLASSES], n_light = s_pre[WORK_CLASSES] - n_heavy;

    // order of a spacepoint among the doublet partners of member j in the reference (canon_key
    // of the doublet kernels); only evaluated for full ties in the seed ranking
    auto tie_key = [&](uint32_t j, uint32_t pos) -> unsigned long long {
        const uint32_t pb = __ldg(a.sorted_bin + pos) % cfg.nPhi;
        const uint32_t w = (pb + cfg.nPhi - W.walk_r0[j]) % cfg.nPhi;
        return (unsigned long long)w * n_valid + pos;
    };
    auto before = [&](uint32_t j, float w1, float s1, uint32_t b1, uint32_t t1, float w2, float s2,
                      uint32_t b2, uint32_t t2) -> bool {
        if (w1 != w2 || s1 != s2) return seed_before(w1, s1, w2, s2);
        const unsigned long long k1 = tie_key(j, b1), k2 = tie_key(j, b2);
        return (k1 != k2) ? (k1 < k2) : (tie_key(j, t1) < tie_key(j, t2));
    };

    while (true) {
        uint32_t t0 = 0;
        if (lane == 0) t0 = atomicAdd(&a.ctrl->ticket_p, POOL_G);
        t0 = __shfl_sync(FULL, t0, 0);
        if (t0 >= n_light) break;
        const uint32_t G = (n_light - t0 < POOL_G) ? (n_light - t0) : POOL_G;
        // ---- members ----
        uint32_t my_nb = 0, my_nt = 0;
        if (lane < G) {
            const uint32_t m = work_item(a.active_list, a.n_sp, s_pre, n_heavy + t0 + lane);  // light classes
            my_nb = a.cnt_b[m];
            my_nt = a.cnt_t[m];
            W.mpos[lane] = m;
            W.nb[lane] = my_nb;
            W.nt[lane] = my_nt;
            W.offB[lane] = a.off_b[m];
            W.offT[lane] = a.off_t[m];
            W.M[lane] = __ldg(a.sp4 + m);
            W.VM[lane] = __ldg(a.var2 + m);
            W.walk_r0[lane] = circular_remap(cfg.nPhi, __ldg(a.sorted_bin + m) % cfg.nPhi, -int(cfg.scope0));
            W.ntop[lane] = 0u;
            W.maxEr[lane] = 0.f;
            W.minEr[lane] = 0.f;
            W.maxIDR[lane] = 0.f;
            acc_tests += (unsigned long long)my_nb * my_nt;
        }
        const uint32_t inclR = warp_incl_scan(my_nb, lane), inclT = warp_incl_scan(my_nt, lane);
        const uint32_t R = __shfl_sync(FULL, inclR, 31), T = __shfl_sync(FULL, inclT, 31);
        if (lane < G) {
            W.rowbase[lane] = inclR - my_nb;
            W.topbase[lane] = inclT - my_nt;
        }
        if (lane == 0) {
            W.rowbase[G] = R;
            W.topbase[G] = T;
        }
        __syncwarp();
        // member that owns pooled index i of a prefix array p[0..G] (G <= 8)
        auto owner = [&](const uint32_t* p, uint32_t i) -> uint32_t {
            uint32_t j = 0;
#pragma unroll
            for (uint32_t step = 4; step >= 1; step >>= 1)
                if (j + step < G && p[j + step] <= i) j += step;
            return j;
        };
        // ---- cotTheta of all mid-tops into shared memory (T <= POOL_TC by construction) ----
        for (uint32_t i = lane; i < T; i += 32) {
            const uint32_t j = owner(W.topbase, i);
            const float4 ta = __ldg(&a.arena_t[W.offT[j] + (i - W.topbase[j])].a);
            W.cot[i] = ta.x;
            if (has_var) {  // bounds over the member's mid-tops for the conservative window
                atomicMax(reinterpret_cast<int*>(&W.maxEr[j]), __float_as_int(fmaxf(ta.z, 0.f)));
                if (ta.z < 0.f) W.minEr[j] = ta.z;  // any negative value switches the pruning off
                atomicMax(reinterpret_cast<int*>(&W.maxIDR[j]), __float_as_int(fmaxf(ta.y, 0.f)));
            }
        }
        __syncwarp();

        uint32_t nlist = 0, npq = 0;
        // ---- flush: entries [0, nc) of the list (complete rows) ----
        auto flush = [&](const uint32_t nc) {
            __syncwarp();
            // (0) list segment of every member (rows are member-major)
            if (lane < G) W.segb[lane] = W.sege[lane] = 0u;
            __syncwarp();
            for (uint32_t i = lane; i < nc; i += 32) {
                const uint32_t j = W.lrow[i] >> 16;
                if (i == 0 || (W.lrow[i - 1] >> 16) != j) W.segb[j] = i;
                if (i + 1 == nc || (W.lrow[i + 1] >> 16) != j) W.sege[j] = i + 1;
            }
            // (1) reference order inside every row: ord[first of row + rank by canon_key]
            for (uint32_t i = lane; i < nc; i += 32) {
                const uint32_t row = W.lrow[i], key = W.list[i].key;
                uint32_t sgm = i, rank = 0;
                while (sgm > 0 && W.lrow[sgm - 1] == row) {
                    --sgm;
                    rank += (W.list[sgm].key < key) ? 1u : 0u;
                }
                for (uint32_t q = i + 1; q < nc && W.lrow[q] == row; ++q)
                    rank += (W.list[q].key < key) ? 1u : 0u;
                W.ord[sgm + rank] = uint16_t(i);
            }
            __syncwarp();
            // (2) compatible-seed bonus (triplet_finding.hpp:107-179), lane per triplet
            for (uint32_t i = lane; i < nc; i += 32) {
                const BlockTriplet cur = W.list[i];
                const uint32_t row = W.lrow[i];
                uint32_t sgm = i, egm = i + 1;
                while (sgm > 0 && W.lrow[sgm - 1] == row) --sgm;
                while (egm < nc && W.lrow[egm] == row) ++egm;
                const float lower = cur.curvature - cfg.deltaInvHelixDiameter;
                const float upper = cur.curvature + cfg.deltaInvHelixDiameter;
                float compat[MAX_COMPAT];
                uint32_t ncompat = 0;
                for (uint32_t q = sgm; q < egm; ++q) {
                    const uint32_t o_i = W.ord[q];
                    if (o_i == i) continue;
                    const BlockTriplet o = W.list[o_i];
                    const float deltaR = cur.rT - o.rT;
                    if (absf(deltaR) < cfg.filterDeltaRMin) continue;
                    if (o.curvature < lower) continue;
                    if (o.curvature > upper) continue;
                    bool newCompSeed = true;
#pragma unroll
                    for (uint32_t c = 0; c < MAX_COMPAT; ++c)
                        if (c < ncompat && absf(compat[c] - o.rT) < cfg.filterDeltaRMin) newCompSeed = false;
                    if (newCompSeed) {
#pragma unroll
                        for (uint32_t c = 0; c < MAX_COMPAT; ++c)
                            if (c == ncompat) compat[c] = o.rT;
                        ++ncompat;
                    }
                    if (ncompat >= cfg.compatSeedLimit) break;
                }
                W.aux[i] = uint8_t(ncompat);
            }
            __syncwarp();
            // (3) final weight, single-seed cut, sorter sum; optional dump
            for (uint32_t i = lane; i < nc; i += 32) {
                BlockTriplet cur = W.list[i];
                const uint32_t j = W.lrow[i] >> 16, rl = W.lrow[i] & 0xFFFFu;
                float w = cur.weight;  // the reference adds compatSeedWeight one at a time (:171)
                for (uint32_t q = W.aux[i]; q > 0; --q) w += cfg.compatSeedWeight;
                const float4 bb = __ldg(&a.arena_b[W.offB[j] + rl].b);
                const uint32_t pos_b = __float_as_uint(bb.w), pos_t = W.lpos[i];
                if (a.dump) {
                    const uint32_t d = atomicAdd(&a.ctrl->dump_cursor, 1u);
                    if (d < a.max_dump) {
                        TripletDumpRec r;
                        r.pos_b = pos_b, r.pos_m = W.mpos[j], r.pos_t = pos_t, r.mb_idx = rl;
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
                cur.key = keep ? pos_b : 0xFFFFFFFFu;
                W.list[i] = cur;
            }
            __syncwarp();
            // (4) merge into the per-member top-K: rank of every kept triplet in the union of its
            //     member's kept triplets and current top-K (triplet_sorter's order; full ties: the
            //     reference's order of discovery)
            for (uint32_t i = lane; i < nc; i += 32) {
                const BlockTriplet c = W.list[i];
                const uint32_t ct = W.lpos[i];
                const uint32_t j = W.lrow[i] >> 16;
                const uint32_t ntop = W.ntop[j];
                const float* tw = top_w + j * K;
                const float* ts = top_s + j * K;
                const uint32_t* tb = top_b + j * K;
                const uint32_t* tt = top_t + j * K;
                uint32_t rank = 0xFFu;
                bool in = (c.key != 0xFFFFFFFFu);
                if (in && ntop == K)  // cannot displace anything: skip the ranking
                    in = before(j, c.weight, c.rT, c.key, ct, tw[K - 1], ts[K - 1], tb[K - 1], tt[K - 1]);
                if (in) {
                    rank = 0;
                    for (uint32_t q = 0; q < ntop; ++q)
                        rank += before(j, tw[q], ts[q], tb[q], tt[q], c.weight, c.rT, c.key, ct) ? 1u : 0u;
                    const uint32_t sb = W.segb[j], se = W.sege[j];
                    for (uint32_t q = sb; q < se && rank < K; ++q) {
                        const BlockTriplet o = W.list[q];
                        if (q != i && o.key != 0xFFFFFFFFu &&
                            before(j, o.weight, o.rT, o.key, W.lpos[q], c.weight, c.rT, c.key, ct))
                            ++rank;
                    }
                    if (rank >= K) rank = 0xFFu;
                }
                W.aux[i] = uint8_t(rank);
            }
            __syncwarp();
            // current entries move down by the number of new entries sorted before them: lanes =
            // (member, slot), 32 / K members per pass; a member's slots are all read before any is
            // rewritten
            {
                const float invK = 1.f / float(K);
                const uint32_t mpp = div_small(32u, invK);  // members per pass (K <= 16)
                const uint32_t jl = div_small(lane, invK), q = lane - jl * K;
                for (uint32_t j0 = 0; j0 < G; j0 += mpp) {
                    const uint32_t j = j0 + jl;
                    float ew = 0.f, es = 0.f, erb = 0.f;
                    uint32_t eb = 0, et = 0, dst = 0xFFFFFFFFu;
                    if (jl < mpp && j < G && q < W.ntop[j]) {
                        const uint32_t o = j * K + q;
                        ew = top_w[o], es = top_s[o], erb = top_rb[o];
                        eb = top_b[o], et = top_t[o];
                        uint32_t npos = q;
                        for (uint32_t x = W.segb[j]; x < W.sege[j]; ++x) {
                            if (W.aux[x] == 0xFFu) continue;
                            const BlockTriplet o2 = W.list[x];
                            npos += before(j, o2.weight, o2.rT, o2.key, W.lpos[x], ew, es, eb, et) ? 1u : 0u;
                        }
                        if (npos < K) dst = j * K + npos;
                    }
                    __syncwarp();
                    if (dst != 0xFFFFFFFFu) {
                        top_w[dst] = ew, top_s[dst] = es, top_rb[dst] = erb;
                        top_b[dst] = eb, top_t[dst] = et;
                    }
                    __syncwarp();
                }
            }
            for (uint32_t i = lane; i < nc; i += 32) {
                if (W.aux[i] == 0xFFu) continue;
                const BlockTriplet c = W.list[i];
                const uint32_t o = (W.lrow[i] >> 16) * K + W.aux[i];
                top_w[o] = c.weight, top_s[o] = c.rT, top_rb[o] = c.curvature;
                top_b[o] = c.key, top_t[o] = W.lpos[i];
            }
            __syncwarp();
            if (lane < G) {  // new fill level of every member's top-K
                uint32_t add = 0;
                for (uint32_t x = W.segb[lane]; x < W.sege[lane]; ++x) add += (W.aux[x] != 0xFFu) ? 1u : 0u;
                const uint32_t nt2 = W.ntop[lane] + add;
                W.ntop[lane] = nt2 < K ? nt2 : K;
            }
            acc_trip += nc;
            // the entries of an unfinished row move to the front
            const uint32_t rest = nlist - nc;
            BlockTriplet mv;
            uint32_t mp = 0, mr = 0;
            for (uint32_t i0 = 0; i0 < rest; i0 += 32) {  // rest <= POOL_NT: ascending copy in chunks is safe
                const uint32_t i = i0 + lane;
                if (i < rest) mv = W.list[nc + i], mp = W.lpos[nc + i], mr = W.lrow[nc + i];
                __syncwarp();
                if (i < rest) W.list[i] = mv, W.lpos[i] = mp, W.lrow[i] = mr;
                __syncwarp();
            }
            nlist = rest;
            __syncwarp();
        };
        // ---- drain: evaluate the queued pairs 32 at a time with the exact reference cuts; `last`:
        //      no more pairs will come for this item, flush everything at the end ----
        auto drain = [&](const bool last) {
            for (uint32_t q0 = 0;; q0 += 32) {
                const bool fin = q0 >= npq;
                if ((fin && last && nlist != 0u) || (!fin && nlist + 32u > POOL_LC)) {
                    // the complete rows: everything before the row of the last entry (more of its
                    // pairs may still be queued), or everything when the item is finished
                    uint32_t nc = nlist;
                    if (!fin) {
                        const uint32_t lastrow = W.lrow[nlist - 1];
                        while (nc > 0 && W.lrow[nc - 1] == lastrow) --nc;
                        if (nc == 0) {  // one row fills the list: cannot happen for POOL_LC >= 2 POOL_NT
                            if (lane == 0) atomicOr(&a.ctrl->overflow, B200SEED_OVF_TRIPLETS);
                            nc = nlist;
                        }
                    }
                    flush(nc);
                }
                if (fin) break;
                const uint32_t qi = q0 + lane;
                bool ok = false;
                uint32_t key = 0, pos_t = 0, lr = 0;
                float curvature = 0.f, impact = 0.f, rT = 0.f;
                if (qi < npq) {
                    const uint32_t e = W.pq[qi];
                    const uint32_t j = e >> 28, rl = (e >> 14) & 0x3FFFu, tt = e & 0x3FFFu;
                    const DoubletRec* LB = a.arena_b + W.offB[j];
                    const DoubletRec* LT = a.arena_t + W.offT[j];
                    const float4 ba = __ldg(&LB[rl].a);
                    const float4 bb = __ldg(&LB[rl].b);
                    const float4 ta = __ldg(&LT[tt].a);
                    const float4 tb = __ldg(&LT[tt].b);
                    const float4 Mj = W.M[j];
                    const float2 Vj = W.VM[j];
                    LinCircle lb, lt;
                    lb.cotTheta = ba.x, lb.iDeltaR = ba.y, lb.Er = ba.z, lb.U = ba.w;
                    lb.V = bb.x, lb.Zo = bb.y;
                    lt.cotTheta = ta.x, lt.iDeltaR = ta.y, lt.Er = ta.z, lt.U = ta.w;
                    lt.V = tb.x, lt.Zo = 0.f;
                    float is2, s2;
                    triplet_row_constants(cfg, lb.cotTheta, is2, s2);
                    ok = triplet_is_compatible(cfg, Mj.w, Vj.y, Vj.x, lb, lt, is2, s2, curvature, impact);
                    key = __float_as_uint(tb.y);
                    rT = tb.z;
                    pos_t = __float_as_uint(tb.w);
                    lr = (j << 16) | rl;
                }
                const uint32_t mk = __ballot_sync(FULL, ok);
                if (ok) {
                    const uint32_t k = nlist + __popc(mk & ltmask);
                    BlockTriplet en;
                    en.key = key;
                    en.curvature = curvature;
                    en.weight = -impact * cfg.impactWeightFactor;
                    en.rT = rT;
                    W.list[k] = en;
                    W.lpos[k] = pos_t;
                    W.lrow[k] = lr;
                }
                nlist += __popc(mk);
                __syncwarp();
            }
            npq = 0;
        };

        // ---- rows of all members, 32 at a time: cotTheta windows -> pair queue; one more round
        //      after the last block drains what is left ----
        for (uint32_t r0 = 0;; r0 += 32) {
            const bool fin = r0 >= R;
            const uint32_t r = r0 + lane;
            uint32_t lo = 0, hi = 0, j = 0, rl = 0;
            if (r < R) {
                j = owner(W.rowbase, r);
                rl = r - W.rowbase[j];
                const uint32_t nt = W.nt[j];
                const float* cot = W.cot + W.topbase[j];
                const float4 la = __ldg(&a.arena_b[W.offB[j] + rl].a);
                const float2 Vj = W.VM[j];
                const float varZM = Vj.x, varRM = Vj.y;
                float iSinTheta2, sir2;
                triplet_row_constants(cfg, la.x, iSinTheta2, sir2);
                // window half-width: see k_triplets (same margins); the list is sorted, so its
                // largest |cotTheta| sits at one of its ends
                const float maxAbsCot = fmaxf(absf(cot[0]), absf(cot[nt - 1u]));
                const float maxEr = W.maxEr[j], maxIDR = W.maxIDR[j];
                const bool sane = (varRM >= 0.f) && (varZM >= 0.f) && (W.minEr[j] >= 0.f) && (maxEr < 1e30f) &&
                                  (maxIDR < 1e30f) && (maxAbsCot < 1e30f);
                const float e2max = la.z + maxEr + 2.f * (absf(la.x) * maxAbsCot * varRM + varZM) * la.y * maxIDR;
                const float Wd = 1.004f * sqrt_rn(e2max) + 1.002f * sqrt_rn(sir2) +
                                 4e-6f * (absf(la.x) + maxAbsCot) + 1e-30f;
                const bool prune = sane && (la.z >= 0.f) && (Wd < 1e30f) && (sir2 >= 0.f);
                lo = prune ? smem_lower_bound(cot, nt, la.x - Wd) : 0u;
                hi = prune ? smem_upper_bound(cot, nt, la.x + Wd) : nt;
                if (hi < lo) hi = lo;
            }
            const uint32_t wdt = hi - lo;
            const uint32_t incl = warp_incl_scan(wdt, lane);
            const uint32_t excl = incl - wdt;
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            acc_visited += total;
            // emit the pairs of this block in slices that fit the queue (row-major order)
            uint32_t e0 = 0;
            do {
                const uint32_t room = POOL_PQ - npq;
                const uint32_t e1 = (total - e0 < room) ? total : (e0 + room);
                // my pairs with flattened index in [e0, e1)
                const uint32_t a0 = (excl > e0) ? excl : e0;
                const uint32_t a1 = (incl < e1) ? incl : e1;
                for (uint32_t x = a0; x < a1; ++x)
                    W.pq[npq + (x - e0)] = (j << 28) | (rl << 14) | (lo + (x - excl));
                npq += e1 - e0;
                e0 = e1;
                __syncwarp();
                if (npq == POOL_PQ || fin) drain(fin);
            } while (e0 < total);
            if (fin) break;
        }
        __syncwarp();
        // ---- final per-member selection (seed_filtering.cpp:84-122) ----
        if (lane < G) {
            const uint32_t m = W.mpos[lane];
            const uint32_t ntop = W.ntop[lane];
            uint32_t nout = 0;
            for (uint32_t i = 0; i < ntop; ++i) {
                const uint32_t o = lane * K + i;
                if (i == 0 || cut_per_middle_sp(cfg, top_rb[o], top_w[o])) {
                    a.seed_b[size_t(m) * K + nout] = top_b[o];
                    a.seed_t[size_t(m) * K + nout] = top_t[o];
                    a.seed_w[size_t(m) * K + nout] = top_w[o];
                    ++nout;
                }
            }
            a.seed_cnt[m] = nout;
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc_tests += __shfl_xor_sync(FULL, acc_tests, o);
    if (lane == 0) {
        atomicAdd(&s_ntrip, acc_trip);
        atomicAdd(&s_tests, acc_tests);
        if (acc_visited) atomicAdd(&a.ctrl->triplet_visited, acc_visited);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_ntrip) atomicAdd(&a.ctrl->n_triplets, s_ntrip);
        if (s_tests) atomicAdd(&a.ctrl->triplet_tests, s_tests);
    }
}

}  // namespace b200seed
