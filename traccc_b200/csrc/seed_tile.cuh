// seed_tile.cuh — k_doublets_tile: the doublet search over GROUPS of neighbouring middles.
//
// Replaces device/common/.../impl/count_doublets.ipp:53-137 + find_doublets.ipp:22-139 (thread per
// middle, every spacepoint of the neighbour bins loaded through three indirections, twice) and
// this library's own warp-per-middle k_doublets<0> (kept as fallback and A/B baseline).
//
// Work item = a group of up to 16 middles that are neighbours in the cell order (same reference
// bin, same r row, a few adjacent z cells; cut by k_cell_scan). Their candidate windows nearly
// coincide, so one warp
//   1. computes the union of the members' per-row z windows (cell_row_window per member, merged
//      with shared-memory atomics in cell space — a superset of every member's own window, and
//      every visited candidate still goes through the exact cuts, so the accepted set is the
//      reference's);
//   2. turns the (row, neighbour bin) windows into contiguous runs of the cell-ordered float4
//      array and lets the copy hardware gather them into ONE dense shared-memory tile
//      (cp.async.bulk per run completing on an mbarrier; STAGING == 1: 16-byte cp.async) —
//      the flattening of ~40 short runs costs no instructions. Runs are ordered by row, so
//      the tile is [lower rows: bottom candidates only][the members' own row][higher rows: top
//      candidates only];
//   3. runs the cuts with lanes = (member g, slice s): lane (g, s) tests member g against tile
//      positions s, s + S, ... (S = 32 / G; one broadcast LDS.128 per S candidates), two
//      candidates per iteration, with the direction-specific form of the first cut block in the
//      bottom / top regions. A lane keeps its verdicts as a bit per step (one 32-bit word per 32
//      steps in shared memory): no queue, no ballot in the loop;
//   4. with the per-member counts known, allocates the group's share of the arena once, expands
//      the bit words into a FIFO of (tile position, member) and computes lin_circle for the
//      doublets of ALL members at full lane occupancy; mid-tops are ranked by cotTheta inside
//      per-member shared-memory segments.
// Arena records, per-middle counts/offsets and the k_triplets work list are those of k_doublets.
// A group that outgrows the fixed-size tables (bit words, mid-top segments, 64 runs), or whose
// members see different neighbour bins (a spacepoint exactly on the upper z edge), is handed to
// k_doublets<2>, the warp-per-middle kernel.
#pragma once

#include "seed_kernels.cuh"

namespace b200seed {

#ifndef B200_TILE_STAGE
#define B200_TILE_STAGE 256
#endif
#ifndef B200_TILE_QT
#define B200_TILE_QT 256
#endif
#ifndef B200_TILE_WORDS
#define B200_TILE_WORDS 8
#endif
#ifndef B200_TILE_WARPS
#define B200_TILE_WARPS 8
#endif
#ifndef B200_TILE_MIN_CTAS
#define B200_TILE_MIN_CTAS 2
#endif
constexpr uint32_t TILE_GCAP = 16;                   // upper bound on the group size
constexpr uint32_t TILE_STAGE = B200_TILE_STAGE;     // candidates per staged chunk
constexpr uint32_t TILE_QT = B200_TILE_QT;           // mid-top doublets per group
constexpr uint32_t TILE_WORDS = B200_TILE_WORDS;     // verdict words per lane (32 steps each)
constexpr uint32_t TILE_RUNS = 64;                   // (row, neighbour bin) runs per group
constexpr uint32_t TILE_POS = 4096;                  // flattened candidate positions per group
constexpr uint32_t TILE_SEG = TILE_QT + 4 * TILE_GCAP;  // cotTheta segments, padded to 4
constexpr uint32_t TILE_FIFO = TILE_STAGE * 4;       // the expansion FIFO reuses the tile
constexpr int TILE_WARPS = B200_TILE_WARPS;
static_assert(TILE_FIFO >= 1024, "one verdict word of every lane must fit the FIFO");
static_assert(TILE_STAGE % 4 == 0 && TILE_QT % 4 == 0, "alignment");

// Shared memory of one warp.
struct __align__(16) TileWarp {
    float4 stage[TILE_STAGE];           // staged candidates {x,y,z,r}; later the expansion FIFO
    float4 mid[32];                     // the members {x,y,z,r}
    float cotT[TILE_SEG];               // per-member segments of mid-top cotTheta
    uint32_t keyT[TILE_SEG];            // ... and of their reference-order keys
    uint32_t qT[TILE_QT];               // mid-tops: (segment slot << 5) | member
    uint32_t auxT[TILE_QT];             // mid-tops: cell-ordered index
    uint32_t bits[TILE_WORDS][32];      // verdict words of every lane
    float2 midcs[32];                   // {xM / rM, yM / rM} of the members
    float2 midvar[32];                  // {varZ, varR} of the members
    uint32_t midpos[32];                // grid position of the members
    uint32_t run_excl[TILE_RUNS + 4];   // start of every run in the flattened space
    uint32_t run_src[TILE_RUNS];        // first cell-ordered index of every run
    uint32_t cntB[32], cntT[32], offB[32], offT[32], segT[32], runB[32], runT[32];
    int winlo[32], winhi[32];           // union z-cell window of every row
    uint8_t blk[TILE_POS / 8];          // run that holds flattened position 8 * i
    unsigned long long mbar;
    unsigned long long pad_;
};

struct TileArgs {
    const uint32_t* bin_off;
    const uint32_t* sorted_bin;
    const float2* var2;
    const uint32_t* cell_off;
    const float4* csp4;
    const uint32_t* ccanon;
    const uint32_t* group_list;
    uint32_t* cnt_b;
    uint32_t* cnt_t;
    uint32_t* off_b;
    uint32_t* off_t;
    DoubletRec* arena_b;
    DoubletRec* arena_t;
    Control* ctrl;
    CellGrid g;
    uint32_t max_doublets;
    uint32_t* fallback_list;
    uint32_t* active_list;
    uint32_t* seed_cnt;
    uint32_t n_sp;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// 1D bulk copy global -> shared through the TMA unit; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// The reference's division / square-root chain of the helix-radius cut, for the < 2 % of the
// pairs the polynomial pre-decision leaves open. Not inlined: the cut loop exists in three
// direction-specific copies with two candidates per iteration, and six inlined copies of this
// chain made the kernel outgrow the instruction cache (stall reason no_instruction 5.7 per
// issued instruction, issue-slot utilisation 31 %).
__device__ __noinline__ bool doublet_stage2_slow(float minHelixRadius2, float helixImpactMargin2, float x1,
                                                 float y1, float x2, float y2) {
    DevCfg c;
    c.minHelixRadius2 = minHelixRadius2;
    c.helixImpactMargin2 = helixImpactMargin2;
    return doublet_stage2(c, x1, y1, x2, y2);
}

template <int STAGING>  // 0: cp.async.bulk + mbarrier (TMA unit), 1: 16-byte cp.async per element
__global__ void __launch_bounds__(TILE_WARPS * 32, B200_TILE_MIN_CTAS)
k_doublets_tile(const DevCfg cfg, const TileArgs a) {
    extern __shared__ __align__(16) unsigned char s_tile_raw[];
    __shared__ unsigned long long s_pairs[2];
    __shared__ uint32_t s_acc[3];  // active, nb, nt
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltmask = lanemask_lt();
    constexpr uint32_t FULL = 0xffffffffu;
    TileWarp& W = reinterpret_cast<TileWarp*>(s_tile_raw)[warp];
    uint32_t* fifo = reinterpret_cast<uint32_t*>(W.stage);
    const uint32_t bar = smem_u32(&W.mbar);
    if (threadIdx.x == 0) {
        s_pairs[0] = s_pairs[1] = 0ull;
        s_acc[0] = s_acc[1] = s_acc[2] = 0u;
    }
    if (STAGING == 0 && lane == 0) {
        mbar_init(bar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    const uint32_t n_valid = a.ctrl->n_valid;
    const bool has_var = a.ctrl->has_variance != 0u;
    const bool bounded = cfg.fast_bounded != 0u;
    const CellGrid g = a.g;
    unsigned long long pairs = 0ull, visited = 0ull;
    uint32_t acc_active = 0, acc_nb = 0, acc_nt = 0;
    const uint32_t n_big = a.ctrl->n_group_big, n_work = n_big + a.ctrl->n_group_small;

    while (true) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&a.ctrl->ticket_g, 1u);
        t = __shfl_sync(FULL, t, 0);
        if (t >= n_work) break;
        const uint32_t desc = __ldg(a.group_list + (t < n_big ? t : a.n_sp - 1u - (t - n_big)));
        const uint32_t c0 = desc >> 5, G = (desc & 31u) + 1u;
        // lanes = (member gi, slice s); lanes beyond G * S idle in the cut loop
        const float invG = 1.f / float(G);
        const uint32_t S = div_small(32u, invG);
        const uint32_t s = div_small(lane, invG), gi = lane - s * G;
        const bool alive = s < S;

        if (lane < G) {
            const uint32_t pm = __ldg(a.ccanon + c0 + lane);
            const float4 Mc = __ldg(a.csp4 + c0 + lane);
            W.mid[lane] = Mc;
            W.midcs[lane] = make_float2(Mc.x / Mc.w, Mc.y / Mc.w);  // cosPhiM, sinPhiM of lin_circle
            W.midpos[lane] = pm;
            W.midvar[lane] = has_var ? __ldg(a.var2 + pm) : make_float2(0.f, 0.f);
            W.cntB[lane] = 0u;
            W.cntT[lane] = 0u;
        }
        __syncwarp();
        const float4 M = W.mid[gi];
        NeighbourWalk walk;
        walk.init(cfg, __ldg(a.sorted_bin + W.midpos[gi]), M.z);
        const bool uniform = __all_sync(
            FULL, walk.r0 == __shfl_sync(FULL, walk.r0, 0) && walk.z0 == __shfl_sync(FULL, walk.z0, 0) &&
                      walk.nz == __shfl_sync(FULL, walk.nz, 0) &&
                      walk.n_phi_seq == __shfl_sync(FULL, walk.n_phi_seq, 0));
        auto hand_back = [&]() {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&a.ctrl->n_fallback, G);
            base = __shfl_sync(FULL, base, 0);
            if (lane < G) a.fallback_list[base + lane] = W.midpos[lane];
            __syncwarp();
        };
        // rows that can hold a doublet partner of any member
        const float rmin = warp_min(M.w), rmax = warp_max(M.w);
        const float er = 1e-2f + 1e-5f * (rmax + absf(cfg.deltaRMax));
        const uint32_t row_lo = cell_row(g, rmin - cfg.deltaRMax - er);
        const uint32_t row_hi = cell_row(g, rmax + cfg.deltaRMax + er);
        const uint32_t nrows = row_hi - row_lo + 1u;  // <= NR <= 32
        const uint32_t ncombo = walk.nq * nrows;
        if (!uniform || ncombo > TILE_RUNS) {
            hand_back();
            continue;
        }
        // the reference tests every spacepoint of the neighbour bins against each member
        uint32_t bin_pop = 0;
        for (uint32_t q = lane; q < walk.nq; q += 32) {
            const uint32_t b = walk.bin(cfg, q);
            bin_pop += __ldg(a.bin_off + b + 1) - __ldg(a.bin_off + b);
        }
        if (lane < nrows) {
            W.winlo[lane] = 0x7fffffff;
            W.winhi[lane] = -1;
        }
        __syncwarp();
        if (alive) {
            for (uint32_t ri = s; ri < nrows; ri += S) {
                float L, U;
                if (cell_row_window(cfg, g, M.w, M.z, row_lo + ri, L, U)) {
                    atomicMin(&W.winlo[ri], cell_zg(g, L));
                    atomicMax(&W.winhi[ri], cell_zg(g, U));
                }
            }
        }
        __syncwarp();

        // ---- runs: combination j = row * nq + neighbour bin, two per lane ----
        const float inv_nq = 1.f / float(walk.nq);
        uint32_t lo[2], len[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t j = 32u * k + lane;
            lo[k] = len[k] = 0;
            if (j < ncombo) {
                const uint32_t ri = div_small(j, inv_nq), q = j - ri * walk.nq;
                const int zlo = W.winlo[ri], zhi = W.winhi[ri];
                if (zhi >= zlo) {
                    const uint32_t zb = walk.zbin(q);
                    const uint32_t base = walk.bin(cfg, q) * g.CPB + (row_lo + ri) * g.NZc;
                    lo[k] = __ldg(a.cell_off + base + cell_z_of(g, zb, zlo));
                    len[k] = __ldg(a.cell_off + base + cell_z_of(g, zb, zhi) + 1u) - lo[k];
                }
            }
        }
        const uint32_t incl0 = warp_incl_scan(len[0], lane);
        const uint32_t tot0 = __shfl_sync(FULL, incl0, 31);
        const uint32_t incl1 = warp_incl_scan(len[1], lane);
        const uint32_t total = tot0 + __shfl_sync(FULL, incl1, 31);
        const uint32_t ex[2] = {incl0 - len[0], tot0 + incl1 - len[1]};
        // steps of the cut loop: step u covers flattened positions u * S ... u * S + S - 1
        const uint32_t nsteps = ((total + S - 1u) / (S ? S : 1u) + 1u) & ~1u;  // even
        if (total > TILE_POS || nsteps > TILE_WORDS * 32u) {
            hand_back();
            continue;
        }
        W.run_excl[lane] = ex[0];
        W.run_excl[32u + lane] = ex[1];
        W.run_src[lane] = lo[0];
        W.run_src[32u + lane] = lo[1];
        if (lane == 0) W.run_excl[TILE_RUNS] = total;
#pragma unroll
        for (int k = 0; k < 2; ++k)  // the run that holds every 8th position
            for (uint32_t b = (ex[k] + 7u) >> 3; b <= ((ex[k] + len[k] - 1u) >> 3) && len[k] != 0u; ++b)
                W.blk[b] = uint8_t(32u * k + lane);
        // region boundaries in the flattened space: rows below the members' row hold bottom
        // candidates only, rows above top candidates only
        const uint32_t my_ri = cell_row(g, M.w) - row_lo;  // the same for all members
        const uint32_t jb = my_ri * walk.nq, jt = (my_ri + 1u) * walk.nq;
        __syncwarp();
        const uint32_t posB = W.run_excl[jb < TILE_RUNS ? jb : TILE_RUNS];   // end of the bottom region
        const uint32_t posT = W.run_excl[jt < TILE_RUNS ? jt : TILE_RUNS];   // start of the top region
        // in steps, rounded to pairs of steps: [0, uB) bottom form, [uB, uT) generic, [uT, nsteps) top
        const uint32_t uB = (posB / S) & ~1u;
        uint32_t uT = ((posT + S - 1u) / S + 1u) & ~1u;
        if (uT > nsteps) uT = nsteps;

        uint32_t accw = 0;           // verdict bits of the current word of this lane
        uint32_t mixB = 0, mixT = 0; // doublets of this lane found in the generic range
        const uint32_t chunk_steps = (TILE_STAGE / S) & ~1u;  // steps per staged chunk (even)
        for (uint32_t u0 = 0; u0 < nsteps; u0 += chunk_steps) {
            const uint32_t u1 = (u0 + chunk_steps < nsteps) ? (u0 + chunk_steps) : nsteps;
            const uint32_t w0 = u0 * S;
            const uint32_t n = ((u1 * S < total) ? (u1 * S) : total) - w0;
            // ---- stage candidates [w0, w0 + n) of the flattened runs ----
            if (STAGING == 0) {
                if (lane == 0) mbar_expect_tx(bar, n * 16u);
                __syncwarp();
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t b0 = (ex[k] > w0) ? ex[k] : w0;
                const uint32_t e1 = ex[k] + len[k];
                const uint32_t b1 = (e1 < w0 + n) ? e1 : (w0 + n);
                if (b0 < b1) {
                    const float4* src = a.csp4 + lo[k] + (b0 - ex[k]);
                    const uint32_t dst = smem_u32(&W.stage[b0 - w0]);
                    if (STAGING == 0) {
                        bulk_g2s(dst, src, (b1 - b0) * 16u, bar);
                    } else {
                        for (uint32_t i = 0; i < b1 - b0; ++i) cp_async16(dst + 16u * i, src + i);
                    }
                }
            }
            if (STAGING == 0) {
                mbar_wait(bar, parity);
                parity ^= 1u;
            } else {
                cp_async_wait_all();
                __syncwarp();
            }
            // ---- cuts: lane (gi, s) tests member gi against positions s, s + S, ...; two steps
            //      per iteration; verdict of step u = bit (u & 31) of word u / 32 ----
            auto helix = [&](const float4& P) -> bool {
                int d = bounded ? doublet_stage2_fast_bounded(cfg, M.x, M.y, P.x, P.y)
                                : doublet_stage2_fast(cfg, M.x, M.y, P.x, P.y);
                if (d == 2)
                    d = doublet_stage2_slow(cfg.minHelixRadius2, cfg.helixImpactMargin2, M.x, M.y, P.x, P.y) ? 1 : 0;
                return d != 0;
            };
            auto range = [&](uint32_t ua, uint32_t ub, auto&& test) {
                for (uint32_t u = ua; u < ub; u += 2u) {
                    const uint32_t pa = u * S + s - w0, pb = pa + S;
                    uint32_t v = 0;
                    if (alive) {
                        if (pa < n && test(W.stage[pa], 0u)) v = 1u;
                        if (pb < n && test(W.stage[pb], 1u)) v |= 2u;
                    }
                    accw |= v << (u & 31u);
                    if ((u & 31u) == 30u) {
                        W.bits[u >> 5][lane] = accw;
                        accw = 0;
                    }
                }
            };
            const uint32_t ca = (u0 > uB) ? u0 : uB, cb = (u1 < uT) ? u1 : uT;
            if (u0 < uB)
                range(u0, (u1 < uB) ? u1 : uB, [&](const float4& P, uint32_t) {
                    return doublet_stage1_bottom(cfg, M.w, M.z, P.w, P.z) && helix(P);
                });
            if (ca < cb)
                range(ca, cb, [&](const float4& P, uint32_t) {
                    const int st = doublet_stage1(cfg, M.w, M.z, P.w, P.z);
                    if (st == 0 || !helix(P)) return false;
                    if (st == 1) ++mixB; else ++mixT;
                    return true;
                });
            if (u1 > uT)
                range((u0 > uT) ? u0 : uT, u1, [&](const float4& P, uint32_t) {
                    return doublet_stage1_top(cfg, M.w, M.z, P.w, P.z) && helix(P);
                });
            __syncwarp();  // every lane is done with the tile before it is overwritten
        }
        if (nsteps & 31u) W.bits[nsteps >> 5][lane] = accw;  // the last, partial word
        const uint32_t nwords = (nsteps + 31u) >> 5;
        __syncwarp();

        // ---- per-lane counts: verdict bits below step uB are bottoms, from uT on tops ----
        {
            uint32_t cB = mixB, cT = mixT;
            for (uint32_t w = 0; w < nwords; ++w) {
                const uint32_t m = W.bits[w][lane];
                const uint32_t b = w * 32u;
                // bits of this word below uB / at or above uT
                const uint32_t nb_bits = (uB > b) ? ((uB - b >= 32u) ? 32u : (uB - b)) : 0u;
                const uint32_t nt_from = (uT > b) ? ((uT - b >= 32u) ? 32u : (uT - b)) : 0u;
                const uint32_t maskB = (nb_bits >= 32u) ? FULL : ((1u << nb_bits) - 1u);
                const uint32_t maskT = (nt_from >= 32u) ? 0u : ~((1u << nt_from) - 1u);
                cB += __popc(m & maskB);
                cT += __popc(m & maskT);
            }
            if (alive && cB) atomicAdd(&W.cntB[gi], cB);
            if (alive && cT) atomicAdd(&W.cntT[gi], cT);
        }
        __syncwarp();
        // ---- per-member results (lane gi < G, slice 0, holds one member) ----
        const uint32_t nBm = (lane < G) ? W.cntB[lane] : 0u, nTm = (lane < G) ? W.cntT[lane] : 0u;
        const bool act = nBm != 0u && nTm != 0u;  // seed_finding.cpp:85-95
        uint32_t actmask = __ballot_sync(FULL, act);
        uint32_t aB = act ? nBm : 0u, aT = act ? nTm : 0u;
        const uint32_t inclB = warp_incl_scan(aB, lane), inclT = warp_incl_scan(aT, lane);
        const uint32_t totB = __shfl_sync(FULL, inclB, 31), totT = __shfl_sync(FULL, inclT, 31);
        const uint32_t aT4 = (aT + 3u) & ~3u;
        const uint32_t inclS = warp_incl_scan(aT4, lane);
        if (totT > TILE_QT) {  // the mid-top segments do not hold this group
            hand_back();
            continue;
        }
        visited += (lane == 0) ? (unsigned long long)nsteps * S * G : 0ull;
        pairs += (unsigned long long)bin_pop * G;  // per-lane partial sums, reduced at the end
        uint32_t baseB = 0, baseT = 0;
        if (lane == 0 && totB != 0u) {
            baseB = atomicAdd(&a.ctrl->cursor[0], totB);
            baseT = atomicAdd(&a.ctrl->cursor[1], totT);
        }
        baseB = __shfl_sync(FULL, baseB, 0);
        baseT = __shfl_sync(FULL, baseT, 0);
        if (baseB > a.max_doublets || totB > a.max_doublets - baseB || baseT > a.max_doublets ||
            totT > a.max_doublets - baseT) {
            if (lane == 0) atomicOr(&a.ctrl->overflow, B200SEED_OVF_DOUBLETS);
            actmask = 0u;
            aB = aT = 0u;
        }
        uint32_t m = 0;
        if (lane < G) {
            m = W.midpos[lane];
            W.offB[lane] = baseB + inclB - aB;
            W.offT[lane] = baseT + inclT - aT;
            W.segT[lane] = inclS - aT4;
            W.cntT[lane] = aT;
            W.runB[lane] = 0u;
            W.runT[lane] = 0u;
            a.cnt_b[m] = aB;
            a.cnt_t[m] = aT;
            a.off_b[m] = baseB + inclB - aB;
            a.off_t[m] = baseT + inclT - aT;
            if (aB == 0u) a.seed_cnt[m] = 0u;
        }
        {
            // work list of k_triplets (see work_class)
            if (aB != 0u) work_push(a.active_list, a.n_sp, a.ctrl, m, aB, aT);
        }
        acc_active += __popc(actmask);
        acc_nb += (actmask ? totB : 0u);
        acc_nt += (actmask ? totT : 0u);
        __syncwarp();
        if (actmask == 0u) continue;

        // ---- expansion: verdict words -> FIFO of (flattened position << 5 | member) -> records ----
        uint32_t wT = 0;  // mid-tops queued so far (warp-uniform)
        const bool my_act = alive && ((actmask >> gi) & 1u);
        for (uint32_t w = 0; w < nwords; ++w) {
            uint32_t mbits = my_act ? W.bits[w][lane] : 0u;
            const uint32_t cnt = __popc(mbits);
            const uint32_t incl = warp_incl_scan(cnt, lane);
            const uint32_t nf = __shfl_sync(FULL, incl, 31);
            if (nf == 0u) continue;
            uint32_t k = incl - cnt;
            while (mbits) {
                const uint32_t b = __ffs(int(mbits)) - 1u;
                mbits &= mbits - 1u;
                fifo[k++] = (((w * 32u + b) * S + s) << 5) | gi;
            }
            __syncwarp();
            for (uint32_t i0 = 0; i0 < nf; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool on = i < nf;
                const uint32_t e = on ? fifo[i] : 0u;
                const uint32_t gg = e & 31u, pl = e >> 5;
                // flattened position -> run -> cell-ordered index
                uint32_t r = on ? W.blk[pl >> 3] : 0u;
                while (on && W.run_excl[r + 1u] <= pl) ++r;
                const uint32_t c = W.run_src[r] + (pl - W.run_excl[r]);
                const float4 Mg = W.mid[gg];
                float4 P = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t pos = 0;
                if (on) {
                    P = __ldg(a.csp4 + c);
                    pos = __ldg(a.ccanon + c);
                }
                // direction: by region; inside the members' own row by the sign of rM - r2
                const uint32_t u = div_small(pl, 1.f / float(S));
                const bool top = (u >= uT) || (u >= uB && (Mg.w - P.w) < 0.f);
                const bool isT = on && top;
                // position inside the member's list = number of earlier entries of the member
                const uint32_t peers =
                    __match_any_sync(FULL, on ? (gg | (top ? 32u : 0u)) : (64u + lane));
                const uint32_t rk = __popc(peers & ltmask);
                uint32_t before = 0;
                if (on) before = top ? W.runT[gg] : W.runB[gg];
                __syncwarp();
                if (on && rk == 0u) (top ? W.runT : W.runB)[gg] = before + __popc(peers);
                __syncwarp();
                const uint32_t mT = __ballot_sync(FULL, isT);
                if (on) {
                    const float2 cs = W.midcs[gg];
                    const float2 VM = W.midvar[gg];
                    const float2 V = has_var ? __ldg(a.var2 + pos) : make_float2(0.f, 0.f);
                    const LinCircle l = transform_coordinates_cs(!top, cs.x, cs.y, Mg.x, Mg.y, Mg.z, Mg.w,
                                                                 VM.x, VM.y, P.x, P.y, P.z, V.x, V.y);
                    if (!top) {
                        DoubletRec rec;
                        rec.a = make_float4(l.cotTheta, l.iDeltaR, l.Er, l.U);
                        rec.b = make_float4(l.V, l.Zo, P.w, __uint_as_float(pos));
                        a.arena_b[W.offB[gg] + before + rk] = rec;
                    } else {
                        // mid-top, pass 1: cotTheta and reference-order key into the member's segment;
                        // the record itself waits for its rank
                        const uint32_t ls = W.segT[gg] + before + rk;
                        const uint32_t wphi = walk.wphi(r - div_small(r, inv_nq) * walk.nq);
                        W.cotT[ls] = l.cotTheta;
                        W.keyT[ls] = canon_key(wphi, n_valid, pos);
                        const uint32_t qi = wT + __popc(mT & ltmask);
                        W.qT[qi] = (ls << 5) | gg;
                        W.auxT[qi] = c;
                    }
                }
                wT += __popc(mT);
            }
            __syncwarp();
        }
        if (lane < G) {  // pad every segment to a multiple of four for the vectorised rank loop
            const uint32_t nT = W.cntT[lane], seg = W.segT[lane];
            for (uint32_t k = nT; k < ((nT + 3u) & ~3u); ++k) W.cotT[seg + k] = __uint_as_float(0x7f800000u);
        }
        __syncwarp();
        // ---- mid-tops, pass 2: position in the (cotTheta, reference order) sort, record ----
        for (uint32_t i0 = 0; i0 < wT; i0 += 32) {
            const uint32_t i = i0 + lane;
            if (i >= wT) continue;
            const uint32_t e = W.qT[i];
            const uint32_t gg = e & 31u, ls = e >> 5;
            const uint32_t seg = W.segT[gg], nT = W.cntT[gg];
            const float ck = W.cotT[ls];
            const uint32_t kk = W.keyT[ls];
            uint32_t lt = 0, eq = 0;
            const float4* c4 = reinterpret_cast<const float4*>(W.cotT + seg);
            for (uint32_t j = 0; j < (nT + 3u) / 4u; ++j) {
                const float4 v = c4[j];
                lt += (v.x < ck) ? 1u : 0u;
                lt += (v.y < ck) ? 1u : 0u;
                lt += (v.z < ck) ? 1u : 0u;
                lt += (v.w < ck) ? 1u : 0u;
                eq += (v.x == ck) ? 1u : 0u;
                eq += (v.y == ck) ? 1u : 0u;
                eq += (v.z == ck) ? 1u : 0u;
                eq += (v.w == ck) ? 1u : 0u;
            }
            uint32_t rank = lt;
            if (eq > 1u)  // equal cotTheta values: the reference order decides
                rank = top_rank([&](uint32_t j) { return W.cotT[seg + j]; },
                                [&](uint32_t j) { return W.keyT[seg + j]; }, nT, ck, kk);
            if (rank >= nT) rank = nT - 1u;  // NaN keys: stay inside the list
            const uint32_t c = W.auxT[i];
            const float4 Mg = W.mid[gg];
            const float2 cs = W.midcs[gg];
            const float2 VM = W.midvar[gg];
            const float4 P = __ldg(a.csp4 + c);
            const uint32_t pos = __ldg(a.ccanon + c);
            const float2 V = has_var ? __ldg(a.var2 + pos) : make_float2(0.f, 0.f);
            const LinCircle l = transform_coordinates_cs(false, cs.x, cs.y, Mg.x, Mg.y, Mg.z, Mg.w, VM.x,
                                                         VM.y, P.x, P.y, P.z, V.x, V.y);
            DoubletRec rec;
            rec.a = make_float4(ck, l.iDeltaR, l.Er, l.U);
            rec.b = make_float4(l.V, __uint_as_float(kk), P.w, __uint_as_float(pos));
            a.arena_t[W.offT[gg] + rank] = rec;
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(FULL, pairs, o);
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
}

}  // namespace b200seed
