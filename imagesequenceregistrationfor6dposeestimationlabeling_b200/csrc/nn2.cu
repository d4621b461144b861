// K2 (production) -- brute-force nearest neighbour: FP32 filter scan + exact FP64 resolve.
//
// Replaces the KD-tree 1-NN of Open3D (verfication.py:97,99; icp.py:97-103,113,115) and
// sklearn (choosePose.py:21-22).  Tiled brute force: a (256-query warp block, 64-target
// sub-tile) pair is either scanned completely or -- when the target carries bounding spheres
// and pruning is on -- skipped because the sphere proves that none of its points can beat
// (or tie) the exact neighbour already found for any of the warp's queries.  With pruning
// off every pair is visited.  The per-pair work is the cheapest FP32 form that can be made
// exact:
//
//   scan    a_j = fma(-2qx, px, fma(-2qy, py, fma(-2qz, pz, |p_j|^2)))   ~ d_j^2 - |q|^2
//           3 FFMA per pair (packed: 3 FFMA2 per target pair) + 1/2 FMNMX3, all on the
//           FP32 CUDA cores.  a_j carries cancellation error, so it is only a FILTER:
//           |a_j - A_j| <= 13 u R^2 (u = 2^-24, R = |q| + d), proven from the three fma
//           roundings and the rounding of |p|^2 (clouds are centred by prepare.cu so R is
//           the object radius, not the camera distance).
//   flag    per 32-target sub-tile one compare per query: does the sub-tile minimum come
//           within the window W of the running minimum?  (one FSETP per query per 32 targets)
//   resolve only then: re-derive the sub-tile's a_j, and for the targets inside the window
//           compute the distance EXACTLY in FP64 from the hi/lo coordinates; keep the
//           smallest (strict <, ascending index => lowest index on exact ties).
//
// Exactness: let j* be the true nearest neighbour and m the running minimum of a when j* is
// visited.  m = a_i for some visited i with D_i >= D_j*, hence
//     a_j* <= D_j* - |q|^2 + E <= D_i - |q|^2 + E <= a_i + 2E = m + 2E <= m + W,
// so j*'s sub-tile is flagged and j* is inside the window: it is always resolved exactly.
// The returned index therefore equals the float64 brute-force argmin of the prepared
// (hi+lo) coordinates; the returned d2 is that FP64 distance rounded once to float32.
//
// Pruning (PRUNE variant).  Clouds are stored in Morton order, so a warp's 256 queries, a
// 1024-target stage and a 64-target sub-tile are all compact patches.  Every query carries
// dq >= its exact best distance so far (FP64 Dbest from the resolve path, rounded up, plus
// the size of its lo part); a tile with sphere (c, r) is skipped by a warp iff for every
// query |q - c| > dq + r (evaluated in FP32 with a 1e-4 relative margin, the sphere radius
// being inflated by prepare.cu for rounding and for the targets' lo parts): then every point
// of the tile is strictly farther than the neighbour already held, so it can be neither the
// minimum nor an equal-distance tie.  Order of work per CTA: (1) the stage nearest to the
// query block, each warp starting with its nearest sub-tile, which gives every query a
// near-final bound; (2) a list of the stages whose sphere comes within the CTA-wide bound
// (all others are never even loaded); (3) those stages, each warp testing the stage sphere
// and then the 16 sub-tile spheres, which ride along with the stage in shared memory.
// The exactness argument above is untouched: the true neighbour's tile is never skipped
// (its distance is <= every bound), so it is visited, flagged and resolved as before.
//
// Roofline: FP32 CUDA cores.  Algorithmic work stays 8 flop per pair;
// executed FP32-pipe work is 3 lane-ops per pair, so the algorithmic rate can exceed the
// nominal FMA peak (cap 8/6) -- bench.py reports both.
#include <math_constants.h>
#include <stdlib.h>

#include "isr_common.cuh"

namespace isr {

struct NN2Params {
    const float *q;          // SoA7 [batch][7][nq_pad]
    long long q_bstride;
    int nq;
    int nq_pad;
    const float *t;          // SoA7 [batch][7][nt_pad]
    long long t_bstride;
    int nt_pad;
    int nt;
    float *out_d2;
    int *out_idx;
    double *part_D;          // [splits][batch*nq] when the target range is split
    int *part_idx;
    long long part_stride;
    const int *skip;
    long long skip_stride;
    int stages_total;
    int stages_per_split;
    int use_lo;
    int debug_no_resolve;  // tuning only: never flag (pure filter-scan rate)
    const float4 *stage_c;   // [batch][stages_total] target stage centroids, or NULL
    long long stage_c_bstride;
    const int *perm_q;       // stored position -> original index (NULL = identity)
    const int *perm_t;
    unsigned long long *dbg;  // tuning only: event counters
    const float4 *sub_c;     // [batch][stages_total * STAGE/SUB] sub-tile spheres (PRUNE)
    long long sub_c_bstride;
    unsigned long long *evaluated;  // profiling: scanned (warp, sub-tile) units, or NULL
};

// Can this lane rule out every point of the tile with sphere S for all of its Q queries?
// q2* = -2 * query (hi part); dmax >= best distance so far of each of the lane's live queries
// (0 for a lane without live queries, whose padded coordinates are far from everything).
template <int Q>
__device__ __forceinline__ bool lane_rules_out(const float4 S, const float (&q2x)[Q],
                                               const float (&q2y)[Q], const float (&q2z)[Q],
                                               float dmax) {
    float m = CUDART_INF_F;
#pragma unroll
    for (int r = 0; r < Q; ++r) {
        const float dx = fmaf(q2x[r], -0.5f, -S.x), dy = fmaf(q2y[r], -0.5f, -S.y),
                    dz = fmaf(q2z[r], -0.5f, -S.z);
        m = fminf(m, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    }
    const float rr = (dmax + S.w) * 1.0001f;
    return S.w < 0.f || m > rr * rr;
}

// Upper threshold for "could still be the nearest neighbour": running minimum + window.
// mt = running min of a (~ d^2 - |q|^2), nq2 = |q|^2, qn = |q| (float32, hi part).
// W >= 2 (13 u R^2 + 8 u R d) with R = |q| + d and d an upper bound on the best distance;
// every constant carries slack for the float32 evaluation of this very formula.
__device__ __forceinline__ float filter_threshold(float mt, float nq2, float qn) {
    const float u = 5.9604645e-8f;  // 2^-24
    const float s = sqrtf(fmaxf(mt + nq2, 0.f));
    const float dub = (s + 1.5e-3f * qn) * 1.002f + 1e-20f;
    const float R = (qn + dub) * 1.00001f;
    const float W = (26.f * u * R * R + 16.f * u * R * dub) * 1.001f + 1e-30f;
    return __fadd_ru(mt, W);
}

constexpr int kListCap = 512;  // stages a pruning CTA can list; more -> it walks all of them

template <int Q, int THREADS, int STAGE, int NSTAGES, int SUB, int MINB, int UNR, bool PRUNE>
__global__ void __launch_bounds__(THREADS, MINB) nn2_kernel(const NN2Params p) {
    static_assert(STAGE % SUB == 0 && SUB % 8 == 0 && Q <= 16, "tile shapes");
    static_assert(!PRUNE || (SUB == ISR_SUB_TILE && STAGE == ISR_SOA_TILE && STAGE / SUB <= 32),
                  "pruning uses the spheres of prepare.cu");
    constexpr int SUBS = STAGE / SUB;
    constexpr int WARPS = THREADS / 32;
    constexpr int SLOT_F = 4 * STAGE + (PRUNE ? 4 * SUBS : 0);  // floats per pipeline slot
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sbuf = reinterpret_cast<float *>(smem_raw);  // [NSTAGES][4][STAGE] (+ [SUBS] spheres)
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSTAGES * SLOT_F * 4);
    int *slist = reinterpret_cast<int *>(full + NSTAGES);  // [kListCap], PRUNE only

    const int b = blockIdx.z;
    if (p.skip != nullptr && p.skip[(long long)b * p.skip_stride] != 0) return;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const float *__restrict__ gq = p.q + (long long)b * p.q_bstride;
    const float *__restrict__ gt = p.t + (long long)b * p.t_bstride;
    const int s_begin = blockIdx.y * p.stages_per_split;
    const int nst = min(p.stages_per_split, p.stages_total - s_begin);

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NSTAGES; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // Scan state, in registers: -2 * query (hi part) and the flag threshold.
    float q2x[Q], q2y[Q], q2z[Q], thr[Q];
    // Resolve state, touched only on the rare path and indexed at run time there, which
    // places it in (L1-resident) local memory and keeps it out of the scan's registers.
    float mt_l[Q], thr_l[Q], tm_l[Q];
    double Dbest_l[Q];
    int ibest_l[Q];
    // a warp owns 32*Q consecutive stored queries (a compact patch under Morton order);
    // lane l holds queries l, 32+l, ...: every load below is one coalesced 128-byte line
    const int q0 = blockIdx.x * (THREADS * Q) + warp * (32 * Q) + lane;
    float dq_l[Q];       // PRUNE: per-query upper bound of the best distance so far
    float dmax = 0.f;    // PRUNE: max of dq_l over the lane's live queries
#pragma unroll
    for (int r = 0; r < Q; ++r) {
        const int i = min(q0 + r * 32, p.nq_pad - 1);
        q2x[r] = -2.0f * gq[i];
        q2y[r] = -2.0f * gq[p.nq_pad + i];
        q2z[r] = -2.0f * gq[2ll * p.nq_pad + i];
        // padded query slots (i >= nq) must never reach the resolve path: their 1e18
        // coordinates would put every target inside the error window
        const bool live = (q0 + r * 32 < p.nq) && !p.debug_no_resolve;
        thr[r] = live ? CUDART_INF_F : -CUDART_INF_F;
        if (PRUNE && q0 + r * 32 < p.nq) dmax = CUDART_INF_F;
    }
    for (int r = 0; r < Q; ++r) {  // run-time loop on purpose
        mt_l[r] = CUDART_INF_F;
        thr_l[r] = (q0 + r * 32 < p.nq) && !p.debug_no_resolve ? CUDART_INF_F : -CUDART_INF_F;
        Dbest_l[r] = CUDART_INF;
        ibest_l[r] = s_begin * STAGE;
        if (PRUNE) dq_l[r] = (q0 + r * 32 < p.nq) ? CUDART_INF_F : 0.f;
    }

    // Scan order: the target stage whose centre is nearest to this CTA's query block goes
    // first (clouds are stored in Morton order, so both are compact patches).  After that
    // one stage every query already holds a near-final bound and the remaining stages
    // almost never reach the resolve path.  The rest follows in ascending order, or -- when
    // pruning -- only the listed stages follow.
    int s_first = 0;
    float cQx = 0.f, cQy = 0.f, cQz = 0.f, rQ = 0.f;  // PRUNE: query-block sphere
    float cWx = 0.f, cWy = 0.f, cWz = 0.f;            // PRUNE: this warp's query centroid
    __shared__ float cred[4][WARPS];
    __shared__ u64 sred[WARPS];
    __shared__ int scount;
    if (p.stage_c != nullptr && (nst > 1 || PRUNE)) {
        float cx = 0.f, cy = 0.f, cz = 0.f, cn = 0.f;
#pragma unroll
        for (int r = 0; r < Q; ++r) {
            if (q0 + r * 32 < p.nq) {
                cx += q2x[r]; cy += q2y[r]; cz += q2z[r]; cn += 1.f;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cx += __shfl_xor_sync(0xffffffffu, cx, o);
            cy += __shfl_xor_sync(0xffffffffu, cy, o);
            cz += __shfl_xor_sync(0xffffffffu, cz, o);
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
        }
        if (PRUNE) {
            const float invw = cn > 0.f ? -0.5f / cn : 0.f;  // q2 = -2 * query
            cWx = cx * invw; cWy = cy * invw; cWz = cz * invw;
        }
        if (lane == 0) {
            cred[0][warp] = cx; cred[1][warp] = cy; cred[2][warp] = cz; cred[3][warp] = cn;
        }
        __syncthreads();
        cx = cy = cz = cn = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            cx += cred[0][w]; cy += cred[1][w]; cz += cred[2][w]; cn += cred[3][w];
        }
        const float inv = cn > 0.f ? -0.5f / cn : 0.f;  // q2 = -2 * query
        cx *= inv; cy *= inv; cz *= inv;
        const float4 *sc = p.stage_c + (long long)b * p.stage_c_bstride + s_begin;
        u64 best = ~0ull;
        for (int s = tid; s < nst; s += THREADS) {
            const float4 c = sc[s];
            const float dx = c.x - cx, dy = c.y - cy, dz = c.z - cz;
            const float d = dx * dx + dy * dy + dz * dz;
            const u64 key = ((u64)__float_as_uint(d) << 32) | (u64)(unsigned)s;
            best = key < best ? key : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u64 other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        if (lane == 0) sred[warp] = best;
        if (PRUNE) {
            // radius of the query block about (cx, cy, cz), live queries only
            float m = 0.f;
#pragma unroll
            for (int r = 0; r < Q; ++r) {
                if (q0 + r * 32 < p.nq) {
                    const float dx = fmaf(q2x[r], -0.5f, -cx), dy = fmaf(q2y[r], -0.5f, -cy),
                                dz = fmaf(q2z[r], -0.5f, -cz);
                    m = fmaxf(m, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            __syncthreads();  // cred is reused
            if (lane == 0) cred[0][warp] = m;
            if (tid == 0) scount = 0;
        }
        __syncthreads();
        best = sred[0];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) best = sred[w] < best ? sred[w] : best;
        s_first = (int)(unsigned)(best & 0xffffffffull);
        if (s_first >= nst) s_first = 0;
        if (PRUNE) {
            float m = cred[0][0];
#pragma unroll
            for (int w = 1; w < WARPS; ++w) m = fmaxf(m, cred[0][w]);
            cQx = cx; cQy = cy; cQz = cz;
            rQ = __fsqrt_ru(m) * 1.00002f;
        }
    }
    // pipeline position -> stage (relative to s_begin).  Position 0 is s_first; then either
    // all other stages in ascending order or (PRUNE, after the list is built) the list.
    bool use_list = false;
    int npos = PRUNE ? 1 : nst;
    auto stage_of = [&](int pos) {
        if (pos == 0) return s_first;
        if (PRUNE && use_list) return slist[pos - 1];
        return pos - 1 < s_first ? pos - 1 : pos;
    };

    auto issue = [&](int sl) {
        const int slot = sl % NSTAGES;
        float *dst = sbuf + (size_t)slot * SLOT_F;
        const int stg = s_begin + stage_of(sl);
        const float *src = gt + (long long)stg * STAGE;
        mbar_expect_tx(&full[slot], 4u * STAGE * 4u + (PRUNE ? SUBS * 16u : 0u));
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
            bulk_g2s(dst + pl * STAGE, src + (long long)pl * p.nt_pad, STAGE * 4u, &full[slot]);
        if (PRUNE)
            bulk_g2s(dst + 4 * STAGE, p.sub_c + (long long)b * p.sub_c_bstride + (long long)stg * SUBS,
                     SUBS * 16u, &full[slot]);
    };
    if (tid == 0) {
        for (int i = 0; i < NSTAGES - 1 && i < npos; ++i) issue(i);
    }
    unsigned nscanned = 0;  // (warp, sub-tile) units this warp evaluated

    // (A per-slot `empty` mbarrier instead of the CTA barrier below was measured 2 % slower:
    // the 5 % of samples parked at the barrier are warps that would otherwise only run ahead.)
    for (int sl = 0; sl < npos; ++sl) {
        if (tid == 0 && sl + NSTAGES - 1 < npos) issue(sl + NSTAGES - 1);
        const int slot = sl % NSTAGES;
        bool skip_stage = false;
        if (PRUNE && sl > 0) {
            // the whole stage first: one sphere test per warp (the sphere comes from L2
            // while the stage's bulk copy is still in flight)
            const float4 S = p.stage_c[(long long)b * p.stage_c_bstride + s_begin + stage_of(sl)];
            skip_stage = __all_sync(0xffffffffu, lane_rules_out<Q>(S, q2x, q2y, q2z, dmax));
        }
        mbar_wait(&full[slot], (sl / NSTAGES) & 1);
        const float4 *sx = reinterpret_cast<const float4 *>(sbuf + (size_t)slot * SLOT_F);
        const float4 *sy = sx + STAGE / 4;
        const float4 *sz = sy + STAGE / 4;
        const float4 *sn = sz + STAGE / 4;
        const float4 *ssph = sn + STAGE / 4;  // PRUNE: the stage's SUBS sub-tile spheres
        int sub_first = 0;
        if (PRUNE && sl == 0) {
            // no bound yet: start with the sub-tile nearest to this warp's queries
            float d = CUDART_INF_F;
            if (lane < SUBS) {
                const float4 S = ssph[lane];
                const float dx = S.x - cWx, dy = S.y - cWy, dz = S.z - cWz;
                d = S.w < 0.f ? CUDART_INF_F : dx * dx + dy * dy + dz * dz;
            }
            u64 key = ((u64)__float_as_uint(d) << 32) | (u64)(unsigned)lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const u64 other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other < key ? other : key;
            }
            sub_first = (int)(unsigned)(key & 31ull);
            if (sub_first >= SUBS) sub_first = 0;
        }
#pragma unroll 1
        for (int k = 0; k < (skip_stage ? 0 : SUBS); ++k) {
            int sub = k;
            if (PRUNE) {
                if (sl == 0) sub = k == 0 ? sub_first : (k <= sub_first ? k - 1 : k);
                if (__all_sync(0xffffffffu, lane_rules_out<Q>(ssph[sub], q2x, q2y, q2z, dmax))) continue;
            }
            ++nscanned;
            float tm[Q];
#pragma unroll
            for (int r = 0; r < Q; ++r) tm[r] = CUDART_INF_F;
#pragma unroll UNR
            for (int g = 0; g < SUB / 4; ++g) {
                const float4 X = sx[sub * (SUB / 4) + g];
                const float4 Y = sy[sub * (SUB / 4) + g];
                const float4 Z = sz[sub * (SUB / 4) + g];
                const float4 N = sn[sub * (SUB / 4) + g];
                const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
                const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
                const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
                const u64 n01 = pack2(N.x, N.y), n23 = pack2(N.z, N.w);
                // level-major over the Q queries: consecutive FFMA2 are independent and share
                // the target-pair operand (operand-reuse cache); the query rides as a scalar
                u64 acc[Q];
#pragma unroll
                for (int r = 0; r < Q; ++r) acc[r] = fma2(pack2(q2z[r], q2z[r]), z01, n01);
#pragma unroll
                for (int r = 0; r < Q; ++r) acc[r] = fma2(pack2(q2y[r], q2y[r]), y01, acc[r]);
#pragma unroll
                for (int r = 0; r < Q; ++r) acc[r] = fma2(pack2(q2x[r], q2x[r]), x01, acc[r]);
#pragma unroll
                for (int r = 0; r < Q; ++r) {
                    float a0, a1;
                    unpack2(acc[r], a0, a1);
                    tm[r] = min3(tm[r], a0, a1);
                }
#pragma unroll
                for (int r = 0; r < Q; ++r) acc[r] = fma2(pack2(q2z[r], q2z[r]), z23, n23);
#pragma unroll
                for (int r = 0; r < Q; ++r) acc[r] = fma2(pack2(q2y[r], q2y[r]), y23, acc[r]);
#pragma unroll
                for (int r = 0; r < Q; ++r) acc[r] = fma2(pack2(q2x[r], q2x[r]), x23, acc[r]);
#pragma unroll
                for (int r = 0; r < Q; ++r) {
                    float a0, a1;
                    unpack2(acc[r], a0, a1);
                    tm[r] = min3(tm[r], a0, a1);
                }
            }
            unsigned flags = 0;
#pragma unroll
            for (int r = 0; r < Q; ++r) flags |= (tm[r] <= thr[r]) ? (1u << r) : 0u;
            if (flags != 0) {
#ifdef ISR_NN_TUNING
                if (p.dbg != nullptr) {
                    const unsigned am = __activemask();
                    if ((int)(__ffs(am) - 1) == (tid & 31)) atomicAdd(p.dbg + 0, 1ull);  // warp-level flagged sub-tiles
                    atomicAdd(p.dbg + 1, (unsigned long long)__popc(flags));            // (query, sub-tile) events
                }
#endif
                // ---- resolve: rare, divergent; one pass serves every flagged lane ----------
#pragma unroll
                for (int r = 0; r < Q; ++r) tm_l[r] = tm[r];
                const float *fx = reinterpret_cast<const float *>(sx) + sub * SUB;
                const float *fy = reinterpret_cast<const float *>(sy) + sub * SUB;
                const float *fz = reinterpret_cast<const float *>(sz) + sub * SUB;
                const float *fn = reinterpret_cast<const float *>(sn) + sub * SUB;
                const int gbase = (s_begin + stage_of(sl)) * STAGE + sub * SUB;
                for (unsigned f = flags; f != 0; f &= f - 1) {
                    const int r = __ffs(f) - 1;
#ifdef ISR_NN_TUNING
                    if (p.dbg != nullptr) {
                        const unsigned am = __activemask();
                        if ((int)(__ffs(am) - 1) == (tid & 31)) atomicAdd(p.dbg + 2, 1ull);  // warp-level resolve passes
                    }
#endif
                    const int qi = min(q0 + r * 32, p.nq_pad - 1);
                    const float qhx = gq[qi], qhy = gq[p.nq_pad + qi], qhz = gq[2ll * p.nq_pad + qi];
                    const float cx = -2.0f * qhx, cy = -2.0f * qhy, cz = -2.0f * qhz;
                    const float nq2 = __fmaf_rn(qhz, qhz, __fmaf_rn(qhy, qhy, qhx * qhx));
                    const float m = fminf(mt_l[r], tm_l[r]);
                    const float th = filter_threshold(m, nq2, sqrtf(nq2));
                    mt_l[r] = m;
                    thr_l[r] = th;
                    double qlx = 0.0, qly = 0.0, qlz = 0.0;
                    if (p.use_lo) {
                        qlx = gq[4ll * p.nq_pad + qi];
                        qly = gq[5ll * p.nq_pad + qi];
                        qlz = gq[6ll * p.nq_pad + qi];
                    }
                    double Db = Dbest_l[r];
                    int ib = ibest_l[r];
#pragma unroll 4
                    for (int j = 0; j < SUB; ++j) {
                        const float px = fx[j], py = fy[j], pz = fz[j];
                        const float a = __fmaf_rn(cx, px, __fmaf_rn(cy, py, __fmaf_rn(cz, pz, fn[j])));
                        if (a <= th) {
#ifdef ISR_NN_TUNING
                            if (p.dbg != nullptr) atomicAdd(p.dbg + 3, 1ull);  // exact evaluations
#endif
                            double dx = (double)qhx - (double)px, dy = (double)qhy - (double)py,
                                   dz = (double)qhz - (double)pz;
                            if (p.use_lo) {
                                const long long gj = gbase + j;
                                dx += qlx - (double)gt[4ll * p.nt_pad + gj];
                                dy += qly - (double)gt[5ll * p.nt_pad + gj];
                                dz += qlz - (double)gt[6ll * p.nt_pad + gj];
                            }
                            const double D = fma(dz, dz, fma(dy, dy, dx * dx));
                            // strict minimum; on an exact tie the lower ORIGINAL index wins
                            // (stages are not visited in index order and storage is permuted)
                            bool take = D < Db;
                            if (D == Db) {
                                const int cand = gbase + j;
                                const int oc = p.perm_t != nullptr ? p.perm_t[min(cand, p.nt - 1)] : cand;
                                const int ob = p.perm_t != nullptr ? p.perm_t[min(ib, p.nt - 1)] : ib;
                                take = oc < ob;
                            }
                            if (take) {
                                Db = D;
                                ib = gbase + j;
                            }
                        }
                    }
                    Dbest_l[r] = Db;
                    ibest_l[r] = ib;
                    // >= the exact best distance of the FP64 (hi + lo) query, rounded up
                    if (PRUNE)
                        dq_l[r] = __double2float_ru(sqrt(Db)) * 1.00002f + 1e-6f * sqrtf(nq2) + 1e-37f;
                }
#pragma unroll
                for (int r = 0; r < Q; ++r) thr[r] = thr_l[r];
                if (PRUNE) {
                    float m = 0.f;
                    for (int r = 0; r < Q; ++r) m = fmaxf(m, dq_l[r]);  // run-time loop on purpose
                    dmax = m;
                }
            }
        }
        __syncthreads();  // every warp is done with this slot before it is refilled
        if (PRUNE && sl == 0 && nst > 1) {
            // every live query now holds a finite bound: list the stages that can still
            // matter to any query of the block, |cQ - c| <= B + rQ + r.
            float B = dmax;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) B = fmaxf(B, __shfl_xor_sync(0xffffffffu, B, o));
            if (lane == 0) cred[1][warp] = B;
            __syncthreads();
            B = cred[1][0];
#pragma unroll
            for (int w = 1; w < WARPS; ++w) B = fmaxf(B, cred[1][w]);
            const float4 *sc = p.stage_c + (long long)b * p.stage_c_bstride + s_begin;
            for (int base = 0; base < nst; base += THREADS) {
                const int sI = base + tid;
                bool need = false;
                if (sI < nst && sI != s_first) {
                    const float4 c = sc[sI];
                    const float dx = c.x - cQx, dy = c.y - cQy, dz = c.z - cQz;
                    const float rr = (B + rQ + c.w) * 1.0001f;
                    need = c.w >= 0.f && !(fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > rr * rr);
                }
                const unsigned m = __ballot_sync(0xffffffffu, need);
                int at = 0;
                if (lane == 0 && m != 0) at = atomicAdd(&scount, __popc(m));
                at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
                if (need && at < kListCap) slist[at] = sI;
            }
            __syncthreads();
            const int cnt = scount;
            use_list = cnt <= kListCap;
            npos = use_list ? 1 + cnt : nst;
            if (tid == 0) {
                for (int i = 1; i < NSTAGES - 1 + 1 && i < npos; ++i) issue(i);
            }
        }
    }
    if (p.evaluated != nullptr && lane == 0 && nscanned != 0)
        atomicAdd(p.evaluated, (unsigned long long)nscanned);

    for (int r = 0; r < Q; ++r) {
        const int i = q0 + r * 32;
        if (i < p.nq) {
            // report in the caller's original indexing
            const int io = p.perm_q != nullptr ? p.perm_q[i] : i;
            const int jo = p.perm_t != nullptr ? p.perm_t[min(ibest_l[r], p.nt - 1)] : ibest_l[r];
            const long long o = (long long)b * p.nq + io;
            if (p.part_D != nullptr) {
                p.part_D[(long long)blockIdx.y * p.part_stride + o] = Dbest_l[r];
                p.part_idx[(long long)blockIdx.y * p.part_stride + o] = jo;
            } else {
                p.out_d2[o] = (float)Dbest_l[r];
                if (p.out_idx != nullptr) p.out_idx[o] = jo;
            }
        }
    }
}

// partial results carry original indices; exact ties go to the lower one
__global__ void nn2_combine_kernel(const double *__restrict__ part_D, const int *__restrict__ part_idx,
                                   long long total, int splits, float *__restrict__ out_d2,
                                   int *__restrict__ out_idx, const int *__restrict__ skip,
                                   long long skip_stride, long long per_batch) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    if (skip != nullptr && skip[(i / per_batch) * skip_stride] != 0) return;
    double best = part_D[i];
    int bi = part_idx[i];
    for (int s = 1; s < splits; ++s) {
        const double d = part_D[(long long)s * total + i];
        const int di = part_idx[(long long)s * total + i];
        if (d < best || (d == best && di < bi)) {
            best = d;
            bi = di;
        }
    }
    out_d2[i] = (float)best;
    if (out_idx != nullptr) out_idx[i] = bi;
}

// ---- host-side launch ----------------------------------------------------------------
template <int Q, int THREADS, int STAGE, int NSTAGES, int SUB, int MINB, int UNR = 2, bool PRUNE = false>
struct NN2Variant {
    static constexpr int kQueriesPerCta = Q * THREADS;
    static constexpr int kStage = STAGE;
    static constexpr bool kPrune = PRUNE;
    static constexpr size_t kSmem =
        (size_t)NSTAGES * (4 * STAGE + (PRUNE ? 4 * (STAGE / SUB) : 0)) * 4 + NSTAGES * 8 +
        (PRUNE ? kListCap * 4 : 0);

    static int launch(const NN2Params &p, dim3 grid, cudaStream_t st) {
        auto kern = nn2_kernel<Q, THREADS, STAGE, NSTAGES, SUB, MINB, UNR, PRUNE>;
        static thread_local int configured_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_dev != dev) {
            ISR_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)kSmem),
                               "nn2 smem attr"));
            configured_dev = dev;
        }
        ProfScope prof(kProfNN, st);
        kern<<<grid, THREADS, kSmem, st>>>(p);
        return launched("nn2_kernel");
    }

    static int ctas_per_sm() {
        int n = 0;
        auto kern = nn2_kernel<Q, THREADS, STAGE, NSTAGES, SUB, MINB, UNR, PRUNE>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, THREADS, kSmem) != cudaSuccess ||
            n < 1)
            n = 1;
        return n;
    }
};

using NN2Main = NN2Variant<8, 128, 1024, 3, 64, 4, 1>;
using NN2Pruned = NN2Variant<8, 128, 1024, 3, 64, 4, 1, true>;
constexpr int kMaxSplits = 32;

// process-wide pruning switch and profiling counters (isr.h)
static std::atomic<int> g_prune{-1};
static std::atomic<unsigned long long> g_answered{0};
static unsigned long long *g_evaluated_dev[64] = {nullptr};  // per device, allocated on first use

static bool pruning_on() {
    int v = g_prune.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("ISR_NN_PRUNE");
        v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
        g_prune.store(v);
    }
    return v != 0;
}

static unsigned long long *evaluated_counter() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (g_evaluated_dev[dev] == nullptr) {
        unsigned long long *ptr = nullptr;
        if (cudaMalloc(&ptr, 64) != cudaSuccess) return nullptr;
        cudaMemset(ptr, 0, 64);
        g_evaluated_dev[dev] = ptr;
    }
    return g_evaluated_dev[dev];
}
constexpr int kSlotsUpperBound = 148 * 8;  // for workspace sizing without a device

static int choose_splits(long long ctas, int stages, int slots) {
    int splits = 1;
    if (ctas < 6ll * slots) {
        long long want = (8ll * slots + ctas - 1) / ctas;
        int max_splits = stages / 4 > 0 ? stages / 4 : 1;  // >= 4 stages per split
        if (max_splits > kMaxSplits) max_splits = kMaxSplits;
        splits = (int)(want < max_splits ? want : max_splits);
        if (splits < 1) splits = 1;
    }
    return splits;
}

struct NN2Call {
    const IsrCloud *q;
    const IsrCloud *t;
    int64_t batch;
    int use_lo;
    float *out_d2; int32_t *out_idx;
    const int32_t *skip; int64_t skip_stride;
    void *workspace; size_t workspace_bytes;
    cudaStream_t st;
};

template <class V>
static int nn2_dispatch(const NN2Call &c) {
    const int64_t nq = c.q->n;
    const int nqb = (int)((nq + V::kQueriesPerCta - 1) / V::kQueriesPerCta);
    const int stages = (int)(c.t->npad / V::kStage);
    static thread_local int slots = 0;
    if (slots == 0) slots = sm_count() * V::ctas_per_sm();
    // a pruning CTA needs the whole target range: its bound comes from the nearest stage
    int splits = V::kPrune ? 1 : choose_splits((long long)nqb * c.batch, stages, slots);
    const int per = (stages + splits - 1) / splits;
    splits = (stages + per - 1) / per;  // no empty split

    NN2Params p;
    p.q = c.q->soa7; p.q_bstride = c.q->bstride; p.nq = (int)nq; p.nq_pad = (int)c.q->npad;
    p.t = c.t->soa7; p.t_bstride = c.t->bstride; p.nt_pad = (int)c.t->npad; p.nt = (int)c.t->n;
    p.out_d2 = c.out_d2; p.out_idx = c.out_idx;
    // stage centroids are laid out per 1024-point SoA tile; only usable when the kernel's
    // stage is that tile
    p.stage_c = (V::kStage == ISR_SOA_TILE) ? reinterpret_cast<const float4 *>(c.t->stage_c) : nullptr;
    p.stage_c_bstride = c.t->bstride == 0 ? 0 : c.t->npad / ISR_SOA_TILE;
    p.perm_q = c.q->perm; p.perm_t = c.t->perm;
    p.sub_c = reinterpret_cast<const float4 *>(c.t->sub_c);
    p.sub_c_bstride = c.t->bstride == 0 ? 0 : c.t->npad / ISR_SUB_TILE;
    p.evaluated = nullptr;
    if (prof_enabled()) {
        p.evaluated = evaluated_counter();
        g_answered.fetch_add((unsigned long long)nq * (unsigned long long)c.t->n *
                             (unsigned long long)c.batch);
    }
    p.dbg = nullptr;
#ifdef ISR_NN_TUNING
    {
        static unsigned long long *dbg = nullptr;
        const char *e = getenv("ISR_NN2_DEBUG");
        if (e && atoi(e)) {
            if (dbg == nullptr) { cudaMalloc(&dbg, 64); cudaMemset(dbg, 0, 64); }
            else {
                unsigned long long h[4];
                cudaMemcpy(h, dbg, 32, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[nn2 dbg, previous launch] warp-flagged-subtiles=%llu query-events=%llu warp-resolve-passes=%llu exact-evals=%llu\n", h[0], h[1], h[2], h[3]);
                cudaMemset(dbg, 0, 64);
            }
            p.dbg = dbg;
        }
    }
#endif
    p.part_D = nullptr; p.part_idx = nullptr; p.part_stride = 0;
    p.skip = c.skip; p.skip_stride = c.skip_stride;
    p.stages_total = stages; p.stages_per_split = per;
    p.use_lo = c.use_lo;
    {
        static int nores = -1;
        if (nores < 0) { const char *e = getenv("ISR_NN2_NORESOLVE"); nores = (e && atoi(e)) ? 1 : 0; }
        p.debug_no_resolve = nores;
    }
    dim3 grid((unsigned)nqb, (unsigned)splits, (unsigned)c.batch);
    if (splits == 1) return V::launch(p, grid, c.st);

    const long long total = (long long)nq * c.batch;
    const size_t need = align256((size_t)total * splits * 8) + (size_t)total * splits * 4;
    ISR_REQUIRE(c.workspace != nullptr && c.workspace_bytes >= need, ISR_E_WORKSPACE,
                "nn: workspace %zu < %zu bytes", c.workspace_bytes, need);
    p.part_D = reinterpret_cast<double *>(c.workspace);
    p.part_idx = reinterpret_cast<int *>(reinterpret_cast<char *>(c.workspace) +
                                         align256((size_t)total * splits * 8));
    p.part_stride = total;
    ISR_TRY(V::launch(p, grid, c.st));
    nn2_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.st>>>(
        p.part_D, p.part_idx, total, splits, c.out_d2, c.out_idx, c.skip, c.skip_stride,
        (long long)nq);
    return launched("nn2_combine_kernel");
}

#ifdef ISR_NN_TUNING
using NN2T13 = NN2Variant<8, 128, 1024, 3, 64, 4, 1>;
using NN2T14 = NN2Variant<8, 128, 1024, 3, 64, 4, 4>;
using NN2T15 = NN2Variant<8, 128, 1024, 3, 128, 4, 1>;
using NN2T16 = NN2Variant<8, 128, 1024, 4, 64, 4, 1>;
using NN2T1 = NN2Variant<8, 128, 1024, 3, 16, 4>;
using NN2T2 = NN2Variant<8, 128, 1024, 3, 64, 4>;
using NN2T3 = NN2Variant<8, 128, 1024, 3, 32, 4, 4>;
using NN2T4 = NN2Variant<8, 128, 1024, 3, 32, 4, 1>;
using NN2T5 = NN2Variant<4, 128, 1024, 3, 32, 6>;
using NN2T6 = NN2Variant<8, 256, 1024, 3, 32, 2>;
using NN2T7 = NN2Variant<12, 128, 1024, 3, 32, 3>;
using NN2T8 = NN2Variant<16, 64, 1024, 3, 32, 4>;
using NN2T9 = NN2Variant<8, 128, 512, 4, 32, 4>;
using NN2T10 = NN2Variant<8, 128, 1024, 3, 32, 3, 4>;
using NN2T11 = NN2Variant<6, 128, 1024, 3, 32, 5>;
using NN2T12 = NN2Variant<8, 128, 1024, 3, 128, 4>;
#endif

}  // namespace isr

extern "C" {

size_t isr_nn2_workspace_bytes(int64_t nq, int64_t nt, int64_t batch) {
    using namespace isr;
    if (nq <= 0 || batch <= 0 || nt <= 0) return 256;
    const int64_t nt_pad = isr_soa_padded_len(nt);
    const int nqb = (int)((nq + 2047) / 2048);  // largest CTA tile of any variant
    const int splits = choose_splits((long long)nqb * batch, (int)(nt_pad / 512), kSlotsUpperBound);
    if (splits <= 1) return 256;
    const size_t total = (size_t)nq * (size_t)batch;
    return align256(total * splits * 8) + align256(total * splits * 4);
}

int isr_nn2(const IsrCloud *q, const IsrCloud *t, int64_t batch, int use_lo, float *out_d2,
            int32_t *out_idx, const int32_t *skip, int64_t skip_stride, void *workspace,
            size_t workspace_bytes, void *stream) {
    using namespace isr;
    ISR_REQUIRE(q != nullptr && t != nullptr, ISR_E_INVALID_ARG, "nn: null cloud descriptor");
    ISR_REQUIRE(q->n >= 0 && t->n >= 1 && batch >= 0, ISR_E_SHAPE,
                "nn: need nq >= 0, nt >= 1, batch >= 0 (nq=%lld nt=%lld batch=%lld)",
                (long long)q->n, (long long)t->n, (long long)batch);
    if (q->n == 0 || batch == 0) return ISR_OK;
    ISR_REQUIRE(q->soa7 && t->soa7 && out_d2, ISR_E_INVALID_ARG, "nn: null pointer");
    ISR_REQUIRE(q->npad >= q->n && q->npad % ISR_SOA_TILE == 0 && t->npad >= t->n &&
                    t->npad % ISR_SOA_TILE == 0,
                ISR_E_SHAPE, "nn: padded lengths must be multiples of %d covering n", ISR_SOA_TILE);
    ISR_REQUIRE(q->npad < (1ll << 31) - 2048 && t->npad < (1ll << 31) - 2048, ISR_E_SHAPE,
                "nn: clouds beyond int32 indexing");
    ISR_REQUIRE(batch <= 65535, ISR_E_SHAPE, "nn: batch %lld > 65535", (long long)batch);
    ISR_REQUIRE(aligned16(t->soa7) && (t->bstride % 4 == 0), ISR_E_ALIGN,
                "nn: target planes must be 16-byte aligned");
    ISR_REQUIRE(t->stage_c == nullptr || aligned16(t->stage_c), ISR_E_ALIGN,
                "nn: stage centroids must be 16-byte aligned");
    const NN2Call c{q, t, batch, use_lo, out_d2, out_idx, skip, skip_stride, workspace,
                    workspace_bytes, (cudaStream_t)stream};
#ifdef ISR_NN_TUNING
    static int variant = -1;
    if (variant < 0) {
        const char *e = getenv("ISR_NN_VARIANT");
        variant = e ? atoi(e) : 0;
    }
    switch (variant) {
        case 1: return nn2_dispatch<NN2T1>(c);
        case 2: return nn2_dispatch<NN2T2>(c);
        case 3: return nn2_dispatch<NN2T3>(c);
        case 4: return nn2_dispatch<NN2T4>(c);
        case 5: return nn2_dispatch<NN2T5>(c);
        case 6: return nn2_dispatch<NN2T6>(c);
        case 7: return nn2_dispatch<NN2T7>(c);
        case 8: return nn2_dispatch<NN2T8>(c);
        case 9: return nn2_dispatch<NN2T9>(c);
        case 10: return nn2_dispatch<NN2T10>(c);
        case 11: return nn2_dispatch<NN2T11>(c);
        case 12: return nn2_dispatch<NN2T12>(c);
        case 13: return nn2_dispatch<NN2T13>(c);
        case 14: return nn2_dispatch<NN2T14>(c);
        case 15: return nn2_dispatch<NN2T15>(c);
        case 16: return nn2_dispatch<NN2T16>(c);
        default: break;
    }
#endif
    if (t->sub_c != nullptr && t->stage_c != nullptr && pruning_on()) {
        ISR_REQUIRE(aligned16(t->sub_c), ISR_E_ALIGN, "nn: sub-tile spheres must be 16-byte aligned");
        return nn2_dispatch<NN2Pruned>(c);
    }
    return nn2_dispatch<NN2Main>(c);
}

int isr_set_nn_pruning(int on) {
    isr::g_prune.store(on != 0 ? 1 : 0);
    return ISR_OK;
}

int isr_get_nn_pruning(void) { return isr::pruning_on() ? 1 : 0; }

int isr_profile_nn_pairs(uint64_t *evaluated_host, uint64_t *answered_host) {
    using namespace isr;
    unsigned long long units = 0;
    unsigned long long *ctr = evaluated_counter();
    if (ctr != nullptr) {
        ISR_TRY(check_cuda(cudaDeviceSynchronize(), "profile_nn_pairs sync"));
        ISR_TRY(check_cuda(cudaMemcpy(&units, ctr, 8, cudaMemcpyDeviceToHost), "profile_nn_pairs read"));
        ISR_TRY(check_cuda(cudaMemset(ctr, 0, 8), "profile_nn_pairs clear"));
    }
    // one unit = one warp block of 32 x 8 queries against one 64-target sub-tile
    if (evaluated_host) *evaluated_host = (uint64_t)units * 256ull * (uint64_t)ISR_SUB_TILE;
    if (answered_host) *answered_host = (uint64_t)g_answered.exchange(0);
    return ISR_OK;
}

}  // extern "C"
