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
//   flag    per 64-target sub-tile one compare per query: does the sub-tile minimum come
//           within the window W of the running minimum?
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
// Pruning (nn2_pruned_kernel).  Clouds are stored along a Hilbert curve (sort.cu), so a warp's
// 256 queries, a 1024-target stage and a 64-target sub-tile are all compact patches.  Every
// query carries dq >= its exact best distance so far (FP64 Dbest from the resolve path, rounded
// up, plus the size of its lo part).  A tile with sphere (c, r) is skipped for a query iff
// |q - c| > dq + r, a sub-tile also iff dist(q, box) > dq for its axis-aligned box (both in
// FP32 with a 1e-4 relative margin; radius and half-extents are inflated by prepare.cu for
// rounding and for the targets' lo parts): every point of the tile is then strictly farther
// than the neighbour already held, so it can be neither the minimum nor an equal-distance
// tie.  Every warp works on its own (one warp per CTA, no barrier, its own shared-memory ring
// fed by 1-D bulk copies):
//   (1) starting bounds: the caller's hints (ICP: the previous iteration's neighbours), else
//       seeds -- for the centre of each of its 8 query rows the sub-tile whose centre is
//       nearest (via the nearest stage), scanned first;
//   (2) a three-level walk over the target's spheres, coarse tests against the 8 query-ROW
//       spheres first (one sphere per lane): chunk spheres (32 stages each) -> stage spheres
//       -> per candidate stage the exact per-query test, then its 16 sub-tile spheres / boxes
//       -> FIFO entries with the mask of rows they may matter to;
//   (3) when everything is queued, the entries are put in nearest-first order;
//   (4) each entry gets the exact per-query test with the bounds of that moment (which also
//       yields the rows that still need it); a survivor's 1 KB bulk copy is issued at once, up
//       to three entries ahead of the scan, and it is scanned for the quarters (two rows each)
//       that need it.
// The exactness argument above is untouched: the true neighbour's tile is never skipped
// (its distance is <= every bound), so it is visited, flagged and resolved as before.
//
// Roofline: FP32 CUDA cores.  Algorithmic work stays 8 flop per pair;
// executed FP32-pipe work is 3 lane-ops per pair, so the algorithmic rate can exceed the
// nominal FMA peak (cap 8/6) -- bench.py reports both.
#include <math_constants.h>
#include <stdlib.h>

#include "icp_device.cuh"
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
    const int *order;        // [batch][gridDim.x] launch-list entries (launch_entry()) run by CTA x, or NULL
    const int *order_count;  // [batch] entries of `order` in use
    int *hint;               // [batch][nq_pad] in/out starting neighbours (stored positions), or NULL
    int nanchors;            // seeds used (<= kAnchors; fewer only for tuning runs)
    int sort_fifo;           // scan the queued sub-tiles nearest-first (0: in stage order)
    const unsigned *sub_h;   // [batch][stages_total * STAGE/SUB] packed half-extents of the sub-tile
    long long sub_h_bstride; //   boxes (isr_tile_spheres), or NULL: sphere tests only
    // target parts (fused ICP, shallow grids): the heaviest single-row entries of the launch list run
    // as kTargetParts CTAs that share the row's queries and own every kTargetParts-th target stage;
    // the last part to finish merges the (FP64 distance, index) pairs and carries on with the row
    const int *order_slot;   // [gridDim.x] merge slot of a launch-list entry with a part code
    double *tp_D;            // [slots][kTargetParts][32]
    int *tp_I;               // [slots][kTargetParts][32]
    unsigned *tp_tick;       // [slots], zero between launches
    unsigned *cost;          // [gridDim.x] out: SM cycles of this launch-list entry's search (fused ICP,
                             //   single start: block_rebalance_kernel re-cuts the launch list from them), or NULL
    unsigned long long *cta_log;  // developer probe (isr_debug_cta_log): 4 words per CTA, or NULL
    long long cta_log_cap;        //   records that fit
};

#ifdef ISR_PHASE_LOG
// developer build (scripts/probe_phases.py): the CTA log carries phase stamps instead of event counts,
// and its LAST record the launch's wall-clock marks (ns): first CTA start, last search end, last
// epilogue end, kernel end (after the reduction tail, the exchange and the solve)
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif
constexpr int kTargetParts = 4;   // CTAs that share the target of one heavy query row (a power of two <= 4)
constexpr int kTargetSlots = 96;  // rows per launch that can run that way
constexpr unsigned kNoBox = 0x3FFFFFFFu;  // three 10-bit fractions of the radius, all ones

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

// The filter scan of SUB targets for query rows R0 .. R0 + QN - 1 of the lane.  The sub-tile is
// walked in PARTS equal pieces; after each piece the rows whose minimum over the piece comes
// within their threshold get bit (piece * 8 + row) of `flags` (the resolve then re-derives only
// those pieces); tm[R0 ..] receives the minimum over the whole sub-tile.  Level-major over the
// rows: consecutive FFMA2 are independent and share the target-pair operand (operand-reuse
// cache); the query rides as a scalar.
template <int Q, int R0, int QN, int SUB, int UNR, int PARTS>
__device__ __forceinline__ void scan_rows(const float4 *sx, const float4 *sy, const float4 *sz,
                                          const float4 *sn, const float (&q2x)[Q], const float (&q2y)[Q],
                                          const float (&q2z)[Q], const float (&thr)[Q], float (&tm)[Q],
                                          unsigned &flags) {
    constexpr int G = SUB / 4 / PARTS;  // groups of 4 targets per piece
    static_assert(G * PARTS * 4 == SUB && PARTS <= 4, "pieces");
#pragma unroll 1
    for (int part = 0; part < PARTS; ++part) {
        float tp[QN];
#pragma unroll
        for (int r = 0; r < QN; ++r) tp[r] = CUDART_INF_F;
#pragma unroll UNR
        for (int g = part * G; g < (part + 1) * G; ++g) {
            const float4 X = sx[g];
            const float4 Y = sy[g];
            const float4 Z = sz[g];
            const float4 N = sn[g];
            const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
            const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
            const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
            const u64 n01 = pack2(N.x, N.y), n23 = pack2(N.z, N.w);
            u64 acc[QN];
#pragma unroll
            for (int r = 0; r < QN; ++r) acc[r] = fma2(pack2(q2z[R0 + r], q2z[R0 + r]), z01, n01);
#pragma unroll
            for (int r = 0; r < QN; ++r) acc[r] = fma2(pack2(q2y[R0 + r], q2y[R0 + r]), y01, acc[r]);
#pragma unroll
            for (int r = 0; r < QN; ++r) acc[r] = fma2(pack2(q2x[R0 + r], q2x[R0 + r]), x01, acc[r]);
#pragma unroll
            for (int r = 0; r < QN; ++r) {
                float a0, a1;
                unpack2(acc[r], a0, a1);
                tp[r] = min3(tp[r], a0, a1);
            }
#pragma unroll
            for (int r = 0; r < QN; ++r) acc[r] = fma2(pack2(q2z[R0 + r], q2z[R0 + r]), z23, n23);
#pragma unroll
            for (int r = 0; r < QN; ++r) acc[r] = fma2(pack2(q2y[R0 + r], q2y[R0 + r]), y23, acc[r]);
#pragma unroll
            for (int r = 0; r < QN; ++r) acc[r] = fma2(pack2(q2x[R0 + r], q2x[R0 + r]), x23, acc[r]);
#pragma unroll
            for (int r = 0; r < QN; ++r) {
                float a0, a1;
                unpack2(acc[r], a0, a1);
                tp[r] = min3(tp[r], a0, a1);
            }
        }
        unsigned f = 0;
#pragma unroll
        for (int r = 0; r < QN; ++r) {
            f |= (tp[r] <= thr[R0 + r]) ? (1u << (R0 + r)) : 0u;
            tm[R0 + r] = PARTS == 1 ? tp[r] : fminf(tm[R0 + r], tp[r]);
        }
        flags |= f << (8 * part);
    }
}

// The FP64 resolve of one scanned (warp, sub-tile) unit.  pflags: bit (piece * 8 + row) set
// where the filter minimum of that piece came within the row's threshold; tm: the filter
// minima over the whole sub-tile.  sx..sn point at the sub-tile's x, y, z, |p|^2 values in
// shared memory; gbase is the stored index of its first target.  The resolve state (mt_l ..
// ibest_l) is indexed at run time, which places it in local memory and keeps it
// out of the scan's registers.  Returns whether anything was flagged.
template <int Q, int SUB, bool PRUNE, int PARTS>
__device__ __forceinline__ bool resolve_flagged(const NN2Params &p, const float *__restrict__ gq,
                                                const float *__restrict__ gt, int q0, int lane,
                                                const float4 *sx, const float4 *sy, const float4 *sz,
                                                const float4 *sn, int gbase, unsigned pflags,
                                                const float (&tm)[Q], float *mt_l, float *thr_l,
                                                float *tm_l, double *Dbest_l, int *ibest_l,
                                                float *dq_l, float &dmax, unsigned &nflag,
                                                unsigned &npass, const float *qsm, const float *qlo) {
    // bit r: row r has a flagged piece
    const unsigned flags = (pflags | (pflags >> 8) | (pflags >> 16) | (pflags >> 24)) & 0xFFu;
    if (PRUNE && (p.evaluated != nullptr || p.cta_log != nullptr)) {  // profiling only
        const unsigned mx = __reduce_max_sync(0xffffffffu, (unsigned)__popc(flags));
        nflag += mx != 0 ? 1u : 0u;
        npass += mx;
    }
    if (flags != 0) {
#ifdef ISR_NN_TUNING
        if (p.dbg != nullptr) {
            const unsigned am = __activemask();
            if ((int)(__ffs(am) - 1) == lane) atomicAdd(p.dbg + 0, 1ull);  // warp-level flagged sub-tiles
            atomicAdd(p.dbg + 1, (unsigned long long)__popc(flags));            // (query, sub-tile) events
        }
#endif
        // ---- resolve: rare, divergent; one pass serves every flagged lane ----------
#pragma unroll
        for (int r = 0; r < Q; ++r) tm_l[r] = tm[r];
        for (unsigned f = flags; f != 0; f &= f - 1) {
            const int r = __ffs(f) - 1;
#ifdef ISR_NN_TUNING
            if (p.dbg != nullptr) {
                const unsigned am = __activemask();
                if ((int)(__ffs(am) - 1) == lane) atomicAdd(p.dbg + 2, 1ull);  // warp-level resolve passes
            }
#endif
            const int qi = min(q0 + r * 32, p.nq_pad - 1);
            // the query's coordinates: from the warp's shared-memory copy when there is one
            // (qsm: [3][32 * Q], hi xyz; the lo parts sit in the lane's local array qlo), else from global memory
            const int qs = r * 32 + lane;
            const float qhx = PRUNE ? qsm[qs] : gq[qi], qhy = PRUNE ? qsm[32 * Q + qs] : gq[p.nq_pad + qi],
                        qhz = PRUNE ? qsm[2 * 32 * Q + qs] : gq[2ll * p.nq_pad + qi];
            const float cx = -2.0f * qhx, cy = -2.0f * qhy, cz = -2.0f * qhz;
            const float nq2 = __fmaf_rn(qhz, qhz, __fmaf_rn(qhy, qhy, qhx * qhx));
            const float m = fminf(mt_l[r], tm_l[r]);
            const float th = filter_threshold(m, nq2, sqrtf(nq2));
            mt_l[r] = m;
            thr_l[r] = th;
            double qlx = 0.0, qly = 0.0, qlz = 0.0;
            if (p.use_lo) {
                qlx = PRUNE ? qlo[r] : gq[4ll * p.nq_pad + qi];
                qly = PRUNE ? qlo[Q + r] : gq[5ll * p.nq_pad + qi];
                qlz = PRUNE ? qlo[2 * Q + r] : gq[6ll * p.nq_pad + qi];
            }
            double Db = Dbest_l[r];
            int ib = ibest_l[r];
            // pass 1, branch-free: which targets fall inside the window?  (Lanes hit in
            // different places; a branch here would run the FP64 code once per hit position.)
            static_assert(SUB <= 64, "window mask is 64 bits");
            u64 wmask = 0;
            // (lanes walk their own flagged pieces side by side: same code, different offsets)
            const u64 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
            // Packed: the same three roundings per target as the scan, two targets per FFMA2 (the query
            // rides as a scalar operand), and the mask of 16 targets assembled with compile-time bit
            // positions -- one 64-bit shift per 16 targets.  (The scalar form -- 12 FFMA and a 64-bit
            // shift per group of four targets -- was a quarter of the verification kernel's instructions:
            // 12.3k -> 13.3k candidates/s in the quick probe, ICP 0.484 -> 0.472 ms per iteration.)
            for (unsigned pf = (pflags >> r) & 0x01010101u; pf != 0; pf &= pf - 1) {
                constexpr int G = SUB / 4 / PARTS;
                static_assert(G % 4 == 0, "batches of four groups");
                const int g0 = ((__ffs(pf) - 1) >> 3) * G;
#pragma unroll 1
                for (int gb = g0; gb < g0 + G; gb += 4) {
                    unsigned pm = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 X = sx[gb + k], Y = sy[gb + k], Z = sz[gb + k], N = sn[gb + k];
                        const u64 a01 = fma2(cx2, pack2(X.x, X.y),
                                             fma2(cy2, pack2(Y.x, Y.y), fma2(cz2, pack2(Z.x, Z.y), pack2(N.x, N.y))));
                        const u64 a23 = fma2(cx2, pack2(X.z, X.w),
                                             fma2(cy2, pack2(Y.z, Y.w), fma2(cz2, pack2(Z.z, Z.w), pack2(N.z, N.w))));
                        float a0, a1, a2, a3;
                        unpack2(a01, a0, a1);
                        unpack2(a23, a2, a3);
                        pm |= (a0 <= th ? 1u << (4 * k) : 0u) | (a1 <= th ? 2u << (4 * k) : 0u) |
                              (a2 <= th ? 4u << (4 * k) : 0u) | (a3 <= th ? 8u << (4 * k) : 0u);
                    }
                    wmask |= (u64)pm << (4 * gb);
                }
            }
            // pass 2: exact FP64 distance of those few (ascending stored position)
            const float *fx = reinterpret_cast<const float *>(sx);
            const float *fy = reinterpret_cast<const float *>(sy);
            const float *fz = reinterpret_cast<const float *>(sz);
            for (; wmask != 0; wmask &= wmask - 1) {
                const int j = __ffsll((long long)wmask) - 1;
                // the point already held (a hinted search meets its own hint again in every
                // iteration): same arithmetic, same D -- nothing to decide, and no trip to the lo planes
                if (gbase + j == ib && Db < CUDART_INF) continue;  // (ib is a placeholder until Db is finite)
                const float px = fx[j], py = fy[j], pz = fz[j];
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
                // (tiles are not visited in index order and storage is permuted)
                bool take = D < Db;
                // (a hinted search meets its own hint again -- the same point, not a tie to break:
                // the two dependent perm loads below would be paid by every query of every iteration)
                if (D == Db && gbase + j != ib) {
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
            Dbest_l[r] = Db;
            ibest_l[r] = ib;
            // >= the exact best distance of the FP64 (hi + lo) query, rounded up
            if (PRUNE)
                dq_l[r * 32] = __fsqrt_ru(__double2float_ru(Db)) * 1.00002f + 1e-6f * sqrtf(nq2) + 1e-37f;
        }
        if (PRUNE) {
            float m = 0.f;
            for (int r = 0; r < Q; ++r) m = fmaxf(m, dq_l[r * 32]);  // run-time loop on purpose
            dmax = m;
        }
    }
    return flags != 0;
}

// One (warp, sub-tile) unit of the exhaustive kernel: the FP32 filter scan of SUB targets
// against the warp's 32 x Q queries (operands and thresholds in registers), then the resolve.
template <int Q, int SUB, int UNR>
__device__ __forceinline__ void scan_subtile(const NN2Params &p, const float *__restrict__ gq,
                                             const float *__restrict__ gt, int q0, int lane,
                                             const float4 *sx, const float4 *sy, const float4 *sz,
                                             const float4 *sn, int gbase, const float (&q2x)[Q],
                                             const float (&q2y)[Q], const float (&q2z)[Q],
                                             float (&thr)[Q], float *mt_l, float *thr_l,
                                             float *tm_l, double *Dbest_l, int *ibest_l,
                                             float *dq_l, float &dmax, unsigned &nflag,
                                             unsigned &npass) {
    float tm[Q];
#pragma unroll
    for (int r = 0; r < Q; ++r) tm[r] = CUDART_INF_F;
    unsigned pflags = 0;
    scan_rows<Q, 0, Q, SUB, UNR, 1>(sx, sy, sz, sn, q2x, q2y, q2z, thr, tm, pflags);
    if (resolve_flagged<Q, SUB, false, 1>(p, gq, gt, q0, lane, sx, sy, sz, sn, gbase, pflags, tm, mt_l, thr_l,
                                          tm_l, Dbest_l, ibest_l, dq_l, dmax, nflag, npass, nullptr, nullptr)) {
#pragma unroll
        for (int r = 0; r < Q; ++r) thr[r] = thr_l[r];
    }
}

// The same unit in the pruned search, which scans only those groups of the warp's query rows
// (GROUPS = 4: quarters of two rows) in which the exact test could not rule out every row
// (`rows`): an unscanned row keeps tm = +inf and never flags.  A group's operands (-2 q and the
// thresholds) are fetched for the scan
// from the warp's shared-memory copy of its queries (qsm: [3][32 * Q], hi xyz)
// and from the resolve state, so that they occupy registers only while a scan runs.
template <int Q, int SUB, int UNR, int PARTS, int GROUPS>
__device__ __forceinline__ void scan_subtile_pruned(const NN2Params &p, const float *__restrict__ gq,
                                                    const float *__restrict__ gt, int q0, int lane,
                                                    const float4 *sx, const float4 *sy, const float4 *sz,
                                                    const float4 *sn, int gbase, float *mt_l, float *thr_l,
                                                    float *tm_l, double *Dbest_l, int *ibest_l, float *dq_l,
                                                    float &dmax, unsigned &nflag, unsigned &npass,
                                                    const float *qsm, const float *qlo, unsigned rows
#ifdef ISR_PHASE_LOG
                                                    , long long &acc_resolve
#endif
                                                    ) {
    static_assert(Q == 8 && (GROUPS == 2 || GROUPS == 4), "halves of four rows or quarters of two");
    constexpr int H = Q / GROUPS;
    float tm[Q];
#pragma unroll
    for (int r = 0; r < Q; ++r) tm[r] = CUDART_INF_F;
    unsigned pflags = 0;  // bit (piece * 8 + row)
#pragma unroll
    for (int half = 0; half < GROUPS; ++half) {
        if ((rows >> (half * H)) & ((1u << H) - 1u)) {
            float hx[H], hy[H], hz[H], ht[H], tmh[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                const int qs = (half * H + r) * 32 + lane;
                hx[r] = -2.0f * qsm[qs];
                hy[r] = -2.0f * qsm[32 * Q + qs];
                hz[r] = -2.0f * qsm[2 * 32 * Q + qs];
                ht[r] = thr_l[half * H + r];
                tmh[r] = CUDART_INF_F;
            }
            unsigned pf = 0;
            scan_rows<H, 0, H, SUB, UNR, PARTS>(sx, sy, sz, sn, hx, hy, hz, ht, tmh, pf);
            pflags |= pf << (half * H);  // rows 0..3 of every piece's byte -> rows half * 4 ..
#pragma unroll
            for (int r = 0; r < H; ++r) tm[half * H + r] = tmh[r];
        }
    }
#ifdef ISR_PHASE_LOG
    const long long t_res = clock64();
#endif
    resolve_flagged<Q, SUB, true, PARTS>(p, gq, gt, q0, lane, sx, sy, sz, sn, gbase, pflags, tm, mt_l, thr_l, tm_l,
                                         Dbest_l, ibest_l, dq_l, dmax, nflag, npass, qsm, qlo);
#ifdef ISR_PHASE_LOG
    acc_resolve += clock64() - t_res;
#endif
}

// ---- exhaustive kernel: every stage, every sub-tile -----------------------------------------
// 128 threads x 8 queries = 1024 queries per CTA; targets stream through shared memory in
// 1024-point stages (4 planes, 16 KB), 3 in flight: four cp.async.bulk (UBLKCP, TMA engine)
// per stage on an mbarrier.
template <int Q, int THREADS, int STAGE, int NSTAGES, int SUB, int MINB, int UNR>
__global__ void __launch_bounds__(THREADS, MINB) nn2_kernel(const NN2Params p) {
    static_assert(STAGE % SUB == 0 && SUB % 8 == 0 && Q <= 16, "tile shapes");
    constexpr int WARPS = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sbuf = reinterpret_cast<float *>(smem_raw);  // [NSTAGES][4][STAGE]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSTAGES * 4 * STAGE * 4);

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
    // Resolve state (see scan_subtile)
    float mt_l[Q], thr_l[Q], tm_l[Q];
    double Dbest_l[Q];
    int ibest_l[Q];
    float dmax = 0.f;
    unsigned nflag_unused = 0;
    // a warp owns 32*Q consecutive stored queries; lane l holds queries l, 32+l, ...: every
    // load below is one coalesced 128-byte line
    const int q0 = blockIdx.x * (THREADS * Q) + warp * (32 * Q) + lane;
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
    }
    for (int r = 0; r < Q; ++r) {  // run-time loop on purpose
        mt_l[r] = CUDART_INF_F;
        thr_l[r] = (q0 + r * 32 < p.nq) && !p.debug_no_resolve ? CUDART_INF_F : -CUDART_INF_F;
        Dbest_l[r] = CUDART_INF;
        ibest_l[r] = s_begin * STAGE;
    }

    // Scan order: the target stage whose centre is nearest to this CTA's query block goes
    // first (clouds are stored in Hilbert order, so both are compact patches).  After that
    // one stage every query already holds a near-final bound and the remaining stages
    // almost never reach the resolve path.  The rest follows in ascending order.
    int s_first = 0;
    if (p.stage_c != nullptr && nst > 1) {
        __shared__ float cred[4][WARPS];
        __shared__ u64 sred[WARPS];
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
        __syncthreads();
        best = sred[0];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) best = sred[w] < best ? sred[w] : best;
        s_first = (int)(unsigned)(best & 0xffffffffull);
        if (s_first >= nst) s_first = 0;
    }
    auto stage_of = [&](int pos) { return pos == 0 ? s_first : (pos - 1 < s_first ? pos - 1 : pos); };

    auto issue = [&](int sl) {
        const int slot = sl % NSTAGES;
        float *dst = sbuf + (size_t)slot * 4 * STAGE;
        const float *src = gt + (long long)(s_begin + stage_of(sl)) * STAGE;
        mbar_expect_tx(&full[slot], 4u * STAGE * 4u);
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
            bulk_g2s(dst + pl * STAGE, src + (long long)pl * p.nt_pad, STAGE * 4u, &full[slot]);
    };
    if (tid == 0) {
        for (int i = 0; i < NSTAGES - 1 && i < nst; ++i) issue(i);
    }

    // (A per-slot `empty` mbarrier instead of the CTA barrier below was measured 2 % slower:
    // the 5 % of samples parked at the barrier are warps that would otherwise only run ahead.)
    for (int sl = 0; sl < nst; ++sl) {
        if (tid == 0 && sl + NSTAGES - 1 < nst) issue(sl + NSTAGES - 1);
        const int slot = sl % NSTAGES;
        mbar_wait(&full[slot], (sl / NSTAGES) & 1);
        const float4 *sx = reinterpret_cast<const float4 *>(sbuf + (size_t)slot * 4 * STAGE);
        const float4 *sy = sx + STAGE / 4;
        const float4 *sz = sy + STAGE / 4;
        const float4 *sn = sz + STAGE / 4;
        const int gstage = (s_begin + stage_of(sl)) * STAGE;
#pragma unroll 1
        for (int sub = 0; sub < STAGE / SUB; ++sub) {
            scan_subtile<Q, SUB, UNR>(p, gq, gt, q0, lane, sx + sub * (SUB / 4),
                                             sy + sub * (SUB / 4), sz + sub * (SUB / 4),
                                             sn + sub * (SUB / 4), gstage + sub * SUB, q2x, q2y, q2z,
                                             thr, mt_l, thr_l, tm_l, Dbest_l, ibest_l, nullptr, dmax,
                                             nflag_unused, nflag_unused);
        }
        __syncthreads();  // every warp is done with this slot before it is refilled
    }
    if (p.evaluated != nullptr && lane == 0) {
        atomicAdd(p.evaluated + 0, 4ull * nst * (STAGE / SUB));  // in quarter units (64 queries x SUB targets)
        atomicAdd(p.evaluated + 1, (unsigned long long)nst);
        atomicAdd(p.evaluated + 4, 1ull);
    }

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

// ---- pruned kernel: warp-autonomous, sub-tile granularity ------------------------------------
// (developer A/B builds override these: scripts/build_variants.sh.  Measured with the bounds in shared
// memory, 20 CTAs per SM unless noted -- verification candidates/s | 1M x 1M ICP ms | ADD-S pairs/s:
// FIFO 96 / ring 3: 10628 | 0.529 | 47.3k;  72 / 4: 10499 | 0.528 | 47.0k;  64 / 4: 10455 | 0.525 | 47.1k;
// 96 / 4 (19 CTAs per SM): 10419 | 0.530 | 47.1k;  before, bounds in local memory and the queries' lo
// parts in shared memory (16 CTAs per SM for ICP): 10560 | 0.550 | 46.6k)
#ifndef ISR_NN_RING
#define ISR_NN_RING 3
#endif
#ifndef ISR_NN_FIFO
#define ISR_NN_FIFO 96
#endif
constexpr int kRing = ISR_NN_RING;  // sub-tile buffers in flight per warp (fused ICP iteration)
constexpr int kFifo = ISR_NN_FIFO;  // candidate sub-tiles queued per warp (the nearest-first sort handles up to 64)
// The plain search (verification, ADD-S, stepwise ICP) needs neither the fused epilogue's scratch nor
// a third buffer (the wait for a sub-tile is 0.4 % of a CTA's cycles with three): a two-slot ring and a
// shorter FIFO bring a warp's shared memory from 10.1 to 8.6 KB, and with the 80 registers the compiler
// then settles for, 24 one-warp CTAs fit an SM instead of 20.
#ifndef ISR_NN_RING_PLAIN
#define ISR_NN_RING_PLAIN 2
#endif
#ifndef ISR_NN_FIFO_PLAIN
#define ISR_NN_FIFO_PLAIN 96
#endif
#ifndef ISR_NN_MINB_PLAIN
#define ISR_NN_MINB_PLAIN 21
#endif
#ifndef ISR_NN_MINB_FUSED
#define ISR_NN_MINB_FUSED 20
#endif
constexpr int kAnchors = 8; // seeds per warp: one per query row
#ifndef ISR_SEED_OWN
#define ISR_SEED_OWN 0
#endif
#ifndef ISR_LO_PREFETCH
#define ISR_LO_PREFETCH 0
#endif
constexpr float kCutGap = 3.0f;   // SPLIT: a row is cut at gaps wider than its radius / kCutGap ...
constexpr float kCutGain = 0.75f; //        ... when every piece is then at most this fraction of its radius wide

template <int SUB, int Q, bool SPLIT, int RING, int FIFO>
struct alignas(128) PrunedWarpSmem {
    float buf[RING][4][SUB];   // x, y, z, |p|^2 of one sub-tile per slot
    float4 sph[FIFO];          // queued candidates: sphere,
    int id[FIFO];              //   sub-tile index in the target (-1: dropped by the exact test),
    unsigned rows[FIFO];       //   query rows that the coarse test could not rule out
    unsigned box[FIFO];        //   packed half-extents of its bounding box about the sphere's centre
    float4 row[Q];             // sphere (c, rho) of query row r = the 32 queries r*32 .. r*32+31
    float4 rowx[SPLIT ? Q : 1][3];  // SPLIT: further spheres of a row that the curve leaves and re-enters (w < 0: none)
    float rowB[Q];             // max of their bounds dq (refreshed between batches of work)
    uint64_t full[RING];
    uint64_t qbar;             // mbarrier of the prologue's bulk copies of the warp's queries
    int seed[kAnchors];        // sub-tiles scanned first (-1: none)
    unsigned seedrows[kAnchors];  // query rows that a seed's first scan covers (the walk queues it for the others)
    int seedpos[kAnchors];     // its FIFO entry
    int seedsd[kAnchors];      // scratch of the seed search: nearest stage, then seed sub-tile, of every anchor
    alignas(16) float qs[3][32 * Q];  // the warp's queries, hi xyz (read by the scan, the tests and the resolve path)
    // dq[r][lane] >= the exact best distance so far of query r * 32 + lane (0: no such query).  Read by
    // every exact test; in local memory (with the rest of the resolve state) five of six of those
    // loads missed the small L1 that the shared-memory carve-out leaves.  The queries' lo parts, which
    // only the hints and the resolve's FP64 pass read, went the other way (local memory).
    float dq[Q][32];
};

// A launch-list entry: query block (20 bits) | the rows of the block that the CTA owns (8-bit mask, a
// run of consecutive rows) << 20 | target part (0: the whole target; k: part k - 1 of kTargetParts) << 28
__host__ __device__ inline int launch_entry(int blk, unsigned rows, int tpart) {
    return (int)((unsigned)blk | (rows << 20) | ((unsigned)tpart << 28));
}
// rows of part c of a block that runs as `parts` CTAs of 8 / parts rows
__host__ __device__ inline unsigned rows_of_part(int parts, int c) {
    const int n = 8 / parts;
    return ((1u << n) - 1u) << (n * c);
}

// FUSED: the kernel is one whole ICP evaluation + update (IcpFuse, icp_device.cuh): the queries
// are made from the original source and the start's pose in the prologue, and the epilogue turns
// the neighbours into the 17 correspondence sums, reduces them across the grid and solves.
template <int Q, int WARPS, int SUB, int MINB, int UNR, int FLAG, int PARTS, int GROUPS = 2, bool FUSED = false,
          bool SPLIT = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
nn2_pruned_kernel(const NN2Params p, const __grid_constant__ IcpFuse f) {
    static_assert(SUB == ISR_SUB_TILE && ISR_SOA_TILE / SUB == 16 && Q == 8, "pruning uses the spheres of prepare.cu");
    static_assert(!FUSED || WARPS == 1, "the fused tail is per one-warp CTA");
    constexpr int SUBS = ISR_SOA_TILE / SUB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.z;
    // (FUSED: the skip flag is the start's `done`, written by the previous iteration's launch, which
    // may still be running -- it is read after pdl_wait() below)
    if (!FUSED && p.skip != nullptr && p.skip[(long long)b * p.skip_stride] != 0) return;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // heaviest query blocks first (order from block_order_kernel): the grid is only a few
    // waves deep for a single cloud pair, and a late-starting heavy block would be its tail.
    // The very widest blocks are split over 8 CTAs that own one query row each 
    int blk = (int)blockIdx.x, tpart = 0;
    unsigned own = 0xFFu;  // rows this CTA is responsible for
    if (p.order != nullptr) {
        if ((int)blockIdx.x >= p.order_count[b]) return;
        const int entry = p.order[(long long)b * gridDim.x + blockIdx.x];
        blk = entry & 0xFFFFF;
        own = ((unsigned)entry >> 20) & 0xFFu;
        tpart = FUSED && SPLIT ? (entry >> 28) & 7 : 0;  // 0: the whole target; k: part k - 1 of kTargetParts
    }
    const int q0 = blk * (WARPS * 32 * Q) + warp * (32 * Q) + lane;
    if (!FUSED && q0 - lane >= p.nq) return;  // this warp has no live query; warps never meet at a barrier
    unsigned livemask = 0;  // bit r: query r * 32 + lane exists and belongs to this CTA
#pragma unroll
    for (int r = 0; r < Q; ++r)
        if (q0 + r * 32 < p.nq && ((own >> r) & 1u)) livemask |= 1u << r;
    const unsigned liverows = __reduce_or_sync(0xffffffffu, livemask);  // rows with a live query
    // (they are consecutive: a CTA owns a run of rows, a cloud ends inside one); the per-row loops
    // of the coarse tests walk this range only
    const int r_lo = liverows != 0 ? __ffs(liverows) - 1 : 0, r_hi = 32 - __clz(liverows);
    if (liverows == 0) {
        if (FUSED) {
            pdl_wait();
            pdl_launch_dependents();
            if (p.skip != nullptr && p.skip[(long long)b * p.skip_stride] != 0) return;
        }
        if (FUSED && p.cost != nullptr && lane == 0 && b == 0) p.cost[blockIdx.x] = 0u;
        if (FUSED && tpart <= 1) {  // nothing to search, but the reduction counts on every row of the launch list
            double rs0[Q];
#pragma unroll 1
            for (int r = 0; r < Q; ++r) rs0[r] = 0.0;
            icp_fused_tail(f, b, blk, own, lane, rs0);
        }
        return;
    }
    long long t_start = clock64();
#ifdef ISR_PHASE_LOG
    long long ph_q = 0, ph_h = 0, ph_r = 0, ph_s = 0;
    if (p.cta_log != nullptr && lane == 0) atomicMin(p.cta_log + 4 * (p.cta_log_cap - 1) + 0, global_ns());
#endif
    constexpr int RING = FUSED ? kRing : ISR_NN_RING_PLAIN, FIFO = FUSED ? kFifo : ISR_NN_FIFO_PLAIN;
    using WS = PrunedWarpSmem<SUB, Q, SPLIT, RING, FIFO>;
    WS &ws = reinterpret_cast<WS *>(smem_raw)[warp];
    const float *__restrict__ gq = p.q + (long long)b * p.q_bstride;
    const float *__restrict__ gt = p.t + (long long)b * p.t_bstride;
    const float4 *__restrict__ stage_c = p.stage_c + (long long)b * p.stage_c_bstride;
    const float4 *__restrict__ sub_c = p.sub_c + (long long)b * p.sub_c_bstride;
    const unsigned *__restrict__ sub_h = p.sub_h != nullptr ? p.sub_h + (long long)b * p.sub_h_bstride : nullptr;
    const int stages = p.stages_total;

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < RING; ++i) mbar_init(&ws.full[i], 1);
        mbar_init(&ws.qbar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    // The warp's 256 queries are 1 KB runs of the stored planes: one bulk copy per plane (TMA
    // engine) straight into the shared-memory copy -- one round trip to L2 instead of one per
    // query row.  FUSED: the planes are the ORIGINAL source; the pose is applied below, in place.
    if (lane == 0) {
        const float *base = FUSED ? f.src7 : gq;
        const bool lo = FUSED || p.use_lo;
        const long long qbase = (long long)(q0 - lane);
        mbar_expect_tx(&ws.qbar, (lo ? 6u : 3u) * 32u * Q * 4u);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl)
            bulk_g2s(&ws.qs[pl][0], base + (long long)pl * p.nq_pad + qbase, 32u * Q * 4u, &ws.qbar);
        if (lo) {  // the lo planes pass through the (still idle) ring buffer on their way to local memory
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
                bulk_g2s(&ws.buf[0][0][0] + pl * 32 * Q, base + (long long)(4 + pl) * p.nq_pad + qbase, 32u * Q * 4u,
                         &ws.qbar);
        }
    }

    // Everything per query row is a RUN-TIME loop over r from here on (the queries live in
    // shared memory, the resolve state in local memory): the kernel's code is what competes
    // for the instruction cache when the warps of an SM sit in different phases, and every
    // loop unrolled over the 8 rows multiplies it (124 KB of SASS before, no_instruction the
    // fastest-growing stall when more warps were made resident).
    static_assert(offsetof(WS, sph) == sizeof(WS::buf) && sizeof(WS::buf) + sizeof(WS::sph) >= 3 * 32 * Q * sizeof(float),
                  "lo planes land in the ring buffer (and, with two slots, the still empty FIFO behind it)");
    float mt_l[Q], thr_l[Q], tm_l[Q];
    float qlo[3 * Q];  // lo parts of the lane's queries: x of rows 0..7, then y, then z (local memory)
    float *dq_l = &ws.dq[0][lane];  // this lane's bounds: dq_l[r * 32]
    double Dbest_l[Q];
    int ibest_l[Q];
    float dmax = 0.f;  // max over the lane's live queries of dq_l (0: no live query)
#pragma unroll 1
    for (int r = 0; r < Q; ++r) {
        const bool live = (livemask >> r) & 1u;
        if (live) dmax = CUDART_INF_F;
        mt_l[r] = CUDART_INF_F;
        thr_l[r] = live ? CUDART_INF_F : -CUDART_INF_F;
        Dbest_l[r] = CUDART_INF;
        ibest_l[r] = 0;
        dq_l[r * 32] = live ? CUDART_INF_F : 0.f;
        qlo[r] = 0.f; qlo[Q + r] = 0.f; qlo[2 * Q + r] = 0.f;
    }
    // FUSED: everything up to here -- the launch-list entry, the barriers, the copy of the ORIGINAL source
    // points -- is the same in every iteration of a run; what follows reads what the previous
    // iteration's launch wrote (done flag, pose, hints, tickets).  With programmatic dependent launch
    // (isr_icp_run*: every launch after the first) this CTA may have started while that launch was
    // still in its tail / exchange / solve: wait for it here, then let the next one queue up behind.
    // (the start's pose is fetched while the query copy is in flight)
    double Tf[12], cf[3];
    if (FUSED) {
        pdl_wait();
        pdl_launch_dependents();
#ifndef ISR_PHASE_LOG
        t_start = clock64();  // (measured cost and CTA log: without the wait for the previous launch)
#endif
        if (p.skip != nullptr && p.skip[(long long)b * p.skip_stride] != 0) {  // warp-uniform
            mbar_wait(&ws.qbar, 0);  // no bulk copy may be in flight into a CTA that exits
            return;
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) Tf[k] = f.states[b].T[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) cf[k] = f.centroid[k];
    }
    mbar_wait(&ws.qbar, 0);
    if (FUSED) {
        // the arithmetic of prepare_soa7_kernel: FP64 R p + t - c, split into a float32 hi/lo pair
#pragma unroll 1
        for (int r = 0; r < Q; ++r) {
            if (!((own >> r) & 1u)) continue;  // (rows of other CTAs keep the raw copy; never read)
            const int qs = r * 32 + lane;
            float hx = ISR_PAD_COORD, hy = ISR_PAD_COORD, hz = ISR_PAD_COORD, lx = 0.f, ly = 0.f, lz = 0.f;
            if (q0 + r * 32 < p.nq) {
                const float *lo_in = &ws.buf[0][0][0];
                const double px = (double)ws.qs[0][qs] + (double)lo_in[qs];
                const double py = (double)ws.qs[1][qs] + (double)lo_in[32 * Q + qs];
                const double pz = (double)ws.qs[2][qs] + (double)lo_in[2 * 32 * Q + qs];
                double x, y, z;
                icp_apply_pose(Tf, px, py, pz, x, y, z);
                x -= cf[0]; y -= cf[1]; z -= cf[2];
                hx = (float)x; hy = (float)y; hz = (float)z;
                lx = (float)(x - (double)hx); ly = (float)(y - (double)hy); lz = (float)(z - (double)hz);
            }
            ws.qs[0][qs] = hx; ws.qs[1][qs] = hy; ws.qs[2][qs] = hz;
            qlo[r] = lx; qlo[Q + r] = ly; qlo[2 * Q + r] = lz;
        }
    } else if (p.use_lo) {
        const float *lo_in = &ws.buf[0][0][0];
#pragma unroll 1
        for (int r = 0; r < Q; ++r) {
            const int qs = r * 32 + lane;
            qlo[r] = lo_in[qs]; qlo[Q + r] = lo_in[32 * Q + qs]; qlo[2 * Q + r] = lo_in[2 * 32 * Q + qs];
        }
    }
    __syncwarp();  // every lane has its lo parts before the ring buffer is handed to the bulk copies
#ifdef ISR_PHASE_LOG
    ph_q = clock64() - t_start;
#endif

    // ---- starting bounds from the caller's hints (the previous search's neighbours) ----------
    // A hinted query starts as if its hint had already been scanned and resolved: running
    // filter minimum = the hint's filter value, best = its exact FP64 distance.
    bool all_hinted = false;
    if (p.hint != nullptr) {
        __syncwarp();  // ws.qs is complete
        bool ok = true;
#pragma unroll 1
        for (int r = 0; r < Q; ++r) {  // run-time loop: the resolve state lives in local memory
            const int i = q0 + r * 32;
            if (!((livemask >> r) & 1u)) continue;
            const int h = p.hint[(long long)b * p.nq_pad + i];
            if (h < 0 || h >= p.nt) { ok = false; continue; }
            const int qs = r * 32 + lane;
            const float qhx = ws.qs[0][qs], qhy = ws.qs[1][qs], qhz = ws.qs[2][qs];
            const float px = gt[h], py = gt[p.nt_pad + h], pz = gt[2ll * p.nt_pad + h];
            const float a = __fmaf_rn(-2.0f * qhx, px, __fmaf_rn(-2.0f * qhy, py,
                                      __fmaf_rn(-2.0f * qhz, pz, gt[3ll * p.nt_pad + h])));
            double dx = (double)qhx - (double)px, dy = (double)qhy - (double)py, dz = (double)qhz - (double)pz;
            if (p.use_lo) {
                dx += (double)qlo[r] - (double)gt[4ll * p.nt_pad + h];
                dy += (double)qlo[Q + r] - (double)gt[5ll * p.nt_pad + h];
                dz += (double)qlo[2 * Q + r] - (double)gt[6ll * p.nt_pad + h];
            }
            const double D = fma(dz, dz, fma(dy, dy, dx * dx));
            const float nq2 = __fmaf_rn(qhz, qhz, __fmaf_rn(qhy, qhy, qhx * qhx));
            mt_l[r] = a;
            thr_l[r] = filter_threshold(a, nq2, sqrtf(nq2));
            Dbest_l[r] = D;
            ibest_l[r] = h;
            dq_l[r * 32] = __fsqrt_ru(__double2float_ru(D)) * 1.00002f + 1e-6f * sqrtf(nq2) + 1e-37f;
        }
        all_hinted = __all_sync(0xffffffffu, ok);
        float m = 0.f;
#pragma unroll 1
        for (int r = 0; r < Q; ++r) m = fmaxf(m, dq_l[r * 32]);
        dmax = m;
    }

#ifdef ISR_PHASE_LOG
    ph_h = clock64() - t_start;
#endif
    // SPLIT: cut rows at curve jumps (below).  A separate instantiation, launched only when the grid
    // is at most one wave of whole blocks deep: there the slowest warp IS the iteration; in deeper
    // grids the extra code and coarse-test work cost more than the shorter tail gains (1M x 1M
    // on one GPU: -5 %, also when the code is merely present behind a run-time flag)
    constexpr bool split_rows = SPLIT;
    // ---- query-row spheres: row r is 32 consecutive stored queries, a compact patch ----------
#pragma unroll 1
    for (int r = 0; r < Q; ++r) {
        if (!((liverows >> r) & 1u)) {  // warp-uniform: a row of another CTA of this block, or beyond the cloud
            if (lane == 0) {
                ws.row[r] = make_float4(0.f, 0.f, 0.f, -1.f);
                ws.rowB[r] = 0.f;
                if (split_rows) ws.rowx[r][0] = ws.rowx[r][1] = ws.rowx[r][2] = make_float4(0.f, 0.f, 0.f, -1.f);
            }
            continue;
        }
        const bool live = (livemask >> r) & 1u;
        const float qx = ws.qs[0][r * 32 + lane], qy = ws.qs[1][r * 32 + lane], qz = ws.qs[2][r * 32 + lane];
        float cx = live ? qx : 0.f, cy = live ? qy : 0.f, cz = live ? qz : 0.f, cn = live ? 1.f : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cx += __shfl_xor_sync(0xffffffffu, cx, o);
            cy += __shfl_xor_sync(0xffffffffu, cy, o);
            cz += __shfl_xor_sync(0xffffffffu, cz, o);
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
        }
        const float inv = cn > 0.f ? 1.0f / cn : 0.f;
        cx *= inv; cy *= inv; cz *= inv;
        const float dx = qx - cx, dy = qy - cy, dz = qz - cz;
        float m = live ? fmaf(dz, dz, fmaf(dy, dy, dx * dx)) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) {
            ws.row[r] = make_float4(cx, cy, cz, cn > 0.f ? __fsqrt_ru(m) * 1.00002f : -1.f);
            ws.rowB[r] = cn > 0.f ? CUDART_INF_F : 0.f;
        }
        if (split_rows) {
            // A row is 32 consecutive points of the curve.  Where the curve leaves the surface and
            // re-enters it elsewhere, the row is several compact patches far apart and its ONE
            // sphere covers everything in between: most of the target's stages and hundreds of
            // sub-tiles pass the coarse tests and get an exact test each -- the slowest warps of a
            // sharded ICP iteration, i.e. its duration.  Such a row is cut at its (up to three)
            // largest gaps between consecutive points that exceed a third of its radius, and keeps
            // one sphere per piece when every piece is clearly tighter than the row; the coarse
            // tests then pass what reaches ANY piece.
            const unsigned full = 0xffffffffu;
            const float nx = __shfl_down_sync(full, qx, 1), ny = __shfl_down_sync(full, qy, 1),
                        nz = __shfl_down_sync(full, qz, 1);
            const bool nlive = __shfl_down_sync(full, live ? 1 : 0, 1) != 0 && lane < 31;
            const float gx = nx - qx, gy = ny - qy, gz = nz - qz;
            float gap = live && nlive ? fmaf(gz, gz, fmaf(gy, gy, gx * gx)) : -1.f;
            unsigned cuts = 0;  // bit l: the row is cut between lanes l and l + 1
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
                unsigned key = gap >= 0.f ? ((__float_as_uint(gap) & 0xFFFFFFE0u) | (unsigned)lane) : 0u;
                key = __reduce_max_sync(full, key);
                if (key == 0u || !(kCutGap * kCutGap * __uint_as_float(key & 0xFFFFFFE0u) > m)) break;  // gap <= radius / kCutGap
                cuts |= 1u << (key & 31u);
                if (lane == (int)(key & 31u)) gap = -1.f;
            }
            float4 seg[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) seg[k] = make_float4(0.f, 0.f, 0.f, -1.f);
            if (cuts != 0u) {  // warp-uniform, rare
                const int mine = __popc(cuts & ((1u << lane) - 1u));  // piece of this lane
                float worst = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const bool in = live && mine == k;
                    float sx = in ? qx : 0.f, sy = in ? qy : 0.f, sz = in ? qz : 0.f, sn = in ? 1.f : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sx += __shfl_xor_sync(full, sx, o); sy += __shfl_xor_sync(full, sy, o);
                        sz += __shfl_xor_sync(full, sz, o); sn += __shfl_xor_sync(full, sn, o);
                    }
                    if (!(sn > 0.f)) continue;
                    const float is = 1.f / sn;
                    sx *= is; sy *= is; sz *= is;
                    const float ux = qx - sx, uy = qy - sy, uz = qz - sz;
                    float d2 = in ? fmaf(uz, uz, fmaf(uy, uy, ux * ux)) : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) d2 = fmaxf(d2, __shfl_xor_sync(full, d2, o));
                    seg[k] = make_float4(sx, sy, sz, __fsqrt_ru(d2) * 1.00002f);
                    worst = fmaxf(worst, d2);
                }
                if (!(worst < kCutGain * kCutGain * m)) {  // the widest piece is not clearly tighter than the row: keep it whole
#pragma unroll
                    for (int k = 0; k < 4; ++k) seg[k].w = -1.f;
                }
            }
            if (lane == 0) {
                if (seg[0].w >= 0.f) ws.row[r] = seg[0];  // (the seed of this row then starts from its first piece)
                ws.rowx[r][0] = seg[0].w >= 0.f ? seg[1] : make_float4(0.f, 0.f, 0.f, -1.f);
                ws.rowx[r][1] = seg[0].w >= 0.f ? seg[2] : make_float4(0.f, 0.f, 0.f, -1.f);
                ws.rowx[r][2] = seg[0].w >= 0.f ? seg[3] : make_float4(0.f, 0.f, 0.f, -1.f);
            }
        }
    }
    __syncwarp();
    // rowB[r] <- max over the row of the current bounds (they only ever shrink, so a stale
    // value is merely conservative)
    auto refresh_bounds = [&]() {
#pragma unroll 1
        for (int r = r_lo; r < r_hi; ++r) {
            float m = dq_l[r * 32];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) ws.rowB[r] = m;
        }
        __syncwarp();
    };
    // coarse test, one candidate sphere per lane: which query rows can it still matter to?
    auto coarse_rows = [&](const float4 S) {
        unsigned rows = 0;
#pragma unroll 1
        for (int r = r_lo; r < r_hi; ++r) {
            const float4 R = ws.row[r];
            const float dx = S.x - R.x, dy = S.y - R.y, dz = S.z - R.z;
            const float rr = (ws.rowB[r] + R.w + S.w) * 1.0001f;
            if (R.w >= 0.f && !(fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > rr * rr)) rows |= 1u << r;
            if (split_rows && ws.rowx[r][0].w >= 0.f) {  // warp-uniform, rare
#pragma unroll 1
                for (int k = 0; k < 3; ++k) {
                    const float4 R2 = ws.rowx[r][k];
                    if (R2.w < 0.f) break;
                    const float ex = S.x - R2.x, ey = S.y - R2.y, ez = S.z - R2.z;
                    const float r2 = (ws.rowB[r] + R2.w + S.w) * 1.0001f;
                    if (!(fmaf(ez, ez, fmaf(ey, ey, ex * ex)) > r2 * r2)) rows |= 1u << r;
                }
            }
        }
        return S.w >= 0.f ? rows : 0u;
    };
    // the same for a sub-tile with its box (see exact_any_box): the row's sphere must reach both
    auto coarse_rows_box = [&](const float4 S, unsigned hb) {
        const float sc = S.w * (1.00001f / 1023.f);
        const float hx = (float)(hb & 1023u) * sc, hy = (float)((hb >> 10) & 1023u) * sc,
                    hz = (float)((hb >> 20) & 1023u) * sc;
        unsigned rows = 0;
#pragma unroll 1
        for (int r = r_lo; r < r_hi; ++r) {
            const float4 R = ws.row[r];
            const float dx = S.x - R.x, dy = S.y - R.y, dz = S.z - R.z;
            const float rb = (ws.rowB[r] + R.w) * 1.0001f, rr = rb + S.w * 1.0001f;
            const float ex = fmaxf(fabsf(dx) - hx, 0.f), ey = fmaxf(fabsf(dy) - hy, 0.f),
                        ez = fmaxf(fabsf(dz) - hz, 0.f);
            const bool out = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > rr * rr ||
                             fmaf(ez, ez, fmaf(ey, ey, ex * ex)) > rb * rb;
            if (R.w >= 0.f && !out) rows |= 1u << r;
            if (split_rows && ws.rowx[r][0].w >= 0.f) {  // warp-uniform, rare
#pragma unroll 1
                for (int k = 0; k < 3; ++k) {
                    const float4 R2 = ws.rowx[r][k];
                    if (R2.w < 0.f) break;
                    const float fx = S.x - R2.x, fy = S.y - R2.y, fz = S.z - R2.z;
                    const float rb2 = (ws.rowB[r] + R2.w) * 1.0001f, rr2 = rb2 + S.w * 1.0001f;
                    const float gx = fmaxf(fabsf(fx) - hx, 0.f), gy = fmaxf(fabsf(fy) - hy, 0.f),
                                gz = fmaxf(fabsf(fz) - hz, 0.f);
                    const bool out2 = fmaf(fz, fz, fmaf(fy, fy, fx * fx)) > rr2 * rr2 ||
                                      fmaf(gz, gz, fmaf(gy, gy, gx * gx)) > rb2 * rb2;
                    if (!out2) rows |= 1u << r;
                }
            }
        }
        return S.w >= 0.f ? rows : 0u;
    };
    // exact test, one query per lane and row: is any query of `rows` not ruled out?
    auto exact_any = [&](const float4 S, unsigned rows) {
        float dq[Q];  // (unrolled with the bounds loaded together, like exact_rows_box below)
#pragma unroll
        for (int r = 0; r < Q; ++r) dq[r] = dq_l[r * 32];
        bool need = false;
#pragma unroll
        for (int r = 0; r < Q; ++r) {
            if (rows & (1u << r)) {  // warp-uniform
                const float dx = ws.qs[0][r * 32 + lane] - S.x, dy = ws.qs[1][r * 32 + lane] - S.y,
                            dz = ws.qs[2][r * 32 + lane] - S.z;
                const float rr = (dq[r] + S.w) * 1.0001f;
                // a dead query slot (dq = 0, padded coordinates ~1e18) is always ruled out
                need = need || !(fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > rr * rr);
            }
        }
        return __any_sync(0xffffffffu, need);
    };
    // the same for a sub-tile that also carries its axis-aligned box (centre = the sphere's,
    // half-extents = hb's three 10-bit fractions of the radius, rounded up): a query is ruled
    // out when EITHER volume is farther than its bound.  The patches are thin sheets of the
    // surface; the box follows them where the sphere is mostly empty.
    auto exact_rows_box = [&](const float4 S, unsigned rows, unsigned hb) {
        const float sc = S.w * (1.00001f / 1023.f);
        const float hx = (float)(hb & 1023u) * sc, hy = (float)((hb >> 10) & 1023u) * sc,
                    hz = (float)((hb >> 20) & 1023u) * sc;
        // (the one per-row loop that stays unrolled: it runs ~60 times per warp, and with a
        // run-time row index every row waited for its own bound from local memory -- 7.5 % of
        // all samples on that one load; here the eight loads go out together)
        float dq[Q];
#pragma unroll
        for (int r = 0; r < Q; ++r) dq[r] = dq_l[r * 32];
        unsigned need = 0;
#pragma unroll
        for (int r = 0; r < Q; ++r) {
            if (rows & (1u << r)) {  // warp-uniform
                const float dx = ws.qs[0][r * 32 + lane] - S.x, dy = ws.qs[1][r * 32 + lane] - S.y,
                            dz = ws.qs[2][r * 32 + lane] - S.z;
                const float rr = (dq[r] + S.w) * 1.0001f;
                const float ex = fmaxf(fabsf(dx) - hx, 0.f), ey = fmaxf(fabsf(dy) - hy, 0.f),
                            ez = fmaxf(fabsf(dz) - hz, 0.f);
                const float rb = dq[r] * 1.0001f;
                const bool out = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > rr * rr ||
                                 fmaf(ez, ez, fmaf(ey, ey, ex * ex)) > rb * rb;
                if (__any_sync(0xffffffffu, !out)) need |= 1u << r;
            }
        }
        return need;
    };

#ifdef ISR_PHASE_LOG
    ph_r = clock64() - t_start;
#endif
    // ---- seeds: for every query row, the sub-tile whose centre is nearest to the row's --------
    int head = 0, look = 0, tail = 0, nloads = 0, nconsumed = 0;  // FIFO / ring state, warp-uniform
    if (lane < kAnchors) { ws.seed[lane] = -1; ws.seedrows[lane] = 0u; }
    __syncwarp();
    if (!all_hinted) {  // (a fully hinted warp already holds near-final bounds)
        // anchors: the centres of the query rows (a dead row falls back to the first live one)
        const int na = kAnchors < p.nanchors ? kAnchors : p.nanchors;
        // (A) the seed sub-tile of every anchor -> ws.seedsd[]
        if (!ISR_SEED_OWN && stages <= 4 * 32) {
            // a target of at most 128 stages (every verification / ADD-S cloud): each lane keeps its (up to)
            // four stage spheres in registers and the sub-tile spheres of all anchors are fetched in one
            // go, two anchors per step -- one round trip each instead of five per anchor (the seeds were
            // 7 % of the verification kernel's samples, nearly all of it waiting for those loads)
            float4 Sg[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int sidx = j * 32 + lane;
                Sg[j] = sidx < stages ? stage_c[sidx] : make_float4(0.f, 0.f, 0.f, -1.f);
            }
#pragma unroll 1
            for (int a2 = 0; a2 < na; ++a2) {
                const float4 R = ws.row[a2].w >= 0.f ? ws.row[a2] : ws.row[__ffs(liverows) - 1];
                unsigned best = ~0u;  // squared distance (low 7 bits dropped) | stage
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float dx = Sg[j].x - R.x, dy = Sg[j].y - R.y, dz = Sg[j].z - R.z;
                    const unsigned key = (__float_as_uint(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) & 0xFFFFFF80u) |
                                         (unsigned)(j * 32 + lane);
                    best = Sg[j].w >= 0.f && key < best ? key : best;
                }
                best = __reduce_min_sync(0xffffffffu, best);
                if (lane == 0) ws.seedsd[a2] = best == ~0u ? 0 : (int)(best & 127u);
            }
            __syncwarp();
            const int half = lane >> 4, sl = lane & (SUBS - 1);
            float4 Sb[4];
            int stv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int a2 = 2 * t + half;
                stv[t] = a2 < na ? ws.seedsd[a2] : 0;
                Sb[t] = a2 < na ? sub_c[(long long)stv[t] * SUBS + sl] : make_float4(0.f, 0.f, 0.f, -1.f);
            }
            __syncwarp();  // (every lane has read the stages before the slots are reused for the sub-tiles)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int a2 = 2 * t + half;
                const float4 R = ws.row[a2].w >= 0.f ? ws.row[a2] : ws.row[__ffs(liverows) - 1];
                const float dx = Sb[t].x - R.x, dy = Sb[t].y - R.y, dz = Sb[t].z - R.z;
                unsigned key = Sb[t].w >= 0.f
                                   ? ((__float_as_uint(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) & 0xFFFFFFF0u) | (unsigned)sl)
                                   : ~0u;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));  // within the half
                if (sl == 0 && a2 < na) ws.seedsd[a2] = stv[t] * SUBS + (key == ~0u ? 0 : (int)(key & 15u));
            }
            __syncwarp();
        } else {
#pragma unroll 1
            for (int a = 0; a < na; ++a) {
#if ISR_SEED_OWN
                if (!((liverows >> a) & 1u)) continue;  // warp-uniform: a row of another CTA, or beyond the cloud
                const float4 R = ws.row[a];
#else
                const float4 R = ws.row[a].w >= 0.f ? ws.row[a] : ws.row[__ffs(liverows) - 1];
#endif
                u64 bk = ~0ull;
                for (int base = 0; base < stages; base += 32) {
                    const int s = base + lane;
                    if (s < stages) {
                        const float4 S = stage_c[s];
                        if (S.w >= 0.f) {
                            const float dx = S.x - R.x, dy = S.y - R.y, dz = S.z - R.z;
                            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            const u64 key = ((u64)__float_as_uint(d) << 32) | (u64)(unsigned)s;
                            bk = key < bk ? key : bk;
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const u64 other = __shfl_xor_sync(0xffffffffu, bk, o);
                    bk = other < bk ? other : bk;
                }
                int st = (int)(unsigned)(bk & 0xffffffffull);
                if (st >= stages) st = 0;
                // nearest sub-tile centre inside that stage
                u64 key = ~0ull;
                if (lane < SUBS) {
                    const float4 S = sub_c[(long long)st * SUBS + lane];
                    if (S.w >= 0.f) {
                        const float dx = S.x - R.x, dy = S.y - R.y, dz = S.z - R.z;
                        key = ((u64)__float_as_uint(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) << 32) | (u64)(unsigned)lane;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const u64 other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other < key ? other : key;
                }
                int sb = (int)(unsigned)(key & 31ull);
                if (key == ~0ull) sb = 0;
                if (lane == 0) ws.seedsd[a] = st * SUBS + sb;
            }
            __syncwarp();
        }
        // (B) queue them, each sub-tile once
#pragma unroll 1
        for (int a = 0; a < na; ++a) {
#if ISR_SEED_OWN
            // a seed is first scanned for the rows of its own row group only: scanned for every row,
            // seed k improved (and resolved, one FP64 pass each) every row that lies nearer to it than
            // to seeds 0 .. k-1.  The regular walk queues the seed's tile for the other rows like any other
            // tile, against the bounds they hold by then.  (Measured: no gain.)
            constexpr int HS = Q / GROUPS;
            const unsigned qrows = (((1u << HS) - 1u) << (HS * (a / HS))) & liverows;
            if (!((liverows >> a) & 1u)) continue;  // warp-uniform
#else
            const unsigned qrows = (1u << Q) - 1u;
#endif
            const int sd = ws.seedsd[a];
            int dup = -1;
            for (int c = 0; c < a; ++c) dup = ws.seed[c] == sd ? c : dup;
            if (dup < 0) {
                if (lane == 0) {
                    // +inf radius: never ruled out
                    ws.sph[tail % FIFO] = make_float4(0.f, 0.f, 0.f, CUDART_INF_F);
                    ws.id[tail % FIFO] = sd;
                    ws.rows[tail % FIFO] = qrows;
                    ws.box[tail % FIFO] = kNoBox;
                    ws.seed[a] = sd;
                    ws.seedrows[a] = qrows;
                    ws.seedpos[a] = tail % FIFO;
                }
                ++tail;
            } else if (lane == 0) {  // the same tile serves another row group as well
                ws.seedrows[dup] |= qrows;
                ws.rows[ws.seedpos[dup]] |= qrows;
            }
            __syncwarp();
        }
    }

    // ---- main loop: ONE code path that either consumes a queued sub-tile or produces more -------
    // (a single scan site keeps the instruction footprint small: the warps of an SM sit in
    // different phases, so every inlined copy of the scan would compete for the instruction cache)
    //   consume: exact-test queued entries up to RING loads ahead (survivors get their bulk
    //            copy issued at once), then wait for the head entry's data and scan it;
    //   produce: next chunk of 32 stage spheres -> coarse row test, one per lane; then per
    //            candidate stage the exact test, and its 16 sub-tile spheres -> coarse row test
    //            -> FIFO.
    unsigned nscanned = 0, nhalves = 0, ntests = 0, ncand = 0, nflag = 0, npass = 0;
#ifdef ISR_PHASE_LOG
    const long long ph_m = clock64() - t_start;  // seeds found and queued
    long long acc_sort = 0, acc_test = 0, acc_wait = 0, acc_scan = 0, acc_resolve = 0, acc_produce = 0;
#define PH_T0 const long long ph_t0 = clock64();
#define PH_ADD(acc) acc += clock64() - ph_t0;
#else
#define PH_T0
#define PH_ADD(acc)
#endif
    unsigned refreshed_at = ~0u;
    bool sorted = p.sort_fifo == 0;
    bool seeding = true;   // the seeds must be scanned before anything is produced
    const int nchunks = (stages + 31) / 32;
    int gchunk = 0;        // next group of 32 chunk spheres (a chunk = 32 stages)
    int cgroup = 0;        // first chunk of the group whose candidates are in cmask
    unsigned cmask = 0;    // candidate chunks of the current group not yet expanded
    int cbase = 0;         // first stage of the chunk whose candidates are in smask
    unsigned smask = 0;    // candidate stages of the current chunk not yet expanded
    float4 Sst = make_float4(0.f, 0.f, 0.f, -1.f);  // this lane's stage sphere of the chunk
    unsigned rows_st = 0;                           // and its coarse row mask
    for (;;) {
        const int pending = tail - head;
        const bool produced_all = smask == 0 && cmask == 0 && gchunk >= nchunks;
        if (produced_all && !sorted) {
            // Everything is queued: put the untested entries in nearest-first order (squared
            // distance between the sub-tile's centre and the nearest of the query rows it may
            // matter to).  Near tiles tighten the bounds first, so that more of the far ones fail
            // their exact test and fewer scans improve a neighbour that a later scan improves
            // again.  Rank sort of at most 64 keys, two per lane, through the idle ring buffer.
            sorted = true;
            PH_T0
            const int cnt = tail - look;
            if (cnt > 2 && cnt <= 64 && look == head) {
                u64 *keys = reinterpret_cast<u64 *>(&ws.buf[0][0][0]);
                float4 S2[2];
                int id2[2];
                unsigned rows2[2], box2[2];
                u64 key2[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int j = k * 32 + lane;
                    key2[k] = ~0ull;
                    if (j < cnt) {
                        const int e = (look + j) % FIFO;
                        S2[k] = ws.sph[e]; id2[k] = ws.id[e]; rows2[k] = ws.rows[e]; box2[k] = ws.box[e];
                        float d = CUDART_INF_F;
#pragma unroll 1
                        for (int r = r_lo; r < r_hi; ++r) {
                            const float4 R = ws.row[r];
                            const float dx = S2[k].x - R.x, dy = S2[k].y - R.y, dz = S2[k].z - R.z;
                            const float dd = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                            d = (rows2[k] >> r) & 1u ? fminf(d, dd) : d;
                        }
                        key2[k] = ((u64)__float_as_uint(d) << 32) | (u64)(unsigned)j;
                    }
                    keys[j] = key2[k];
                }
                __syncwarp();
                int rank0 = 0, rank1 = 0;
                for (int j = 0; j < cnt; ++j) {
                    const u64 kj = keys[j];
                    rank0 += kj < key2[0] ? 1 : 0;
                    rank1 += kj < key2[1] ? 1 : 0;
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k * 32 + lane < cnt) {
                        const int e = (look + (k == 0 ? rank0 : rank1)) % FIFO;
                        ws.sph[e] = S2[k]; ws.id[e] = id2[k]; ws.rows[e] = rows2[k]; ws.box[e] = box2[k];
                    }
                }
                __syncwarp();
            }
            PH_ADD(acc_sort)
        }
        if (pending > 0 && (seeding || produced_all || pending > FIFO - 2 * SUBS)) {
#ifdef ISR_PHASE_LOG
            const long long ph_c0 = clock64();
#endif
            while (look < tail && nloads - nconsumed < RING) {
                const int e = look % FIFO;
                ++ntests;
                const unsigned need = exact_rows_box(ws.sph[e], ws.rows[e], ws.box[e]);
                if (need != 0) {
                    if (lane == 0) {
                        ws.rows[e] = need;
                        const int slot = nloads % RING;
                        const long long src = (long long)ws.id[e] * SUB;
                        mbar_expect_tx(&ws.full[slot], 4u * SUB * 4u);
#pragma unroll
                        for (int pl = 0; pl < 4; ++pl)
                            bulk_g2s(&ws.buf[slot][pl][0], gt + (long long)pl * p.nt_pad + src, SUB * 4u,
                                     &ws.full[slot]);
                    }
#if ISR_LO_PREFETCH
                    // the lo planes of the sub-tile (read by the resolve's FP64 pass, one dependent trip per
                    // window target) are asked into L2 now: six 128-byte lines, one per lane
                    else if (lane <= 6 && p.use_lo) {
                        const float *a = gt + (long long)(4 + (lane - 1) / 2) * p.nt_pad + (long long)ws.id[e] * SUB +
                                         ((lane - 1) & 1) * 32;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                    }
#endif
                    ++nloads;
                } else {
                    if (lane == 0) ws.id[e] = -1;
                }
                ++look;
            }
            __syncwarp();
#ifdef ISR_PHASE_LOG
            acc_test += clock64() - ph_c0;
#endif
            const int id = ws.id[head % FIFO];
            if (id >= 0) {
                const int slot = nconsumed % RING;
#ifdef ISR_PHASE_LOG
                const long long ph_w0 = clock64();
#endif
                mbar_wait(&ws.full[slot], (nconsumed / RING) & 1);
#ifdef ISR_PHASE_LOG
                const long long ph_w1 = clock64();
                acc_wait += ph_w1 - ph_w0;
                const long long res0 = acc_resolve;
#endif
                const unsigned rows_e = ws.rows[head % FIFO];
                ++nscanned;
                // counted in half units (4 rows x SUB targets): a quarter (2 rows) is half of one
                if (GROUPS == 2) {
                    nhalves += ((rows_e & 0x0Fu) ? 2u : 0u) + ((rows_e & 0xF0u) ? 2u : 0u);
                } else {
                    nhalves += ((rows_e & 0x03u) ? 1u : 0u) + ((rows_e & 0x0Cu) ? 1u : 0u) +
                               ((rows_e & 0x30u) ? 1u : 0u) + ((rows_e & 0xC0u) ? 1u : 0u);
                }
                // scanned in FLAG-sized pieces (FLAG == SUB in the shipped variant)
                const float4 *sx = reinterpret_cast<const float4 *>(&ws.buf[slot][0][0]);
#pragma unroll 1
                for (int h = 0; h < SUB / FLAG; ++h)
                    scan_subtile_pruned<Q, FLAG, UNR, PARTS, GROUPS>(p, gq, gt, q0, lane, sx + h * (FLAG / 4),
                                                     sx + SUB / 4 + h * (FLAG / 4),
                                                     sx + 2 * (SUB / 4) + h * (FLAG / 4),
                                                     sx + 3 * (SUB / 4) + h * (FLAG / 4), id * SUB + h * FLAG,
                                                     mt_l, thr_l, tm_l, Dbest_l, ibest_l, dq_l, dmax, nflag,
                                                     npass, &ws.qs[0][0], qlo, rows_e
#ifdef ISR_PHASE_LOG
                                                     , acc_resolve
#endif
                                                     );
#ifdef ISR_PHASE_LOG
                acc_scan += (clock64() - ph_w1) - (acc_resolve - res0);
#endif
                ++nconsumed;
                __syncwarp();  // every lane is done with the slot before lane 0 refills it
            }
            ++head;
            continue;
        }
        seeding = false;
        if (produced_all) break;
#ifdef ISR_PHASE_LOG
        struct PhScope { long long &a; long long t0; __device__ PhScope(long long &x) : a(x), t0(clock64()) {}
                         __device__ ~PhScope() { a += clock64() - t0; } } ph_scope(acc_produce);
#endif
        if (smask == 0) {
            // refresh the row bounds (if any scan ran since)
            if (nscanned != refreshed_at) {
                refresh_bounds();
                refreshed_at = nscanned;
            }
            if (cmask == 0) {
                // next group: coarse-test 32 chunk spheres, one per lane (a chunk sphere bounds
                // the spheres of 32 consecutive stages)
                cgroup = gchunk;
                gchunk += 32;
                if (nchunks <= 8) {  // a small target (<= 256 k points): its few chunks are all walked
                    cmask = (1u << nchunks) - 1u;
                    continue;
                }
                const int c = cgroup + lane;
                const float4 C = stage_c[stages + min(c, nchunks - 1)];
                cmask = __ballot_sync(0xffffffffu, c < nchunks && coarse_rows(C) != 0);
                continue;
            }
            // next candidate chunk: coarse-test its 32 stage spheres
            cbase = (cgroup + __ffs(cmask) - 1) * 32;
            cmask &= cmask - 1;
            const int s = cbase + lane;
            Sst = stage_c[min(s, stages - 1)];
            rows_st = s < stages ? coarse_rows(Sst) : 0u;
            smask = __ballot_sync(0xffffffffu, rows_st != 0);
            // a target part owns every kTargetParts-th stage (cbase is a multiple of 32); which
            // part looks at a stage does not depend on anyone's bounds, so no stage is lost
            if (FUSED && SPLIT && tpart != 0) smask &= (0xFFFFFFFFu / ((1u << kTargetParts) - 1u)) << (tpart - 1);
            continue;
        }
        // two candidate stages per step: lanes 0-15 take the sub-tiles of the first, lanes
        // 16-31 those of the second (one dependent global load serves both)
        const int l1 = __ffs(smask) - 1;
        smask &= smask - 1;
        const int l2 = smask != 0 ? __ffs(smask) - 1 : -1;
        smask &= smask - 1;  // (0 & anything = 0)
        bool pass1, pass2 = false;
        {
            const float4 S = make_float4(__shfl_sync(0xffffffffu, Sst.x, l1), __shfl_sync(0xffffffffu, Sst.y, l1),
                                         __shfl_sync(0xffffffffu, Sst.z, l1), __shfl_sync(0xffffffffu, Sst.w, l1));
            pass1 = exact_any(S, __shfl_sync(0xffffffffu, rows_st, l1));
        }
        if (l2 >= 0) {
            const float4 S = make_float4(__shfl_sync(0xffffffffu, Sst.x, l2), __shfl_sync(0xffffffffu, Sst.y, l2),
                                         __shfl_sync(0xffffffffu, Sst.z, l2), __shfl_sync(0xffffffffu, Sst.w, l2));
            pass2 = exact_any(S, __shfl_sync(0xffffffffu, rows_st, l2));
        }
        if (!pass1 && !pass2) continue;
        ncand += (pass1 ? 1u : 0u) + (pass2 ? 1u : 0u);
        {
            float4 S = make_float4(0.f, 0.f, 0.f, -1.f);
            unsigned rows = 0;
            const bool mine = lane < SUBS ? pass1 : pass2;
            const int gid = (cbase + (lane < SUBS ? l1 : l2)) * SUBS + (lane & (SUBS - 1));
            unsigned hb = kNoBox;
            if (mine) {
                S = sub_c[gid];
                if (sub_h != nullptr) hb = sub_h[gid];
                rows = coarse_rows_box(S, hb);
#pragma unroll
                for (int a = 0; a < kAnchors; ++a) rows = gid == ws.seed[a] ? rows & ~ws.seedrows[a] : rows;  // already scanned
            }
            const unsigned m32 = __ballot_sync(0xffffffffu, rows != 0);
            if (rows != 0) {
                const int pos = (tail + __popc(m32 & ((1u << lane) - 1u))) % FIFO;
                ws.sph[pos] = S;
                ws.id[pos] = gid;
                ws.rows[pos] = rows;
                ws.box[pos] = hb;
            }
            tail += __popc(m32);
            __syncwarp();
        }
    }

    if (FUSED && p.cost != nullptr && lane == 0 && b == 0) {
        const long long cyc = clock64() - t_start;
        p.cost[blockIdx.x] = (unsigned)(cyc < 0x7FFFFFFFll ? cyc : 0x7FFFFFFFll);
    }
#ifdef ISR_PHASE_LOG
    ph_s = clock64() - t_start;
    if (p.cta_log != nullptr && lane == 0) atomicMax(p.cta_log + 4 * (p.cta_log_cap - 1) + 1, global_ns());
    if (false) {
        const long long rec = 0;
        {
#else
    if (p.cta_log != nullptr && lane == 0) {
        const long long rec = ((long long)b * gridDim.x + blockIdx.x);
        if (rec < p.cta_log_cap) {
#endif
            unsigned long long *o = p.cta_log + 4 * rec;
            o[0] = (unsigned long long)(clock64() - t_start);
            o[1] = ((unsigned long long)nscanned << 32) | ntests;
            o[2] = ((unsigned long long)nhalves << 32) | ncand;
            o[3] = ((unsigned long long)(unsigned)blk << 32) | ((unsigned long long)own << 24) |
                   ((unsigned long long)tpart << 20) | (npass & 0xFFFFFu);
        }
    }
    if (p.evaluated != nullptr && lane == 0) {
        atomicAdd(p.evaluated + 0, (unsigned long long)nhalves);  // scanned quarter units (2 query rows x SUB targets)
        atomicAdd(p.evaluated + 1, (unsigned long long)stages);   // stage spheres tested
        atomicAdd(p.evaluated + 2, (unsigned long long)ncand);    // stages that passed
        atomicAdd(p.evaluated + 3, (unsigned long long)ntests);   // exact sub-tile tests
        atomicAdd(p.evaluated + 4, 1ull);                         // warps
        atomicAdd(p.evaluated + 5, (unsigned long long)nflag);    // scanned units that flagged
        atomicAdd(p.evaluated + 6, (unsigned long long)npass);    // resolve passes (warp level)
        // slowest warp: cycles / 1024, with its scanned units and exact tests
        const unsigned long long cyc = (unsigned long long)(clock64() - t_start) >> 10;
        atomicMax(p.evaluated + 7, (cyc << 44) | ((unsigned long long)(nscanned & 0xFFFFFu) << 24) |
                                       (unsigned long long)(ntests & 0xFFFFFFu));
    }

    if (FUSED && SPLIT && tpart != 0) {
        // ---- target parts -> the row's neighbours ---------------------------------------------------
        // Every part has searched its share of the stages (all of them from the same hints): the
        // row's neighbour is the best of the kTargetParts answers.  The last part to arrive merges.
        const int r0 = __ffs(own) - 1;  // a part entry owns exactly one row
        const int slot = p.order_slot[blockIdx.x];
        double *mD = p.tp_D + ((size_t)slot * kTargetParts) * 32;
        int *mI = p.tp_I + ((size_t)slot * kTargetParts) * 32;
        mD[(tpart - 1) * 32 + lane] = Dbest_l[r0];
        mI[(tpart - 1) * 32 + lane] = ibest_l[r0];
        __threadfence();
        __syncwarp();
        unsigned last = 0;
        if (lane == 0) last = atomicAdd(&p.tp_tick[slot], 1u) + 1u == (unsigned)kTargetParts ? 1u : 0u;
        if (__shfl_sync(0xffffffffu, last, 0) == 0) return;
        __threadfence();
        double Db = Dbest_l[r0];
        int ib = ibest_l[r0];
#pragma unroll 1
        for (int pp = 0; pp < kTargetParts; ++pp) {
            if (pp == tpart - 1) continue;
            const double D = __ldcg(&mD[pp * 32 + lane]);
            const int i2 = __ldcg(&mI[pp * 32 + lane]);
            bool take = D < Db;
            if (D == Db && i2 != ib) {  // exact tie between different points: the lower ORIGINAL index
                const int oc = p.perm_t != nullptr ? p.perm_t[min(i2, p.nt - 1)] : i2;
                const int ob = p.perm_t != nullptr ? p.perm_t[min(ib, p.nt - 1)] : ib;
                take = oc < ob;
            }
            if (take) { Db = D; ib = i2; }
        }
        Dbest_l[r0] = Db;
        ibest_l[r0] = ib;
        if (lane == 0) p.tp_tick[slot] = 0;
    }
    if (FUSED) {
        // ---- correspondences -> per-row sums (icp_device.cuh) --------------------------------------
        // Distances are re-derived in FP64 from the ORIGINAL source point, the FP64 pose and the
        // neighbour's original float32 coordinates (one 16-byte gather; neighbours of consecutive
        // stored queries are stored close together): the strict d2 < max_d2 test, fitness, rmse
        // and the Kabsch sums carry no FP32 error, only the choice of neighbour was made in FP32.
        static_assert(!FUSED || (offsetof(WS, sph) == sizeof(WS::buf) && offsetof(WS, id) == offsetof(WS, sph) + sizeof(WS::sph) &&
                                 offsetof(WS, rows) == offsetof(WS, id) + sizeof(WS::id) &&
                                 offsetof(WS, box) == offsetof(WS, rows) + sizeof(WS::rows) &&
                                 offsetof(WS, box) + sizeof(WS::box) >= 32 * kNS * sizeof(double)),
                      "the row reduction borrows the (idle) ring buffer and FIFO arrays");
        double *scratch = reinterpret_cast<double *>(&ws.buf[0][0][0]);  // [32 lanes][17]
        __syncwarp();
        double Te[12];  // (re-read: the pose must not sit in registers across the search)
#pragma unroll
        for (int k = 0; k < 12; ++k) Te[k] = __ldcg(&f.states[b].T[k]);
        double rs[Q];
#pragma unroll 1
        for (int r = 0; r < Q; ++r) rs[r] = 0.0;
#pragma unroll 1
        for (int r = 0; r < Q; ++r) {
            if (!((own >> r) & 1u)) continue;  // warp-uniform
            double c[kNS];
#pragma unroll
            for (int k = 0; k < kNS; ++k) c[k] = 0.0;
            if ((livemask >> r) & 1u) {
                const int i = q0 + r * 32;
                const int j = ibest_l[r];
                p.hint[(long long)b * p.nq_pad + i] = j;
                const float4 t4 = f.tgt4[j];
                const float *s7 = f.src7;
                const double px = (double)s7[i] + (double)s7[4ll * p.nq_pad + i];
                const double py = (double)s7[p.nq_pad + i] + (double)s7[5ll * p.nq_pad + i];
                const double pz = (double)s7[2ll * p.nq_pad + i] + (double)s7[6ll * p.nq_pad + i];
                double sx, sy, sz;
                icp_apply_pose(Te, px, py, pz, sx, sy, sz);
                const double d2 = icp_dist2(sx, sy, sz, (double)t4.x, (double)t4.y, (double)t4.z);
                const bool in = d2 < f.max_d2;
                f.inlier[(long long)b * p.nq_pad + i] = in ? 1 : 0;
                if (in) icp_contrib(sx, sy, sz, (double)t4.x, (double)t4.y, (double)t4.z, d2, c);
            }
#pragma unroll
            for (int k = 0; k < kNS; ++k) scratch[lane * kNS + k] = c[k];
            __syncwarp();
            rs[r] = icp_row_sum(scratch, lane);  // lane k: component k over the row's 32 queries
            __syncwarp();
        }
#ifdef ISR_PHASE_LOG
        const long long ph_e = clock64() - t_start;
        if (p.cta_log != nullptr && lane == 0) atomicMax(p.cta_log + 4 * (p.cta_log_cap - 1) + 2, global_ns());
#endif
        icp_fused_tail(f, b, blk, own, lane, rs);
#ifdef ISR_PHASE_LOG
        if (p.cta_log != nullptr && lane == 0) {
            atomicMax(p.cta_log + 4 * (p.cta_log_cap - 1) + 3, global_ns());
            const long long rec = ((long long)b * gridDim.x + blockIdx.x);
            if (rec < p.cta_log_cap - 1) {
                unsigned long long *o = p.cta_log + 4 * rec;
                o[0] = (unsigned long long)ph_s;
                o[1] = (unsigned long long)ph_q | ((unsigned long long)ph_h << 32);
                o[2] = (unsigned long long)ph_r | ((unsigned long long)ph_e << 32);
                o[3] = ((unsigned long long)(unsigned)blk << 32) | ((unsigned long long)own << 24) |
                       ((unsigned long long)tpart << 20) | (unsigned long long)((clock64() - t_start) >> 8 & 0xFFFFF);
                unsigned long long *o2 = p.cta_log + 4 * (p.cta_log_cap / 2 + rec);  // main-loop shares
                if (rec < p.cta_log_cap / 2 - 1) {
                    o2[1] = (unsigned long long)acc_scan | ((unsigned long long)acc_resolve << 32);
                    o2[0] = (unsigned long long)acc_test | ((unsigned long long)acc_wait << 32);
                    o2[2] = (unsigned long long)acc_produce | ((unsigned long long)acc_sort << 32);
                    o2[3] = ((unsigned long long)nscanned << 48) | ((unsigned long long)ntests << 32) |
                            ((unsigned long long)npass << 16) | (unsigned long long)nhalves;
                }
            }
        }
#endif
    } else {
#pragma unroll 1
        for (int r = 0; r < Q; ++r) {
            const int i = q0 + r * 32;
            if ((livemask >> r) & 1u) {
                const int io = p.perm_q != nullptr ? p.perm_q[i] : i;
                const int jo = p.perm_t != nullptr ? p.perm_t[min(ibest_l[r], p.nt - 1)] : ibest_l[r];
                const long long o = (long long)b * p.nq + io;
                p.out_d2[o] = (float)Dbest_l[r];
                if (p.out_idx != nullptr) p.out_idx[o] = jo;
                if (p.hint != nullptr) p.hint[(long long)b * p.nq_pad + i] = ibest_l[r];
            }
        }
#ifdef ISR_PHASE_LOG
        if (p.cta_log != nullptr && lane == 0) {
            const long long rec = ((long long)b * gridDim.x + blockIdx.x);
            if (rec < p.cta_log_cap / 2 - 1) {
                unsigned long long *o = p.cta_log + 4 * rec;
                o[0] = (unsigned long long)ph_s;
                o[1] = (unsigned long long)ph_q | ((unsigned long long)ph_h << 32);
                o[2] = (unsigned long long)ph_r | ((unsigned long long)ph_m << 32);
                o[3] = ((unsigned long long)(unsigned)blk << 32) | ((unsigned long long)own << 24) |
                       (unsigned long long)((clock64() - t_start) >> 8 & 0xFFFFF);
                unsigned long long *o2 = p.cta_log + 4 * (p.cta_log_cap / 2 + rec);
                o2[0] = (unsigned long long)acc_test | ((unsigned long long)acc_wait << 32);
                o2[1] = (unsigned long long)acc_scan | ((unsigned long long)acc_resolve << 32);
                o2[2] = (unsigned long long)acc_produce | ((unsigned long long)acc_sort << 32);
                o2[3] = ((unsigned long long)nscanned << 48) | ((unsigned long long)ntests << 32) |
                        ((unsigned long long)npass << 16) | (unsigned long long)nhalves;
            }
        }
#endif
    }
}

// ---- launch order of the pruned kernel's query blocks -----------------------------------------
// Work per block grows with the spread of its queries (a block that straddles a gap of the
// curve needs the neighbourhoods of both sides).  weight = squared radius of the block about
// its centroid; blocks are launched in descending weight (longest-processing-time first).
constexpr int kOrderMax = 4096;  // blocks per batch item that the single-CTA sort handles

// one warp per (batch item, block of QB queries)
template <int QB>
__global__ void __launch_bounds__(128)
block_weight_kernel(const float *__restrict__ q, long long q_bstride, int nq, int nq_pad, int nqb,
                    int batch, u64 *__restrict__ keys, float *__restrict__ rowrad) {
    const int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= nqb * batch) return;
    const int b = w / nqb, blk = w % nqb;
    const float *gq = q + (long long)b * q_bstride;
    float x[QB / 32], y[QB / 32], z[QB / 32];
    float sx = 0.f, sy = 0.f, sz = 0.f, sn = 0.f;
#pragma unroll
    for (int r = 0; r < QB / 32; ++r) {
        const int i = blk * QB + r * 32 + lane;
        const bool live = i < nq;
        const int ii = min(i, nq_pad - 1);
        x[r] = gq[ii]; y[r] = gq[nq_pad + ii]; z[r] = gq[2ll * nq_pad + ii];
        if (live) { sx += x[r]; sy += y[r]; sz += z[r]; sn += 1.f; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
        sn += __shfl_xor_sync(0xffffffffu, sn, o);
    }
    const float inv = sn > 0.f ? 1.f / sn : 0.f;
    sx *= inv; sy *= inv; sz *= inv;
    float m = 0.f;
#pragma unroll
    for (int r = 0; r < QB / 32; ++r) {
        if (blk * QB + r * 32 + lane < nq) {
            const float dx = x[r] - sx, dy = y[r] - sy, dz = z[r] - sz;
            m = fmaxf(m, dx * dx + dy * dy + dz * dz);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // ascending sort of ~bits(weight) = descending weight; ties by ascending block index
    if (lane == 0) keys[w] = ((u64)(~__float_as_uint(m)) << 32) | (u64)(unsigned)blk;
    if (rowrad != nullptr) {  // squared radius of every query row about its own centroid (target parts)
#pragma unroll 1
        for (int r = 0; r < QB / 32; ++r) {
            const bool live = blk * QB + r * 32 + lane < nq;
            float cx = live ? x[r] : 0.f, cy = live ? y[r] : 0.f, cz = live ? z[r] : 0.f, cn = live ? 1.f : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                cx += __shfl_xor_sync(0xffffffffu, cx, o);
                cy += __shfl_xor_sync(0xffffffffu, cy, o);
                cz += __shfl_xor_sync(0xffffffffu, cz, o);
                cn += __shfl_xor_sync(0xffffffffu, cn, o);
            }
            const float in = cn > 0.f ? 1.f / cn : 0.f;
            const float dx = x[r] - cx * in, dy = y[r] - cy * in, dz = z[r] - cz * in;
            float rm = live ? dx * dx + dy * dy + dz * dz : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, o));
            if (lane == 0) rowrad[(long long)w * (QB / 32) + r] = rm;
        }
    }
}

// one CTA per batch item: bitonic sort of its <= kOrderMax keys; then the launch list:
// the (at most kSplitMax) blocks whose radius exceeds twice the median radius -- blocks that
// straddle a gap of the curve and need the neighbourhoods of several patches -- are split
// into Q single-row entries, everything else follows as whole blocks in descending weight.
constexpr int kSplitMax = 128;  // (512 was measured slower: 25 % more total work, tail no longer the limit)
__global__ void __launch_bounds__(1024)
block_order_kernel(const u64 *__restrict__ keys, int nqb, int rows, int stride, int split_max, int parts,
                   float split_factor, int *__restrict__ order, int *__restrict__ order_count,
                   const float *__restrict__ rowrad, float tp_factor, int tp_slots, int *__restrict__ order_slot,
                   unsigned *__restrict__ tp_tick, int mix_slots) {
    __shared__ u64 s[kOrderMax];
    __shared__ int nsplit, nsplit_entries, nmixed;
    __shared__ short rowpos[kSplitMax * 8];   // first launch-list entry of row r of split block i
    __shared__ signed char rowslot[kSplitMax * 8];  // its merge slot when it runs as target parts, else -1
    static_assert(kSplitMax == 128, "order_workspace_bytes sizes the launch list for 128 split blocks");
    const int b = blockIdx.x;
    int n2 = 1;
    while (n2 < nqb) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += 1024) s[i] = i < nqb ? keys[(long long)b * nqb + i] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n2 / 2; t += 1024) {
                const int i = 2 * t - (t & (j - 1));
                const bool asc = (i & k) == 0;
                const u64 a = s[i], c = s[i + j];
                if ((a > c) == asc) { s[i] = c; s[i + j] = a; }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        const float wmed = __uint_as_float(~(unsigned)(s[nqb / 2] >> 32));
        int h = 0;
        while (h < split_max && h < nqb && __uint_as_float(~(unsigned)(s[h] >> 32)) > split_factor * wmed) ++h;
        nsplit = h;
        // rows of the split blocks that are themselves several patches far apart (a jump of the curve
        // inside the row) run as kTargetParts CTAs over disjoint shares of the target
        int pos = rowrad != nullptr ? 0 : h * rows, used = 0;
        for (int i = 0; i < h && rowrad != nullptr; ++i) {
            const int blk = (int)(unsigned)(s[i] & 0xffffffffull);
            for (int r = 0; r < rows; ++r) {
                const bool tp = used < tp_slots && rowrad[((long long)b * nqb + blk) * rows + r] > tp_factor * wmed;
                rowpos[i * rows + r] = (short)pos;
                rowslot[i * rows + r] = tp ? (signed char)used : (signed char)-1;
                if (tp && tp_tick != nullptr) tp_tick[used] = 0;
                pos += tp ? kTargetParts : 1;
                used += tp ? 1 : 0;
            }
        }
        nsplit_entries = pos;
        // A launch that leaves resident slots empty (mix_slots > 0: a one-wave grid) runs the widest
        // of the remaining blocks as twice as many CTAs, as far as the slots go
        int nm = 0;
        if (mix_slots > 0 && parts < rows) {
            nm = (mix_slots - pos - parts * (nqb - h)) / parts;
            nm = nm < 0 ? 0 : (nm > nqb - h ? nqb - h : nm);
        }
        nmixed = nm;
        order_count[b] = parts * (nqb - h) + parts * nm + pos;
    }
    __syncthreads();
    const int h = nsplit, hbase = nsplit_entries, nm = nmixed;
    int *out = order + (long long)b * stride;
    for (int i = threadIdx.x; i < nqb; i += 1024) {
        const int blk = (int)(unsigned)(s[i] & 0xffffffffull);
        if (i < h) {
            for (int r = 0; r < rows; ++r) {
                const int at = rowrad != nullptr ? rowpos[i * rows + r] : i * rows + r;
                const int sl = rowrad != nullptr ? rowslot[i * rows + r] : -1;
                if (sl < 0) {
                    out[at] = launch_entry(blk, 1u << r, 0);
                } else {
                    for (int k = 0; k < kTargetParts; ++k) {
                        out[at + k] = launch_entry(blk, 1u << r, k + 1);
                        order_slot[at + k] = sl;
                    }
                }
            }
        } else {
            // a grid that does not fill the machine: every block runs as `parts` CTAs of 8 / parts
            // query rows (the scan skips the rows a CTA does not own, so the split costs little
            // and shortens the tail)
            if (i - h < nm) {
                for (int c = 0; c < 2 * parts; ++c)
                    out[hbase + 2 * parts * (i - h) + c] = launch_entry(blk, rows_of_part(2 * parts, c), 0);
            } else {
                for (int c = 0; c < parts; ++c)
                    out[hbase + parts * nm + parts * (i - h) + c] = launch_entry(blk, rows_of_part(parts, c), 0);
            }
        }
    }
}

// ---- the launch list re-cut from measured costs (fused ICP, single start, one-wave grids) -------
// An ICP run repeats nearly the same search 30-50 times, and in a one-wave grid an iteration lasts
// as long as its slowest CTA (measured on a 1/8 shard of 1M points: median 121 k cycles, slowest
// 228 k, with every slot of the machine taken).  Block radii predict the cost of a CTA only roughly;
// the previous iteration MEASURED it (NN2Params::cost).  This kernel turns the measurements into a
// variable cost per query row, v_r = (cycles - F) / rows of the CTA (F: the fixed cost of a CTA, the
// cheapest one seen), and cuts every block again: a run of rows stays one CTA while its estimate
// F + sum v_r stays under a level T, else it is halved (8 -> 4 -> 2 -> 1 rows, a single row -> target
// parts); T is the lowest level (bisection) whose CTA count still fits the resident slots.  Light
// blocks merge, heavy ones split, and the list is re-sorted heaviest-first.  The reduction tree of
// the sums does not depend on how blocks are cut into CTAs (icp_fused_tail), so the iteration's
// result is bit-identical before and after.
#ifndef ISR_SPLIT_INFLATE
#define ISR_SPLIT_INFLATE 1.15f
#endif
constexpr float kSplitInflate = ISR_SPLIT_INFLATE;  // a halved run costs more than half (repeated coarse walk and tests)
// cost of a row inside a run of rows_new rows, measured inside a CTA of rows_old rows
__device__ __forceinline__ float rebalance_scale(int rows_new, int rows_old) {
    float f = 1.f;
    for (int a = rows_old; a > rows_new; a >>= 1) f *= kSplitInflate;
    // (merged runs are taken at the sum of their parts: a row that straddles a jump of the curve is far
    // cheaper on its own -- its CTA's coarse tests see one small sphere -- than inside a wider CTA)
    return f;
}
// The CTAs of one block at level T: a run of rows stays one CTA while its estimate F + sum v_r stays
// under T, else it is halved (8 -> 4 -> 2 -> 1 rows; a single row above T runs as target parts).
// emit(rows mask, CTAs (1 or kTargetParts), estimate).
// (Runs of any length, packed greedily under T, were measured too: 1 882 single-row CTAs, the same
// slowest CTA, iteration 4 % slower -- the estimates, not the granularity, limit the balance.)
// (merge: may rows that were measured in narrower CTAs be joined?  Only a one-wave list needs that, to
// free slots for the cuts)
template <class Emit>
__device__ __forceinline__ void rebalance_block(const float *v, const unsigned char *g, float F, float T,
                                                bool allow_tp, bool merge, Emit emit) {
    auto est = [&](int r0, int n) {
        float sum = 0.f;
        for (int r = r0; r < r0 + n; ++r) {
            sum += v[r] * rebalance_scale(n, g[r]);
            if (!merge && g[r] < n) sum = CUDART_INF_F;
        }
        return F + sum;
    };
    const float e8 = est(0, 8);
    if (e8 <= T) { emit(0xFFu, 1, e8); return; }
    for (int h = 0; h < 2; ++h) {
        const float e4 = est(4 * h, 4);
        if (e4 <= T) { emit(0x0Fu << (4 * h), 1, e4); continue; }
        for (int q = 2 * h; q < 2 * h + 2; ++q) {
            const float e2 = est(2 * q, 2);
            if (e2 <= T) { emit(0x03u << (2 * q), 1, e2); continue; }
            for (int r = 2 * q; r < 2 * q + 2; ++r) {
                const float e1 = est(r, 1);
                if (e1 <= T || !allow_tp) { emit(1u << r, 1, e1); continue; }
                emit(1u << r, kTargetParts, F + (e1 - F) * (1.1f / kTargetParts) + 0.15f * F);
            }
        }
    }
}
// the eight row costs and measured CTA widths of one block, in three vector loads
struct RebalanceRows {
    float v[8];
    unsigned char g[8];
};
__device__ __forceinline__ RebalanceRows rebalance_load(const float *vrow, const unsigned char *grow, int b) {
    RebalanceRows o;
    const float4 a = *reinterpret_cast<const float4 *>(vrow + b * 8), c = *reinterpret_cast<const float4 *>(vrow + b * 8 + 4);
    const uint2 w = *reinterpret_cast<const uint2 *>(grow + b * 8);
    o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = c.x; o.v[5] = c.y; o.v[6] = c.z; o.v[7] = c.w;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        o.g[r] = (unsigned char)((w.x >> (8 * r)) & 0xFFu);
        o.g[4 + r] = (unsigned char)((w.y >> (8 * r)) & 0xFFu);
    }
    return o;
}
__global__ void __launch_bounds__(1024)
block_rebalance_kernel(int *__restrict__ order, int *__restrict__ order_slot, int *__restrict__ order_count,
                       const unsigned *__restrict__ cost, float *__restrict__ vrow, unsigned char *__restrict__ grow,
                       int nqb, int slots, int max_entries, int tp_slots, unsigned *__restrict__ tp_tick,
                       float deep_level) {
    constexpr int kBins = 256;  // launch order: descending estimate, in this many levels
    __shared__ unsigned fmin_s, n_s, tp_s, bin_n[kBins], bin_at[kBins];
    __shared__ float top_s, sum_s;
    const int tid = threadIdx.x;
    const int n_old = order_count[0];
    if (tid == 0) { fmin_s = 0x7FFFFFFFu; top_s = 0.f; sum_s = 0.f; }
    __syncthreads();
    // F: the cheapest CTA that searched at all
    for (int e = tid; e < n_old; e += 1024) {
        const unsigned c = cost[e];
        if (c > 0u) atomicMin(&fmin_s, c);
    }
    for (int i = tid; i < nqb * 8; i += 1024) { vrow[i] = 0.f; grow[i] = 8; }
    __syncthreads();
    const float F = fmin_s == 0x7FFFFFFFu ? 1.f : (float)fmin_s;
    // variable cost per query row (the parts of a target-part row add up)
    float csum = 0.f;
    for (int e = tid; e < n_old; e += 1024) {
        const int entry = order[e];
        const int blk = entry & 0xFFFFF;
        const unsigned rows = ((unsigned)entry >> 20) & 0xFFu;
        const int nrows = __popc(rows);
        const float c = (float)cost[e];
        const float share = fmaxf(c - F, 0.f) / (float)nrows;
        const bool part = ((entry >> 28) & 7) != 0;  // (only the parts of a target-part row share a row)
        for (int r = 0; r < 8; ++r)
            if ((rows >> r) & 1u) {
                if (part) atomicAdd(&vrow[blk * 8 + r], share); else vrow[blk * 8 + r] = share;
                grow[blk * 8 + r] = (unsigned char)nrows;
            }
        csum += c;
    }
    csum = (float)warp_sum((double)csum);
    if ((tid & 31) == 0) atomicAdd(&sum_s, csum);
    __syncthreads();
    // the heaviest whole block: upper end of the bisection and of the launch-order levels
    for (int b = tid; b < nqb; b += 1024) {
        float sum = 0.f;
        for (int r = 0; r < 8; ++r) sum += vrow[b * 8 + r];  // (deflated estimates are below this)
        atomicMax(reinterpret_cast<int *>(&top_s), __float_as_int(F + sum));
    }
    __syncthreads();
    const float top = top_s * 1.0001f + 1.f;
    float level = top;
    if (nqb <= slots) {
        // one wave: the lowest level whose CTAs are all resident at once
        float lo = F;
        const int cap = min(slots, max_entries);
        for (int it = 0; it < 12; ++it) {  // (to 2^-12 of the range: a few dozen cycles)
            const float T = 0.5f * (lo + level);
            if (tid == 0) { n_s = 0; tp_s = 0; }
            __syncthreads();
            unsigned n = 0, ntp = 0;
            for (int b = tid; b < nqb; b += 1024) {
                const RebalanceRows R = rebalance_load(vrow, grow, b);
                rebalance_block(R.v, R.g, F, T, tp_slots > 0, true, [&](unsigned, int parts, float) {
                    n += (unsigned)parts;
                    ntp += parts > 1 ? 1u : 0u;
                });
            }
            atomicAdd(&n_s, n);
            atomicAdd(&tp_s, ntp);
            __syncthreads();
            const bool ok = n_s <= (unsigned)cap && tp_s <= (unsigned)tp_slots;
            __syncthreads();
            if (ok) level = T; else lo = T;
        }
    } else {
        // several waves, launched heaviest first: only a CTA that outlasts the average load of a slot
        // can be the tail -- those are cut, everything else keeps its shape
        level = fmaxf(deep_level * sum_s / (float)slots, 2.f * F);
    }
    // launch list at that level, heaviest first (counting sort over kBins levels of the estimate)
    if (tid == 0) { n_s = 0; tp_s = 0; }
    for (int i = tid; i < kBins; i += 1024) { bin_n[i] = 0; bin_at[i] = 0; }
    __syncthreads();
    const float bin_scale = (float)(kBins - 1) / fmaxf(top - F, 1.f);
    auto bin_of = [&](float e) {
        const int k = (int)(fminf(fmaxf(e - F, 0.f) * bin_scale, (float)(kBins - 1)));
        return kBins - 1 - k;
    };
    for (int b = tid; b < nqb; b += 1024) {
        const RebalanceRows R = rebalance_load(vrow, grow, b);
        rebalance_block(R.v, R.g, F, level, tp_slots > 0, nqb <= slots,
                        [&](unsigned, int parts, float e) { atomicAdd(&bin_n[bin_of(e)], (unsigned)parts); });
    }
    __syncthreads();
    if (tid == 0) {
        unsigned at = 0;
        for (int i = 0; i < kBins; ++i) { bin_at[i] = at; at += bin_n[i]; }
        n_s = at;
    }
    __syncthreads();
    const int n_new = (int)n_s;
    if (n_new > max_entries) return;  // (cannot happen: the list has room for every row of every block)
    for (int b = tid; b < nqb; b += 1024) {
        const RebalanceRows R = rebalance_load(vrow, grow, b);
        rebalance_block(R.v, R.g, F, level, tp_slots > 0, nqb <= slots, [&](unsigned rows, int parts, float e) {
            const unsigned at = atomicAdd(&bin_at[bin_of(e)], (unsigned)parts);
            const unsigned slot = parts > 1 ? atomicAdd(&tp_s, 1u) : 0u;
            for (int k = 0; k < parts; ++k) {
                order[at + k] = launch_entry(b, rows, parts > 1 ? k + 1 : 0);
                order_slot[at + k] = (int)slot;
            }
        });
    }
    for (int i = tid; i < tp_slots; i += 1024) tp_tick[i] = 0u;
    if (tid == 0) order_count[0] = n_new;
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
template <int Q, int THREADS, int STAGE, int NSTAGES, int SUB, int MINB, int UNR = 2>
struct NN2Variant {
    static constexpr int kQueriesPerCta = Q * THREADS;
    static constexpr int kStage = STAGE;
    static constexpr bool kPrune = false;
    static constexpr size_t kSmem = (size_t)NSTAGES * 4 * STAGE * 4 + NSTAGES * 8;

    static constexpr bool kFused = false;
    static constexpr bool kSplit = false;
    static int launch(const NN2Params &p, dim3 grid, cudaStream_t st, const IcpFuse &, bool) {
        auto kern = nn2_kernel<Q, THREADS, STAGE, NSTAGES, SUB, MINB, UNR>;
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

    static int ctas_per_sm(bool) {
        int n = 0;
        auto kern = nn2_kernel<Q, THREADS, STAGE, NSTAGES, SUB, MINB, UNR>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, THREADS, kSmem) != cudaSuccess ||
            n < 1)
            n = 1;
        return n;
    }
};

using NN2Main = NN2Variant<8, 128, 1024, 3, 64, 4, 1>;

// the pruned kernel: WARPS independent warps of 32 x Q queries per CTA
template <int Q, int WARPS, int SUB, int MINB, int UNR, int FLAG, int PARTS, int GROUPS = 2, bool FUSED = false,
          bool SPLIT = false>
struct NN2PrunedVariant {
    static constexpr int kQueriesPerCta = Q * WARPS * 32;
    static constexpr int kStage = ISR_SOA_TILE;
    static constexpr bool kPrune = true;
    static constexpr bool kFused = FUSED;
    static constexpr bool kSplit = SPLIT;
    static constexpr size_t kSmem =
        (size_t)WARPS * sizeof(PrunedWarpSmem<SUB, Q, SPLIT, FUSED ? kRing : ISR_NN_RING_PLAIN, FUSED ? kFifo : ISR_NN_FIFO_PLAIN>);

    static int launch(const NN2Params &p, dim3 grid, cudaStream_t st, const IcpFuse &fuse, bool pdl) {
        auto kern = nn2_pruned_kernel<Q, WARPS, SUB, MINB, UNR, FLAG, PARTS, GROUPS, FUSED, SPLIT>;
        static thread_local int configured_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_dev != dev) {
            // MINB single-warp CTAs of ~10 KB each only fit with the largest shared-memory carve-out
            ISR_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                    (int)cudaSharedmemCarveoutMaxShared),
                               "nn2 pruned carve-out"));
            configured_dev = dev;
        }
        ProfScope prof(kProfNN, st);
        if (FUSED && pdl) {
            // the previous operation of the stream is the previous iteration's launch of this kernel
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = grid;
            cfg.blockDim = dim3(WARPS * 32);
            cfg.dynamicSmemBytes = kSmem;
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            ISR_TRY(check_cuda(cudaLaunchKernelEx(&cfg, kern, p, fuse), "nn2 fused launch (dependent)"));
        } else {
            kern<<<grid, WARPS * 32, kSmem, st>>>(p, fuse);
        }
        return launched(FUSED ? "nn2_pruned_kernel<fused icp>" : "nn2_pruned_kernel");
    }
    // resident CTAs per SM: MINB by registers, and what shared memory allows
    static int ctas_per_sm(bool) {
        const size_t smem = kSmem + 1024;
        const int by_smem = (int)((size_t)233472 / smem);
        return by_smem < MINB ? by_smem : MINB;
    }
};
// one warp per CTA: a finished warp frees its slot at once (measured 5 % faster than 4-warp
// CTAs, whose slowest warp holds the registers and shared memory of the other three)
// (flag granularity FLAG = 32 / 16 targets inside the 64-target pruning unit was measured
// 2 % / 19 % slower than 64: more, shorter resolve passes)
// (20 warps per SM -- a 102-register cap, 224 B of spills -- was measured 2 % slower on the
// verification workload and 8 % slower on the ICP search)
// (PARTS = 2 / 4 -- flagging halves / quarters of the sub-tile so that the resolve re-derives
// only those -- removes 3/4 of the resolve's pass-1 instructions but was measured 2 % slower
// on the verification workload and equal on the ICP search: the kernel is bound by the
// latency of its serial phases at 4 warps per scheduler, not by instruction count)
// (GROUPS: the scan runs per quarter of the query rows (2 rows); halves (4 rows) evaluate 25 %
// more pairs and were 2 % slower on verification and ICP, 9 % on ADD-S with its sparse queries;
// single rows evaluate 6 % fewer pairs still but leave one dependent FFMA2 chain per lane: -7 %)
// (but a scan that needs most rows costs ~40 % more in two-row pieces -- twice the shared-memory
// reads per FFMA2, two dependent chains per lane instead of four: the multi-start ICP batch,
// whose starts are mostly far from aligned, took 0.40 s with quarters against 0.27 s with
// halves, so a batched ICP search uses the halves; choosing per half at run time -- 4-row scan
// when both quarters are needed -- kept config 5 at 0.30 s but gave the 4 % on verification back)
// (round 2 re-measured PARTS on the verification workload, same box: 1 -> 10 650, 2 -> 10 500, 4 -> 10 470
// candidates/s although PARTS = 4 removes 18 % of all instructions: pass 1 is 16 independent
// groups that issue back to back; what a warp waits for are its dependent chains elsewhere)
#ifndef ISR_NN_VERIFY_PARTS
#define ISR_NN_VERIFY_PARTS 1
#endif
using NN2Pruned = NN2PrunedVariant<8, 1, 64, ISR_NN_MINB_PLAIN, 1, 64, ISR_NN_VERIFY_PARTS, 4>;
using NN2PrunedHalves = NN2PrunedVariant<8, 1, 64, ISR_NN_MINB_PLAIN, 1, 64, 1, 2>;
using NN2PrunedFused = NN2PrunedVariant<8, 1, 64, ISR_NN_MINB_FUSED, 1, 64, 1, 4, true>;
using NN2PrunedHalvesFused = NN2PrunedVariant<8, 1, 64, ISR_NN_MINB_FUSED, 1, 64, 1, 2, true>;
// (flags per quarter of the sub-tile, so that a resolve pass re-derives 16 filter values instead of
// 64: in a shallow grid a warp's instruction count IS its latency -- 0.161 -> 0.155 ms per
// iteration on a 1/8 shard; in deep grids the same change was measured 0 .. -2 %)
#ifndef ISR_SPLIT_UNR
#define ISR_SPLIT_UNR 1
#endif
using NN2PrunedFusedSplit = NN2PrunedVariant<8, 1, 64, ISR_NN_MINB_FUSED, ISR_SPLIT_UNR, 64, 4, 4, true, true>;
#ifdef ISR_NN_TUNING
using NN2PrunedP2 = NN2PrunedVariant<8, 1, 64, 20, 1, 64, 2>;
using NN2PrunedP4 = NN2PrunedVariant<8, 1, 64, 20, 1, 64, 4>;
using NN2PrunedU2 = NN2PrunedVariant<8, 1, 64, 16, 2, 64, 1>;
using NN2PrunedU4 = NN2PrunedVariant<8, 1, 64, 16, 4, 64, 1>;
#endif
static_assert(NN2Pruned::kSmem <= 48 * 1024, "pruned kernel uses the default dynamic shared memory limit");
constexpr int kMaxSplits = 32;

// process-wide pruning switch and profiling counters (isr.h)
static std::atomic<int> g_prune{-1};
static std::atomic<unsigned long long> g_answered{0};
static unsigned long long *g_cta_log = nullptr;  // isr_debug_cta_log
static long long g_cta_log_cap = 0;
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

constexpr size_t kTargetPartBytes =  // row radii are sized per call; these are the fixed pieces
    (size_t)kTargetSlots * kTargetParts * 32 * (8 + 4) + (size_t)kTargetSlots * 4;
static size_t order_workspace_bytes(long long nqb, long long batch) {
    // keys, launch list (room for every block as 8 single-row entries, plus the extra entries of
    // rows that run as target parts), entry counts; for target parts: merge slots of the entries,
    // row radii, merge buffers and tickets
    const size_t list = (size_t)(8 * nqb + kTargetParts * kTargetSlots) * batch * 4;
    // (+ measured cost per launch-list entry and the rows-per-CTA of every row: block_rebalance_kernel)
    return align256((size_t)nqb * batch * 8) + 2 * align256(list) + align256((size_t)batch * 4) +
           align256((size_t)nqb * batch * 8 * 4) + align256(kTargetPartBytes) + align256(list) +
           align256((size_t)nqb * batch * 8);
}

// run-time tuning knob, read once: ISR_<name> in the environment, else the default
static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e != nullptr && *e != 0 ? atoi(e) : dflt;
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
    int reuse_order;
    const IcpFuse *fuse;  // the fused ICP iteration (icp_device.cuh), or NULL
};

template <class V>
static int nn2_dispatch(const NN2Call &c) {
    const int64_t nq = c.q->n;
    const int nqb = (int)((nq + V::kQueriesPerCta - 1) / V::kQueriesPerCta);
    const int stages = (int)(c.t->npad / V::kStage);
    const int slots = sm_count() * V::ctas_per_sm(c.use_lo != 0);  // (per call: the device may differ)
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
    p.stage_c_bstride = c.t->bstride == 0 ? 0 : isr_stage_sphere_count(c.t->npad);
    p.perm_q = c.q->perm; p.perm_t = c.t->perm;
    p.sub_c = reinterpret_cast<const float4 *>(c.t->sub_c);
    p.sub_c_bstride = c.t->bstride == 0 ? 0 : c.t->npad / ISR_SUB_TILE;
    p.sub_h = reinterpret_cast<const unsigned *>(c.t->sub_box);
    p.sub_h_bstride = p.sub_c_bstride;
#ifdef ISR_NN_TUNING
    if (const char *e = getenv("ISR_NN_BOX")) { if (atoi(e) == 0) p.sub_h = nullptr; }
#endif
    p.cta_log = g_cta_log;
    p.cta_log_cap = g_cta_log_cap;
    p.evaluated = nullptr;
    if (prof_enabled()) {
        p.evaluated = evaluated_counter();
        g_answered.fetch_add((unsigned long long)nq * (unsigned long long)c.t->n *
                             (unsigned long long)c.batch);
    }
    p.hint = V::kPrune ? c.q->hint : nullptr;
    p.nanchors = kAnchors;  // measured on the verification workload: 1 seed 6.4k, 2 7.6k, 4 8.8k, 8 9.0k candidates/s
#ifdef ISR_NN_TUNING
    if (const char *e = getenv("ISR_NN_ANCHORS")) p.nanchors = atoi(e);
#endif
    p.sort_fifo = 1;  // measured on the verification workload: +3 %
#ifdef ISR_NN_TUNING
    if (const char *e = getenv("ISR_NN_SORT")) p.sort_fifo = atoi(e);
#endif
    p.order = nullptr;
    p.order_count = nullptr;
    p.order_slot = nullptr; p.tp_D = nullptr; p.tp_I = nullptr; p.tp_tick = nullptr;
    p.cost = nullptr;
    int grid_x = nqb;
    if (V::kPrune && nqb > 1 && nqb <= kOrderMax && c.workspace != nullptr &&
        c.workspace_bytes >= order_workspace_bytes(nqb, c.batch)) {
        constexpr int kRows = V::kQueriesPerCta / 32;
        // splitting only pays when the grid is a few waves deep (a single cloud pair, a few
        // ICP starts); a big batch hides its wide blocks behind the others
        int split_max = (long long)nqb * c.batch <= 16384 ? kSplitMax : 0;
        float split_factor = 4.0f;  // squared radius > 4 x the median's: radius > twice the median
#ifdef ISR_NN_TUNING
        if (const char *e = getenv("ISR_NN_SPLIT_MAX")) { if (split_max) split_max = atoi(e); }
        if (const char *e = getenv("ISR_NN_SPLIT_THR")) split_factor = (float)atof(e);
#endif
        // a grid that does not fill the machine runs every block as 2 / 4 / 8 CTAs of 4 / 2 / 1
        // query rows, as long as all of them are resident at once (round 1, halves only, ICP
        // search: 100k points 0.203 -> 0.163 ms, 250k 0.248 -> 0.206 ms; from one full wave on --
        // 500k, 1M points -- the repeated per-CTA tests cost more than the shorter tail gains:
        // 0.350 -> 0.381 ms, 0.63 -> 0.81 ms)
        static const int parts_max = env_int("ISR_NN_PARTS_MAX", 8);
        int parts = 1;
        while (parts < parts_max && 2ll * parts * nqb * c.batch <= slots) parts *= 2;
        { const int force = env_int("ISR_NN_PARTS_FORCE", 0); if (force > 0) parts = force; }
        if (split_max > nqb) split_max = nqb;
        // target parts: only in the split instantiation of the fused ICP iteration (one start,
        // shallow grid), where the slowest single row is the iteration
        // (read per call, not cached: the tests switch them within one process)
        const int tp_slots_env = env_int("ISR_NN_TP_SLOTS", kTargetSlots);
        const float tp_factor = (float)env_int("ISR_NN_TP_FACTOR_X10", 40) * 0.1f;
        const int tp_slots = V::kFused && V::kSplit && c.batch == 1 && split_max > 0
                                 ? (tp_slots_env < kTargetSlots ? tp_slots_env : kTargetSlots) : 0;
        // (mixing: only the fused single-start iteration, whose launch list is built once per run)
        const int mix_env = env_int("ISR_NN_MIX", 1);
        const int base_ctas = parts * nqb + (kRows - parts) * split_max + (kTargetParts - 1) * tp_slots;
        const int mix_slots = V::kFused && V::kSplit && c.batch == 1 && mix_env != 0 && parts < kRows &&
                                      (long long)parts * nqb <= slots ? slots : 0;
        const int stride = mix_slots > base_ctas ? mix_slots : base_ctas;
        char *w = reinterpret_cast<char *>(c.workspace);
        const size_t list = align256((size_t)(8 * nqb + kTargetParts * kTargetSlots) * c.batch * 4);
        u64 *keys = reinterpret_cast<u64 *>(w);
        w += align256((size_t)nqb * c.batch * 8);
        int *order = reinterpret_cast<int *>(w);
        w += list;
        int *order_slot = reinterpret_cast<int *>(w);
        w += list;
        int *count = reinterpret_cast<int *>(w);
        w += align256((size_t)c.batch * 4);
        float *rowrad = reinterpret_cast<float *>(w);
        w += align256((size_t)nqb * c.batch * 8 * 4);
        p.tp_D = reinterpret_cast<double *>(w);
        p.tp_I = reinterpret_cast<int *>(w + (size_t)kTargetSlots * kTargetParts * 32 * 8);
        p.tp_tick = reinterpret_cast<unsigned *>(w + (size_t)kTargetSlots * kTargetParts * 32 * 12);
        w += align256(kTargetPartBytes);
        unsigned *cost = reinterpret_cast<unsigned *>(w);
        w += list;
        unsigned char *grow = reinterpret_cast<unsigned char *>(w);
        p.order_slot = order_slot;
        // measured-cost re-cut of the launch list (block_rebalance_kernel): the fused single-start
        // iteration in a one-wave grid; the launch grid leaves room for every resident slot
        const int rebalance_env = env_int("ISR_ICP_REBALANCE", 1);  // (read per call: the tests switch it)
        const int max_entries = 8 * nqb + kTargetParts * kTargetSlots;
        const bool can_rebalance = V::kFused && c.batch == 1 && rebalance_env != 0 && split_max > 0;
        if (can_rebalance) p.cost = cost;
        const long long warps = (long long)nqb * c.batch;
        if (!c.reuse_order) {
            block_weight_kernel<V::kQueriesPerCta><<<(unsigned)((warps + 3) / 4), 128, 0, c.st>>>(
                p.q, p.q_bstride, p.nq, p.nq_pad, nqb, (int)c.batch, keys, tp_slots > 0 ? rowrad : nullptr);
            ISR_TRY(launched("block_weight_kernel"));
            block_order_kernel<<<(unsigned)c.batch, 1024, 0, c.st>>>(
                keys, nqb, kRows, stride, split_max, parts, split_factor, order, count,
                tp_slots > 0 ? rowrad : nullptr, tp_factor, tp_slots, order_slot, p.tp_tick, mix_slots);
            ISR_TRY(launched("block_order_kernel"));
        }
        p.order = order;
        p.order_count = count;
        grid_x = stride;
        if (can_rebalance) {
            // (entries beyond order_count return at once; a one-wave list grows to the resident slots at
            // most, a deeper one by the CTAs that the cut adds -- a quarter more is never reached)
            const int room = nqb <= slots ? (slots < max_entries ? slots : max_entries)
                                          : (stride + stride / 4 < max_entries ? stride + stride / 4 : max_entries);
            if (grid_x < room) grid_x = room;
            if (c.reuse_order == 2) {
                const float deep_level = (float)env_int("ISR_ICP_RECUT_LEVEL_X100", 100) * 0.01f;
                block_rebalance_kernel<<<1, 1024, 0, c.st>>>(order, order_slot, count, cost, rowrad, grow, nqb, slots,
                                                             room, V::kSplit ? tp_slots : 0, p.tp_tick, deep_level);
                ISR_TRY(launched("block_rebalance_kernel"));
            }
        }
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
    dim3 grid((unsigned)grid_x, (unsigned)splits, (unsigned)c.batch);
    IcpFuse fuse{};
    if (V::kFused) {
        ISR_REQUIRE(c.fuse != nullptr && splits == 1, ISR_E_INVALID_ARG, "nn: fused search without its descriptor");
        fuse = *c.fuse;
        ISR_REQUIRE(fuse.nqb == nqb && fuse.ngroups == (nqb + kFuseGroup - 1) / kFuseGroup, ISR_E_SHAPE,
                    "nn: fused search: reduction layout for %d blocks, launch has %d", fuse.nqb, nqb);
    }
    // programmatic dependent launch of iteration k + 1 behind iteration k (same kernel, nothing else
    // enqueued in between; profiling brackets every launch with events, which would break the pair)
    static const int pdl_env = env_int("ISR_ICP_PDL", 1);
    const bool pdl = V::kFused && c.reuse_order == 1 && pdl_env != 0 && !prof_enabled();
    if (splits == 1) return V::launch(p, grid, c.st, fuse, pdl);

    const long long total = (long long)nq * c.batch;
    const size_t need = align256((size_t)total * splits * 8) + (size_t)total * splits * 4;
    ISR_REQUIRE(c.workspace != nullptr && c.workspace_bytes >= need, ISR_E_WORKSPACE,
                "nn: workspace %zu < %zu bytes", c.workspace_bytes, need);
    p.part_D = reinterpret_cast<double *>(c.workspace);
    p.part_idx = reinterpret_cast<int *>(reinterpret_cast<char *>(c.workspace) +
                                         align256((size_t)total * splits * 8));
    p.part_stride = total;
    ISR_TRY(V::launch(p, grid, c.st, fuse, false));
    nn2_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.st>>>(
        p.part_D, p.part_idx, total, splits, c.out_d2, c.out_idx, c.skip, c.skip_stride,
        (long long)nq);
    return launched("nn2_combine_kernel");
}


}  // namespace isr

extern "C" {

size_t isr_nn2_workspace_bytes(int64_t nq, int64_t nt, int64_t batch) {
    using namespace isr;
    if (nq <= 0 || batch <= 0 || nt <= 0) return 256;
    const int64_t nt_pad = isr_soa_padded_len(nt);
    const int nqb = (int)((nq + 2047) / 2048);  // largest CTA tile of any variant
    const int splits = choose_splits((long long)nqb * batch, (int)(nt_pad / 512), kSlotsUpperBound);
    // launch order of the pruned kernel's 256-query blocks
    const size_t ord = order_workspace_bytes((nq + 255) / 256, batch);
    if (splits <= 1) return ord;
    const size_t total = (size_t)nq * (size_t)batch;
    const size_t part = align256(total * splits * 8) + align256(total * splits * 4);
    return part > ord ? part : ord;
}

int isr_nn2(const IsrCloud *q, const IsrCloud *t, int64_t batch, int use_lo, float *out_d2,
            int32_t *out_idx, const int32_t *skip, int64_t skip_stride, void *workspace,
            size_t workspace_bytes, void *stream) {
    return isr::nn2_search(q, t, batch, use_lo, out_d2, out_idx, skip, skip_stride, workspace, workspace_bytes,
                           stream, 0, nullptr);
}

}  // extern "C"

namespace isr {

bool nn2_fusable(const IsrCloud *t) {
    return t != nullptr && t->sub_c != nullptr && t->stage_c != nullptr && pruning_on();
}

int nn2_query_blocks(int64_t nq) { return (int)((nq + NN2Pruned::kQueriesPerCta - 1) / NN2Pruned::kQueriesPerCta); }

int nn2_search(const IsrCloud *q, const IsrCloud *t, int64_t batch, int use_lo, float *out_d2,
               int32_t *out_idx, const int32_t *skip, int64_t skip_stride, void *workspace,
               size_t workspace_bytes, void *stream, int reuse_order, const IcpFuse *fuse) {
    ISR_REQUIRE(q != nullptr && t != nullptr, ISR_E_INVALID_ARG, "nn: null cloud descriptor");
    ISR_REQUIRE(q->n >= 0 && t->n >= 1 && batch >= 0, ISR_E_SHAPE,
                "nn: need nq >= 0, nt >= 1, batch >= 0 (nq=%lld nt=%lld batch=%lld)",
                (long long)q->n, (long long)t->n, (long long)batch);
    if (q->n == 0 || batch == 0) return ISR_OK;
    ISR_REQUIRE(q->soa7 && t->soa7 && (out_d2 || fuse), ISR_E_INVALID_ARG, "nn: null pointer");
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
                    workspace_bytes, (cudaStream_t)stream, reuse_order, fuse};
    if (fuse != nullptr) {
        ISR_REQUIRE(nn2_fusable(t) && q->hint != nullptr && use_lo, ISR_E_INVALID_ARG,
                    "nn: the fused ICP iteration needs the pruned search, hints and the lo planes");
        ISR_REQUIRE(aligned16(t->sub_c), ISR_E_ALIGN, "nn: sub-tile spheres must be 16-byte aligned");
        // (a batch of starts = multi-start ICP: scan-heavy, see NN2PrunedHalves)
        if (batch > 1) return nn2_dispatch<NN2PrunedHalvesFused>(c);
        // a single start whose whole blocks fit one wave (a source shard of a multi-GPU run, a
        // small cloud): the iteration lasts as long as its slowest warp -- cut rows at curve jumps
        static const int split_env = env_int("ISR_NN_SPLIT_ROWS", -1);
        const bool shallow = nn2_query_blocks(q->n) <= sm_count() * NN2PrunedFused::ctas_per_sm(true);
        if (split_env >= 0 ? split_env != 0 : shallow) return nn2_dispatch<NN2PrunedFusedSplit>(c);
        return nn2_dispatch<NN2PrunedFused>(c);
    }
    if (t->sub_c != nullptr && t->stage_c != nullptr && pruning_on()) {
        ISR_REQUIRE(aligned16(t->sub_c), ISR_E_ALIGN, "nn: sub-tile spheres must be 16-byte aligned");
#ifdef ISR_NN_TUNING
        static int parts = -1;
        if (parts < 0) { const char *e = getenv("ISR_NN_PARTS"); parts = e ? atoi(e) : 1; }
        if (parts == 2) return nn2_dispatch<NN2PrunedP2>(c);
        if (parts == 4) return nn2_dispatch<NN2PrunedP4>(c);
        if (parts == 24) return nn2_dispatch<NN2PrunedHalves>(c);
        if (parts == 12) return nn2_dispatch<NN2PrunedU2>(c);
        if (parts == 14) return nn2_dispatch<NN2PrunedU4>(c);
#endif
        // a batch of hinted searches = multi-start ICP: scan-heavy, see NN2PrunedHalves
        if (q->hint != nullptr && batch > 1) return nn2_dispatch<NN2PrunedHalves>(c);
        return nn2_dispatch<NN2Pruned>(c);
    }
    return nn2_dispatch<NN2Main>(c);
}

}  // namespace isr

extern "C" {

int isr_profile_nn_counters(uint64_t *out8_host) {
    using namespace isr;
    ISR_REQUIRE(out8_host != nullptr, ISR_E_INVALID_ARG, "profile_nn_counters: null pointer");
    unsigned long long *ctr = evaluated_counter();
    ISR_REQUIRE(ctr != nullptr, ISR_E_CUDA, "profile_nn_counters: no counter buffer");
    ISR_TRY(check_cuda(cudaDeviceSynchronize(), "profile_nn_counters sync"));
    return check_cuda(cudaMemcpy(out8_host, ctr, 64, cudaMemcpyDeviceToHost), "profile_nn_counters read");
}

int isr_debug_cta_log(uint64_t *dev_buf, int64_t capacity_records) {
    isr::g_cta_log = reinterpret_cast<unsigned long long *>(dev_buf);
    isr::g_cta_log_cap = dev_buf != nullptr ? capacity_records : 0;
    return ISR_OK;
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
        ISR_TRY(check_cuda(cudaMemset(ctr, 0, 64), "profile_nn_pairs clear"));
    }
    // counted in quarter units: 32 x 2 queries against one 64-target sub-tile
    if (evaluated_host) *evaluated_host = (uint64_t)units * 64ull * (uint64_t)ISR_SUB_TILE;
    if (answered_host) *answered_host = (uint64_t)g_answered.exchange(0);
    return ISR_OK;
}

}  // extern "C"
