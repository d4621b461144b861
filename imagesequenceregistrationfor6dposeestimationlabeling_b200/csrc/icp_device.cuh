// Device-side pieces of the ICP iteration shared by icp.cu (stepwise kernels) and nn2.cu (the
// fused iteration: search + correspondence sums + exchange + Kabsch in ONE kernel).
//
//   PeerView         the kernel's view of the CUDA-IPC exchange buffers of one box (isr_peer_*)
//   icp_solve_state  Open3D's break test + Eigen::umeyama (no scaling) + T <- U T on one state
//   IcpFuse          what the fused search kernel needs beyond the search itself
//   icp_fused_tail   reduction of the per-row sums in a launch-independent fixed order, the
//                    peer exchange and the solve, run by the last warp to arrive
#pragma once

#include <math_constants.h>

#include "isr_common.cuh"

namespace isr {

constexpr int kNS = ISR_ICP_NSUMS;

// ---- exchange of the 17 sums between the GPUs of one box ---------------------------------------
// Every rank owns one small buffer that all peers map through CUDA IPC (NVLink / NVSwitch
// peer memory):  data[2][world][kPeerStarts][17] doubles and flag[2][world][kPeerStarts].
// The warp that finishes a rank's reduction stores the rank's sums for start s into slot
// [seq & 1][rank][s] of EVERY rank's buffer, fences, and then stores the message number seq into
// the matching flags; it then waits until its own buffer holds seq from all ranks and adds the
// `world` vectors in rank order -- the same order everywhere, so all ranks solve bit-identical
// problems.  No collective library call, no extra launch, nothing read by the host.  Two slots
// suffice: a rank can only start message seq + 1 after it has received every peer's seq, i.e.
// after every peer has finished reading message seq - 1 from the slot that seq + 1 overwrites.
constexpr int kPeerRanks = ISR_PEER_MAX_RANKS;
constexpr int kPeerStarts = ISR_PEER_MAX_STARTS;
struct PeerView {
    double *data[kPeerRanks];
    unsigned long long *flag[kPeerRanks];
    int rank, world;  // world == 0: no exchange
    unsigned long long seq;
    long long timeout_cycles;  // give up waiting for a peer after this many SM cycles
};
__host__ __device__ inline size_t peer_data_index(const PeerView &v, int from_rank, int start) {
    return (((size_t)(v.seq & 1) * kPeerRanks + from_rank) * kPeerStarts + start) * kNS;
}
__host__ __device__ inline size_t peer_flag_index(const PeerView &v, int from_rank, int start) {
    return ((size_t)(v.seq & 1) * kPeerRanks + from_rank) * kPeerStarts + start;
}
constexpr size_t kPeerDataBytes = (size_t)2 * kPeerRanks * kPeerStarts * kNS * sizeof(double);
constexpr size_t kPeerFlagBytes = (size_t)2 * kPeerRanks * kPeerStarts * sizeof(unsigned long long);

#ifdef __CUDACC__
// ---- 3x3 SVD (one-sided Jacobi, FP64) and Kabsch ---------------------------------------
// (Every index below is a compile-time constant once the loops are unrolled: the matrices live in
// registers.  With the column pair (p, q) as a run-time index they sat in local memory, and this
// function -- one thread, at the very end of every ICP iteration, the whole machine waiting -- spent
// most of its ~15 k cycles on the round trips.)
template <int P, int Q>
__device__ __forceinline__ bool svd3_rotate(double (&A)[3][3], double (&V)[3][3]) {
    double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        alpha += A[k][P] * A[k][P];
        beta += A[k][Q] * A[k][Q];
        gamma += A[k][P] * A[k][Q];
    }
    // one square root, one division and one reciprocal square root per rotation: the convergence
    // test on squares, and
    //   t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)),  zeta = (beta - alpha) / (2 gamma)
    //     = sign(beta - alpha) 2 gamma / (|beta - alpha| + sqrt((beta - alpha)^2 + 4 gamma^2))
    if (gamma == 0.0 || gamma * gamma <= (2.3e-16 * 2.3e-16) * (alpha * beta)) return false;
    const double delta = beta - alpha, g2 = 2.0 * gamma;
    double t = g2 / (fabs(delta) + sqrt(fma(delta, delta, g2 * g2)));
    t = delta < 0.0 ? -t : t;
    const double c = rsqrt(fma(t, t, 1.0)), sn = c * t;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double ap = A[k][P], aq = A[k][Q];
        A[k][P] = c * ap - sn * aq;
        A[k][Q] = sn * ap + c * aq;
        const double vp = V[k][P], vq = V[k][Q];
        V[k][P] = c * vp - sn * vq;
        V[k][Q] = sn * vp + c * vq;
    }
    return true;
}
template <int A_, int B_>
__device__ __forceinline__ void svd3_order(double (&D)[3], double (&A)[3][3], double (&V)[3][3]) {
    if (D[B_] > D[A_]) {
        const double td = D[A_]; D[A_] = D[B_]; D[B_] = td;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double ta = A[k][A_]; A[k][A_] = A[k][B_]; A[k][B_] = ta;
            const double tv = V[k][A_]; V[k][A_] = V[k][B_]; V[k][B_] = tv;
        }
    }
}
__device__ inline void svd3(const double (&M)[3][3], double (&U)[3][3], double (&D)[3], double (&V)[3][3]) {
    double A[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            A[i][j] = M[i][j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
#pragma unroll 1
    for (int sweep = 0; sweep < 64; ++sweep) {
        bool rotated = svd3_rotate<0, 1>(A, V);
        rotated = svd3_rotate<0, 2>(A, V) || rotated;
        rotated = svd3_rotate<1, 2>(A, V) || rotated;
        if (!rotated) break;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
        D[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    // sort singular values descending (column permutation of A and V)
    svd3_order<0, 1>(D, A, V);
    svd3_order<0, 2>(D, A, V);
    svd3_order<1, 2>(D, A, V);
    const double tiny = D[0] * 1e-14;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int k = 0; k < 3; ++k) U[k][j] = 0.0;
        if (D[j] > tiny && D[j] > 0.0) {
            const double inv = 1.0 / D[j];
#pragma unroll
            for (int k = 0; k < 3; ++k) U[k][j] = A[k][j] * inv;
            ++rank;
        }
    }
    if (rank == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) U[i][j] = (i == j) ? 1.0 : 0.0;
    } else {
        if (rank == 1) {
            // any unit vector orthogonal to U[:,0]: the axis of its smallest component, projected
            int m = 0;
            double um = U[0][0];
            if (fabs(U[1][0]) < fabs(um)) { m = 1; um = U[1][0]; }
            if (fabs(U[2][0]) < fabs(um)) { m = 2; um = U[2][0]; }
            double w[3], nn = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) { w[k] = (k == m ? 1.0 : 0.0) - um * U[k][0]; nn += w[k] * w[k]; }
            nn = sqrt(nn);
#pragma unroll
            for (int k = 0; k < 3; ++k) U[k][1] = w[k] / nn;
        }
        if (rank <= 2) {
            U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
            U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
            U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
        }
    }
}

__device__ __forceinline__ double det3(const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) -
           M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// One state, one thread: consume the 17 sums S of an evaluation (sum s[3], sum t[3], sum t s^T
// [9], sum d2, count; already reduced over every source shard).  Open3D's RegistrationICP:
// fitness / rmse of this evaluation, the break test against the previous one, else
// Eigen::umeyama without scaling (Sigma = (1/n) sum (t - mu_t)(s - mu_s)^T, SVD, det-sign fix)
// and T <- U T.  `final_eval` marks the evaluation after the last allowed update.
__device__ inline void icp_solve_state(IsrIcpState &st, const double *S, int64_t ns_total,
                                       double rel_fitness, double rel_rmse, int final_eval) {
    const double cnt = S[16];
    const double fitness = ns_total > 0 ? cnt / (double)ns_total : 0.0;
    const double rmse = cnt > 0.0 ? sqrt(S[15] / cnt) : 0.0;
    const bool had_prev = st.evals > 0;
    const double pf = st.fitness, pr = st.inlier_rmse;
    st.prev_fitness = pf;
    st.prev_rmse = pr;
    st.fitness = fitness;
    st.inlier_rmse = rmse;
    st.n_corr = (int64_t)cnt;
    st.evals += 1;
    if (had_prev && fabs(pf - fitness) < rel_fitness && fabs(pr - rmse) < rel_rmse) {
        st.done = 1;
        return;
    }
    if (final_eval) {
        st.done = 1;
        return;
    }
    st.iters += 1;
    if (!(cnt > 0.0)) return;  // empty correspondence set: U = I

    const double inv = 1.0 / cnt;
    const double ms[3] = {S[0] * inv, S[1] * inv, S[2] * inv};
    const double mt[3] = {S[3] * inv, S[4] * inv, S[5] * inv};
    double Sig[3][3], U[3][3], V[3][3], D[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Sig[i][j] = S[6 + 3 * i + j] * inv - mt[i] * ms[j];
    svd3(Sig, U, D, V);
    const double sgn = (det3(U) * det3(V) < 0.0) ? -1.0 : 1.0;
    double R[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + sgn * U[i][2] * V[j][2];
    double tr[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        tr[i] = mt[i] - (R[i][0] * ms[0] + R[i][1] * ms[1] + R[i][2] * ms[2]);
    // T <- [R tr; 0 1] . T   (the old pose is fetched in one go)
    double To[12], Tn[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) To[k] = st.T[k];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double v = R[i][0] * To[0 + j] + R[i][1] * To[4 + j] + R[i][2] * To[8 + j];
            if (j == 3) v += tr[i];
            Tn[4 * i + j] = v;
        }
#pragma unroll
    for (int k = 0; k < 12; ++k) st.T[k] = Tn[k];
    st.T[12] = 0.0; st.T[13] = 0.0; st.T[14] = 0.0; st.T[15] = 1.0;
}
// The FP64 arithmetic of one correspondence, spelled out operation by operation so that every
// kernel that evaluates it (fused search epilogue, stepwise accumulate, exact-distance kernel)
// produces the same bits: s = T p (row-major 3x4, fma chain from the left), d = s - t,
// d2 = fma(dz, dz, fma(dy, dy, dx dx)).
__device__ __forceinline__ void icp_apply_pose(const double (&T)[12], double px, double py, double pz,
                                               double &sx, double &sy, double &sz) {
    sx = __dadd_rn(__fma_rn(T[2], pz, __fma_rn(T[1], py, __dmul_rn(T[0], px))), T[3]);
    sy = __dadd_rn(__fma_rn(T[6], pz, __fma_rn(T[5], py, __dmul_rn(T[4], px))), T[7]);
    sz = __dadd_rn(__fma_rn(T[10], pz, __fma_rn(T[9], py, __dmul_rn(T[8], px))), T[11]);
}
__device__ __forceinline__ double icp_dist2(double sx, double sy, double sz, double tx, double ty, double tz) {
    const double dx = __dsub_rn(sx, tx), dy = __dsub_rn(sy, ty), dz = __dsub_rn(sz, tz);
    return __fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx)));
}
// the 17 contributions of an inlier correspondence (s transformed source, t target, d2)
__device__ __forceinline__ void icp_contrib(double sx, double sy, double sz, double tx, double ty, double tz,
                                            double d2, double (&c)[kNS]) {
    c[0] = sx; c[1] = sy; c[2] = sz;
    c[3] = tx; c[4] = ty; c[5] = tz;
    c[6] = __dmul_rn(tx, sx); c[7] = __dmul_rn(tx, sy); c[8] = __dmul_rn(tx, sz);
    c[9] = __dmul_rn(ty, sx); c[10] = __dmul_rn(ty, sy); c[11] = __dmul_rn(ty, sz);
    c[12] = __dmul_rn(tz, sx); c[13] = __dmul_rn(tz, sy); c[14] = __dmul_rn(tz, sz);
    c[15] = d2;
    c[16] = 1.0;
}
// Sum of one query row: every lane has stored its 17 contributions at scratch[lane * 17 + k]
// (shared memory, 32 x 17 doubles); lane k < 17 returns component k summed in lane order.
__device__ __forceinline__ double icp_row_sum(const double *scratch, int lane) {
    double sum = 0.0;
    if (lane < kNS) {
#pragma unroll 8
        for (int j = 0; j < 32; ++j) sum = __dadd_rn(sum, scratch[j * kNS + lane]);
    }
    return sum;
}
#endif  // __CUDACC__

// ---- the fused iteration ------------------------------------------------------------------------
// nn2_pruned_kernel<FUSED> is one whole ICP evaluation + update for `batch` starts:
//   prologue  every warp transforms its 256 stored source points by the start's FP64 pose,
//             centres them on the target's centroid and splits them into the hi/lo pair the
//             search runs on (what isr_prepare_cloud writes to HBM for the stepwise path);
//   search    unchanged (hints in, exact neighbours out);
//   epilogue  per source point: its neighbour's ORIGINAL coordinates (one 16-byte gather from
//             tgt4), the FP64 distance from the FP64-transformed source point, the strict
//             d2 < max_d2 test, and its 17 contributions; per query ROW a fixed-order sum;
//   tail      rows -> block -> group of 64 blocks -> start, each level summed in index order
//             by the last warp to arrive (tickets), so the result does not depend on how the
//             launch was decomposed into CTAs or on what else is in the batch; then the peer
//             exchange and the solve, by the very last warp.
constexpr int kFuseGroup = 64;  // query blocks per reduction group
struct IcpFuse {
    IsrIcpState *states;     // [batch]; NULL: not fused
    const float *src7;       // [7][nq_pad] the ORIGINAL source in stored order: hi planes 0-2, lo planes 4-6
    const float4 *tgt4;      // [nt_pad] the ORIGINAL target coordinates in stored order
    const double *centroid;  // [3] the centre the target was prepared with
    double max_d2;
    double *rowsum;          // [batch][nqb][8][17]   (only written by CTAs that own part of a block)
    double *blocksum;        // [batch][nqb][17]
    double *groupsum;        // [batch][ngroups][17]
    unsigned *tickets;       // [batch][nqb + ngroups + 1], zero between launches
    double *sums;            // [batch][17] out: this rank's sums of the evaluation
    uint8_t *inlier;         // [batch][nq_pad] out, stored order
    int nqb, ngroups;
    long long ns_total;
    double rel_fitness, rel_rmse;
    int final_eval;
    int do_solve;            // 0: stop after `sums` (the stepwise accumulate kernel uses the same tail)
    PeerView px;
};

#if defined(__CUDACC__)
__device__ __forceinline__ double ldcg_f64(const double *p) { return __ldcg(p); }

// rs[r] (r = 0..7): component `lane` (< 17) of the sum of query row r of block `blk`, for the rows
// in `own` (the rows this CTA is responsible for, live or not).  One warp.
static __device__ __noinline__ void icp_fused_tail(const IcpFuse &f, int b, int blk, unsigned own, int lane,
                                            const double *rs) {
    const unsigned full = 0xffffffffu;
    const int k = lane < kNS ? lane : 0;
    unsigned *tick = f.tickets + (size_t)b * (f.nqb + f.ngroups + 1);
    double bs = 0.0;
    // ---- rows -> block --------------------------------------------------------------------
    if (own == 0xFFu) {
        for (int r = 0; r < 8; ++r) bs = __dadd_rn(bs, rs[r]);
    } else {
        double *row = f.rowsum + (((size_t)b * f.nqb + blk) * 8) * kNS;
        if (lane < kNS)
            for (int r = 0; r < 8; ++r)
                if ((own >> r) & 1u) row[r * kNS + k] = rs[r];
        __threadfence();
        __syncwarp();
        unsigned last = 0;
        if (lane == 0) {
            const unsigned mine = (unsigned)__popc(own);
            last = atomicAdd(&tick[blk], mine) + mine == 8u ? 1u : 0u;
        }
        if (__shfl_sync(full, last, 0) == 0) return;
        __threadfence();
        for (int r = 0; r < 8; ++r) bs = __dadd_rn(bs, ldcg_f64(row + r * kNS + k));
        if (lane == 0) tick[blk] = 0;
    }
    // ---- blocks -> group ------------------------------------------------------------------
    const int g = blk / kFuseGroup;
    const int g_first = g * kFuseGroup;
    const int g_count = min(kFuseGroup, f.nqb - g_first);
    double gs = bs;
    if (g_count > 1) {
        double *bsum = f.blocksum + ((size_t)b * f.nqb) * kNS;
        if (lane < kNS) bsum[(size_t)blk * kNS + k] = bs;
        __threadfence();
        __syncwarp();
        unsigned last = 0;
        if (lane == 0) last = atomicAdd(&tick[f.nqb + g], 1u) + 1u == (unsigned)g_count ? 1u : 0u;
        if (__shfl_sync(full, last, 0) == 0) return;
        __threadfence();
        gs = 0.0;
#pragma unroll 16
        for (int x = 0; x < g_count; ++x) gs = __dadd_rn(gs, ldcg_f64(bsum + (size_t)(g_first + x) * kNS + k));
        if (lane == 0) tick[f.nqb + g] = 0;
    }
    // ---- groups -> start ------------------------------------------------------------------
    double S = gs;
    if (f.ngroups > 1) {
        double *gsum = f.groupsum + ((size_t)b * f.ngroups) * kNS;
        if (lane < kNS) gsum[(size_t)g * kNS + k] = gs;
        __threadfence();
        __syncwarp();
        unsigned last = 0;
        if (lane == 0) last = atomicAdd(&tick[f.nqb + f.ngroups], 1u) + 1u == (unsigned)f.ngroups ? 1u : 0u;
        if (__shfl_sync(full, last, 0) == 0) return;
        __threadfence();
        S = 0.0;
#pragma unroll 8
        for (int x = 0; x < f.ngroups; ++x) S = __dadd_rn(S, ldcg_f64(gsum + (size_t)x * kNS + k));
        if (lane == 0) tick[f.nqb + f.ngroups] = 0;
    }
    // ---- this warp holds the start's sums: exchange with the peers, then solve -------------------
    if (lane < kNS) f.sums[(size_t)b * kNS + k] = S;
    if (!f.do_solve) return;
    IsrIcpState &st = f.states[b];
    const PeerView &px = f.px;
    if (px.world > 0) {
        if (lane < kNS)
            for (int r = 0; r < px.world; ++r) px.data[r][peer_data_index(px, px.rank, b) + k] = S;
        __threadfence_system();
        __syncwarp();
        if (lane < px.world) {
            volatile unsigned long long *fl = px.flag[lane] + peer_flag_index(px, px.rank, b);
            *fl = px.seq;
        }
        bool ok = true;
        if (lane < px.world) {
            volatile unsigned long long *fl = px.flag[px.rank] + peer_flag_index(px, lane, b);
            const long long t0 = clock64();
            while (*fl != px.seq) {
                if (clock64() - t0 > px.timeout_cycles) { ok = false; break; }  // a peer died
                __nanosleep(64);
            }
        }
        ok = __all_sync(full, ok);
        __threadfence_system();
        if (!ok) {
            if (lane == 0) {
                st.done = 1;
                st.reserved = 1;  // exchange timed out
                st.fitness = CUDART_NAN;
                st.inlier_rmse = CUDART_NAN;
            }
            return;
        }
        double t = 0.0;
        for (int r = 0; r < px.world; ++r) {
            const volatile double *d = px.data[px.rank] + peer_data_index(px, r, b);
            t += d[k];
        }
        S = t;
    }
    double Sv[kNS];
#pragma unroll
    for (int i = 0; i < kNS; ++i) Sv[i] = __shfl_sync(full, S, i);
    if (lane == 0) icp_solve_state(st, Sv, f.ns_total, f.rel_fitness, f.rel_rmse, f.final_eval);
}
#endif  // __CUDACC__

}  // namespace isr
