// K3 -- point-to-point ICP with the iteration loop resident on the device.
//
// Replaces o3d.pipelines.registration.evaluate_registration / registration_icp with
// TransformationEstimationPointToPoint (icp.py:96-103).  Upstream semantics restated:
//   evaluation   : transform source by T, 1-NN per point, keep d2 < max_dist^2 (strict),
//                  fitness = #corr / n_source, inlier_rmse = sqrt(sum d2 / #corr);
//   update       : Eigen::umeyama(src[corr], tgt[corr], no scaling) -> U;  T <- U T;
//   loop         : evaluate; repeat { update; evaluate; break if |dfitness| < rf and
//                  |drmse| < rr } at most max_iteration times; the result is the last
//                  evaluation; an empty correspondence set gives U = I.
//
// One evaluation = K1 (FP64 transform -> FP32 SoA) + K2 (FP32 brute-force 1-NN, indices)
// + icp_accumulate_kernel (HBM-bound gather: 12 B source + 4 B index + 12 B gathered
// target (+ 1 B flag written) per source point).  The gather re-derives every
// correspondence distance in FP64 from the original source point and the FP64 pose, so
// the inlier test, fitness, rmse and the 17 Kabsch sums carry no FP32 error; only the
// choice of neighbour is made in FP32.  Sums are reduced warp -> CTA -> last-arriving CTA
// in a fixed order (no floating-point atomics).  icp_solve_kernel (one warp per start)
// applies the break test, solves the 3x3 SVD by one-sided Jacobi in FP64 and composes
// T <- U T.  A finished start sets `done`; all later kernels for it exit at once, so the
// host enqueues max_iteration + 1 passes without ever reading the device.
#include <math_constants.h>
#include <string.h>

#include "isr_common.cuh"

namespace isr {

constexpr int kAccThreads = 256;
constexpr int kNS = ISR_ICP_NSUMS;
constexpr int kStateInts = (int)(sizeof(IsrIcpState) / sizeof(int32_t));
constexpr int kStateDoubles = (int)(sizeof(IsrIcpState) / sizeof(double));
static_assert(sizeof(IsrIcpState) % 8 == 0, "IsrIcpState must be a whole number of doubles");

// ---- exchange of the 17 sums between the GPUs of one box, fused into the two ICP kernels ----
// Every rank owns one small buffer that all peers map through CUDA IPC (NVLink / NVSwitch
// peer memory):  data[2][world][kPeerStarts][17] doubles and flag[2][world][kPeerStarts].
// The last-arriving CTA of the accumulate kernel stores this rank's sums for start s into
// slot [seq & 1][rank][s] of EVERY rank's buffer, fences, and then stores the message number
// seq into the matching flags; the solve kernel of every rank waits until its own buffer
// holds seq from all ranks and adds the `world` vectors in rank order -- the same order
// everywhere, so all ranks solve bit-identical problems.  No collective library call, no
// extra launch, nothing read by the host.  Two slots suffice: a rank can only start message
// seq + 1 after it has received every peer's seq, i.e. after every peer has finished reading
// message seq - 1 from the slot that seq + 1 overwrites.
constexpr int kPeerRanks = ISR_PEER_MAX_RANKS;
constexpr int kPeerStarts = ISR_PEER_MAX_STARTS;
struct PeerView {
    double *data[kPeerRanks];
    unsigned long long *flag[kPeerRanks];
    int rank, world;  // world == 0: no exchange
    unsigned long long seq;
};
__host__ __device__ inline size_t peer_data_index(const PeerView &v, int from_rank, int start) {
    return (((size_t)(v.seq & 1) * kPeerRanks + from_rank) * kPeerStarts + start) * kNS;
}
__host__ __device__ inline size_t peer_flag_index(const PeerView &v, int from_rank, int start) {
    return ((size_t)(v.seq & 1) * kPeerRanks + from_rank) * kPeerStarts + start;
}
constexpr size_t kPeerDataBytes = (size_t)2 * kPeerRanks * kPeerStarts * kNS * sizeof(double);
constexpr size_t kPeerFlagBytes = (size_t)2 * kPeerRanks * kPeerStarts * sizeof(unsigned long long);

// grid: (nblk, starts)
__global__ void __launch_bounds__(kAccThreads)
icp_accumulate_kernel(const IsrIcpState *__restrict__ states, const float *__restrict__ src,
                      const float *__restrict__ src_lo, int64_t ns, const float *__restrict__ tgt, const int32_t *__restrict__ idx,
                      double max_d2, double *__restrict__ partials, unsigned *__restrict__ tickets,
                      double *__restrict__ sums, uint8_t *__restrict__ inlier, const PeerView px) {
    const int s = blockIdx.y;
    const IsrIcpState &stt = states[s];
    if (stt.done != 0) return;
    __shared__ double red[kAccThreads / 32][kNS];
    __shared__ bool is_last;

    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = stt.T[k];
    double acc[kNS];
#pragma unroll
    for (int k = 0; k < kNS; ++k) acc[k] = 0.0;

    const int32_t *ids = idx + (int64_t)s * ns;
    uint8_t *inl = inlier != nullptr ? inlier + (int64_t)s * ns : nullptr;
    // 4 points per thread and trip, all loads of a trip issued before the first use: the
    // dependent gather (index -> target row) is what bounds this kernel, so keep 4 in flight.
    constexpr int U = 4;
    const int64_t stride = (int64_t)gridDim.x * kAccThreads;
    for (int64_t i0 = (int64_t)blockIdx.x * kAccThreads + threadIdx.x; i0 < ns; i0 += U * stride) {
        float fx[U], fy[U], fz[U], lx[U], ly[U], lz[U];
        int j[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            ok[u] = i < ns;
            const int64_t ii = ok[u] ? i : i0;
            fx[u] = src[3 * ii]; fy[u] = src[3 * ii + 1]; fz[u] = src[3 * ii + 2];
            lx[u] = ly[u] = lz[u] = 0.f;
            if (src_lo != nullptr) {
                lx[u] = src_lo[3 * ii]; ly[u] = src_lo[3 * ii + 1]; lz[u] = src_lo[3 * ii + 2];
            }
            j[u] = ids[ii];
            // a negative index: this source point has no correspondence here (target-sharded
            // ICP: its neighbour lives on another rank)
            ok[u] = ok[u] && j[u] >= 0;
        }
        float gx[U], gy[U], gz[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t jj = j[u] >= 0 ? j[u] : 0;
            gx[u] = tgt[3 * jj]; gy[u] = tgt[3 * jj + 1]; gz[u] = tgt[3 * jj + 2];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double px = (double)fx[u] + (double)lx[u], py = (double)fy[u] + (double)ly[u],
                         pz = (double)fz[u] + (double)lz[u];
            const double sx = ((T[0] * px + T[1] * py) + T[2] * pz) + T[3];
            const double sy = ((T[4] * px + T[5] * py) + T[6] * pz) + T[7];
            const double sz = ((T[8] * px + T[9] * py) + T[10] * pz) + T[11];
            const double tx = gx[u], ty = gy[u], tz = gz[u];
            const double dx = sx - tx, dy = sy - ty, dz = sz - tz;
            const double d2 = dx * dx + dy * dy + dz * dz;
            const bool in = ok[u] && d2 < max_d2;
            if (inl != nullptr && i0 + u * stride < ns) inl[i0 + u * stride] = in ? 1 : 0;
            if (in) {
                acc[0] += sx; acc[1] += sy; acc[2] += sz;
                acc[3] += tx; acc[4] += ty; acc[5] += tz;
                acc[6] += tx * sx; acc[7] += tx * sy; acc[8] += tx * sz;
                acc[9] += ty * sx; acc[10] += ty * sy; acc[11] += ty * sz;
                acc[12] += tz * sx; acc[13] += tz * sy; acc[14] += tz * sz;
                acc[15] += d2;
                acc[16] += 1.0;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kNS; ++k) acc[k] = warp_sum(acc[k]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kNS; ++k) red[warp][k] = acc[k];
    }
    __syncthreads();
    double *my = partials + ((int64_t)s * gridDim.x + blockIdx.x) * kNS;
    if (threadIdx.x < kNS) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kAccThreads / 32; ++w) v += red[w][threadIdx.x];
        my[threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&tickets[s], 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        // final reduction over the CTAs' partials by the whole block: thread t takes blocks
        // t, t+256, ... (fixed mapping), then the same warp -> CTA tree as above
        __threadfence();
        const double *base = partials + (int64_t)s * gridDim.x * kNS;
        double v[kNS];
#pragma unroll
        for (int k = 0; k < kNS; ++k) v[k] = 0.0;
        for (unsigned bk = threadIdx.x; bk < gridDim.x; bk += kAccThreads) {
#pragma unroll
            for (int k = 0; k < kNS; ++k) v[k] += base[(int64_t)bk * kNS + k];
        }
#pragma unroll
        for (int k = 0; k < kNS; ++k) v[k] = warp_sum(v[k]);
        __syncthreads();  // `red` is reused
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kNS; ++k) red[warp][k] = v[k];
        }
        __syncthreads();
        if (threadIdx.x < kNS) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kAccThreads / 32; ++w) t += red[w][threadIdx.x];
            sums[(int64_t)s * kNS + threadIdx.x] = t;
            // push this rank's sums into every rank's exchange buffer (peer stores over NVLink)
            for (int r = 0; r < px.world; ++r) px.data[r][peer_data_index(px, px.rank, s) + threadIdx.x] = t;
        }
        if (px.world > 0) {
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x < px.world) {
                volatile unsigned long long *f = px.flag[threadIdx.x] + peer_flag_index(px, px.rank, s);
                *f = px.seq;
            }
        }
        if (threadIdx.x == 0) tickets[s] = 0;
    }
}

// out_D[start][i] = FP64 squared distance between T.src[i] and tgt[idx[start][i]] (the same
// arithmetic as the accumulate kernel); +inf where idx < 0.  grid (blocks, starts).
__global__ void __launch_bounds__(256)
icp_corr_dist_kernel(const IsrIcpState *__restrict__ states, const float *__restrict__ src,
                     const float *__restrict__ src_lo, int64_t ns, const float *__restrict__ tgt,
                     const int32_t *__restrict__ idx, double *__restrict__ out_D) {
    const int s = blockIdx.y;
    const IsrIcpState &stt = states[s];
    if (stt.done != 0) return;
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = stt.T[k];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < ns; i += (int64_t)gridDim.x * 256) {
        const int j = idx[(int64_t)s * ns + i];
        double D = CUDART_INF;
        if (j >= 0) {
            double px = src[3 * i], py = src[3 * i + 1], pz = src[3 * i + 2];
            if (src_lo != nullptr) { px += (double)src_lo[3 * i]; py += (double)src_lo[3 * i + 1]; pz += (double)src_lo[3 * i + 2]; }
            const double sx = ((T[0] * px + T[1] * py) + T[2] * pz) + T[3];
            const double sy = ((T[4] * px + T[5] * py) + T[6] * pz) + T[7];
            const double sz = ((T[8] * px + T[9] * py) + T[10] * pz) + T[11];
            const double dx = sx - (double)tgt[3ll * j], dy = sy - (double)tgt[3ll * j + 1],
                         dz = sz - (double)tgt[3ll * j + 2];
            D = dx * dx + dy * dy + dz * dz;
        }
        out_D[(int64_t)s * ns + i] = D;
    }
}

// hint[start][*] = -1 for the starts that have not been evaluated yet (their workspace may
// hold anything); a no-op for every later iteration
__global__ void __launch_bounds__(256)
icp_hint_reset_kernel(const IsrIcpState *__restrict__ states, int64_t nsp, long long total,
                      int32_t *__restrict__ hint) {
    const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= total) return;
    const IsrIcpState &st = states[i0 / nsp];  // nsp is a multiple of 1024: 4 slots, one start
    if (st.evals != 0 || st.done != 0) return;
    *reinterpret_cast<int4 *>(hint + i0) = make_int4(-1, -1, -1, -1);
}

// ---- 3x3 SVD (one-sided Jacobi, FP64) and Kabsch ---------------------------------------
__device__ void svd3(const double M[3][3], double U[3][3], double D[3], double V[3][3]) {
    double A[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[i][j] = M[i][j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    const int P[3] = {0, 0, 1}, Qc[3] = {1, 2, 2};
    for (int sweep = 0; sweep < 64; ++sweep) {
        bool rotated = false;
        for (int pr = 0; pr < 3; ++pr) {
            const int p = P[pr], q = Qc[pr];
            double alpha = 0, beta = 0, gamma = 0;
            for (int k = 0; k < 3; ++k) {
                alpha += A[k][p] * A[k][p];
                beta += A[k][q] * A[k][q];
                gamma += A[k][p] * A[k][q];
            }
            if (gamma == 0.0 || fabs(gamma) <= 2.3e-16 * sqrt(alpha * beta)) continue;
            rotated = true;
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
            for (int k = 0; k < 3; ++k) {
                const double ap = A[k][p], aq = A[k][q];
                A[k][p] = c * ap - sn * aq;
                A[k][q] = sn * ap + c * aq;
                const double vp = V[k][p], vq = V[k][q];
                V[k][p] = c * vp - sn * vq;
                V[k][q] = sn * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    for (int j = 0; j < 3; ++j)
        D[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    // sort singular values descending (column permutation of A and V)
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (D[b] > D[a]) {
                const double td = D[a]; D[a] = D[b]; D[b] = td;
                for (int k = 0; k < 3; ++k) {
                    const double ta = A[k][a]; A[k][a] = A[k][b]; A[k][b] = ta;
                    const double tv = V[k][a]; V[k][a] = V[k][b]; V[k][b] = tv;
                }
            }
    const double tiny = D[0] * 1e-14;
    int rank = 0;
    for (int j = 0; j < 3; ++j) {
        if (D[j] > tiny && D[j] > 0.0) {
            for (int k = 0; k < 3; ++k) U[k][j] = A[k][j] / D[j];
            ++rank;
        }
    }
    if (rank == 0) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) U[i][j] = (i == j) ? 1.0 : 0.0;
    } else {
        if (rank == 1) {
            // any unit vector orthogonal to U[:,0]
            int m = 0;
            if (fabs(U[1][0]) < fabs(U[m][0])) m = 1;
            if (fabs(U[2][0]) < fabs(U[m][0])) m = 2;
            double e[3] = {0, 0, 0};
            e[m] = 1.0;
            const double dp = U[m][0];
            double w[3], nn = 0;
            for (int k = 0; k < 3; ++k) { w[k] = e[k] - dp * U[k][0]; nn += w[k] * w[k]; }
            nn = sqrt(nn);
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

// one warp per start; lane 0 carries the (tiny, serial) FP64 solve.
__global__ void __launch_bounds__(32)
icp_solve_kernel(IsrIcpState *__restrict__ states, const double *__restrict__ sums,
                 int64_t ns_total, double rel_fitness, double rel_rmse, int final_eval, const PeerView px) {
    IsrIcpState &st = states[blockIdx.x];
    if (st.done != 0) return;
    __shared__ double Sx[kNS];
    if (px.world > 0) {
        // wait for message seq of every rank in THIS rank's buffer, then add in rank order
        const int s = blockIdx.x;
        bool ok = true;
        if ((int)threadIdx.x < px.world) {
            volatile unsigned long long *f = px.flag[px.rank] + peer_flag_index(px, threadIdx.x, s);
            const long long t0 = clock64();
            while (*f != px.seq) {
                if (clock64() - t0 > 20000000000ll) { ok = false; break; }  // ~10 s: a peer died
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        __threadfence_system();
        if (!ok) {
            if (threadIdx.x == 0) {
                st.done = 1;
                st.reserved = 1;  // exchange timed out
                st.fitness = CUDART_NAN;
                st.inlier_rmse = CUDART_NAN;
            }
            return;
        }
        if (threadIdx.x < kNS) {
            double t = 0.0;
            for (int r = 0; r < px.world; ++r) {
                const volatile double *d = px.data[px.rank] + peer_data_index(px, r, s);
                t += d[threadIdx.x];
            }
            Sx[threadIdx.x] = t;
        }
        __syncwarp();
    }
    if (threadIdx.x != 0) return;
    const double *S = px.world > 0 ? Sx : sums + (int64_t)blockIdx.x * kNS;
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

    // Eigen::umeyama without scaling: Sigma = (1/n) sum (t - mu_t)(s - mu_s)^T
    const double inv = 1.0 / cnt;
    const double ms[3] = {S[0] * inv, S[1] * inv, S[2] * inv};
    const double mt[3] = {S[3] * inv, S[4] * inv, S[5] * inv};
    double Sig[3][3], U[3][3], V[3][3], D[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Sig[i][j] = S[6 + 3 * i + j] * inv - mt[i] * ms[j];
    svd3(Sig, U, D, V);
    const double sgn = (det3(U) * det3(V) < 0.0) ? -1.0 : 1.0;
    double R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            R[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + sgn * U[i][2] * V[j][2];
    double tr[3];
    for (int i = 0; i < 3; ++i)
        tr[i] = mt[i] - (R[i][0] * ms[0] + R[i][1] * ms[1] + R[i][2] * ms[2]);
    // T <- [R tr; 0 1] . T
    double Tn[12];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            double v = R[i][0] * st.T[0 + j] + R[i][1] * st.T[4 + j] + R[i][2] * st.T[8 + j];
            if (j == 3) v += tr[i];
            Tn[4 * i + j] = v;
        }
    for (int k = 0; k < 12; ++k) st.T[k] = Tn[k];
    st.T[12] = 0.0; st.T[13] = 0.0; st.T[14] = 0.0; st.T[15] = 1.0;
}

// The grid (and with it the grouping of the partial sums) depends on the source size only,
// never on the batch of starts or the device: a start gives bit-identical sums whether it
// runs alone or next to 63 others.
static int acc_blocks(int64_t ns) {
    int64_t want = (ns + kAccThreads * 4 - 1) / (kAccThreads * 4);
    if (want > 2048) want = 2048;
    if (want < 1) want = 1;
    return (int)want;
}

int icp_search_impl(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                    const int32_t *src_perm, int64_t ns, const IsrCloud *tgt_cloud, const double *centroid,
                    int32_t *corr_idx, void *workspace, size_t workspace_bytes, void *stream, int repeat);

struct IcpLayout {
    size_t xs, d2, partials, tickets, hint, nnws, total;
    int nblk;
};

static IcpLayout icp_layout(int64_t ns, int64_t nt, int64_t starts) {
    IcpLayout L;
    const int64_t nsp = isr_soa_padded_len(ns);
    size_t off = 0;
    L.xs = off;       off += align256((size_t)starts * 7 * nsp * 4);
    L.d2 = off;       off += align256((size_t)starts * ns * 4);
    const int64_t max_blk = acc_blocks(ns);
    L.partials = off; off += align256((size_t)starts * max_blk * kNS * 8);
    L.tickets = off;  off += align256((size_t)starts * 4);
    L.hint = off;     off += align256((size_t)starts * nsp * 4);
    const size_t w1 = isr_nn_workspace_bytes(ns, nt, starts), w2 = isr_nn2_workspace_bytes(ns, nt, starts);
    L.nnws = off;     off += w1 > w2 ? w1 : w2;
    L.total = off;
    L.nblk = 0;
    return L;
}

}  // namespace isr

// Host side of the exchange: this rank's buffer, the peers' mappings, the message counter.
struct IsrPeer {
    int rank = 0, world = 0;
    void *own = nullptr;                       // data (kPeerDataBytes) followed by the flags
    void *mapped[ISR_PEER_MAX_RANKS] = {};     // mapped[rank] == own
    bool connected = false;
    unsigned long long seq = 0;                // number of exchanges enqueued so far
};

namespace isr {

static PeerView peer_view(const IsrPeer *p, bool next_message) {
    PeerView v{};
    if (p == nullptr) return v;  // world == 0
    for (int r = 0; r < p->world; ++r) {
        char *base = reinterpret_cast<char *>(p->mapped[r]);
        v.data[r] = reinterpret_cast<double *>(base);
        v.flag[r] = reinterpret_cast<unsigned long long *>(base + kPeerDataBytes);
    }
    v.rank = p->rank;
    v.world = p->world;
    v.seq = p->seq + (next_message ? 1 : 0);
    return v;
}

static int accumulate_corr_px(const IsrIcpState *states, int64_t starts, const float *src,
                              const float *src_lo, int64_t ns, const float *tgt, int64_t nt,
                              const int32_t *corr_idx, double max_dist, double *sums, uint8_t *inlier,
                              void *workspace, size_t workspace_bytes, const PeerView &px, void *stream,
                              int repeat = 0) {
    ISR_REQUIRE(starts >= 1 && starts <= 65535 && ns >= 1 && nt >= 1, ISR_E_SHAPE,
                "icp_accumulate_corr: bad size");
    ISR_REQUIRE(states && src && tgt && corr_idx && sums, ISR_E_INVALID_ARG,
                "icp_accumulate_corr: null pointer");
    IcpLayout L = icp_layout(ns, nt, starts);
    ISR_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, ISR_E_WORKSPACE,
                "icp: workspace %zu < %zu bytes", workspace_bytes, L.total);
    ISR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ISR_E_ALIGN,
                "icp: workspace not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    double *partials = reinterpret_cast<double *>(ws + L.partials);
    unsigned *tickets = reinterpret_cast<unsigned *>(ws + L.tickets);
    // (the kernel's last CTA leaves the tickets at zero for the next launch)
    if (!repeat) ISR_TRY(check_cuda(cudaMemsetAsync(tickets, 0, (size_t)starts * 4, st), "icp memset"));
    const int nblk = acc_blocks(ns);
    dim3 grid((unsigned)nblk, (unsigned)starts);
    ProfScope prof(kProfIcpAcc, st);
    icp_accumulate_kernel<<<grid, kAccThreads, 0, st>>>(states, src, src_lo, ns, tgt, corr_idx,
                                                        max_dist * max_dist, partials, tickets, sums,
                                                        inlier, px);
    return launched("icp_accumulate_kernel");
}

static int solve_px(IsrIcpState *states, int64_t starts, const double *sums, int64_t ns_total,
                    double rel_fitness, double rel_rmse, int final_eval, const PeerView &px, void *stream) {
    ISR_REQUIRE(starts >= 1 && states && sums, ISR_E_INVALID_ARG, "icp_solve: bad argument");
    ProfScope prof(kProfIcpSolve, (cudaStream_t)stream);
    icp_solve_kernel<<<(unsigned)starts, 32, 0, (cudaStream_t)stream>>>(
        states, sums, ns_total, rel_fitness, rel_rmse, final_eval, px);
    return launched("icp_solve_kernel");
}

}  // namespace isr

extern "C" {

int isr_peer_create(int rank, int world, IsrPeer **out, unsigned char *handle_out) {
    using namespace isr;
    static_assert(sizeof(cudaIpcMemHandle_t) <= ISR_PEER_HANDLE_BYTES, "IPC handle size");
    ISR_REQUIRE(out != nullptr && handle_out != nullptr, ISR_E_INVALID_ARG, "peer_create: null pointer");
    ISR_REQUIRE(world >= 1 && world <= ISR_PEER_MAX_RANKS && rank >= 0 && rank < world, ISR_E_INVALID_ARG,
                "peer_create: rank %d of %d (at most %d ranks)", rank, world, ISR_PEER_MAX_RANKS);
    IsrPeer *p = new IsrPeer();
    p->rank = rank;
    p->world = world;
    int s = check_cuda(cudaMalloc(&p->own, kPeerDataBytes + kPeerFlagBytes), "peer_create: cudaMalloc");
    if (s == ISR_OK) s = check_cuda(cudaMemset(p->own, 0, kPeerDataBytes + kPeerFlagBytes), "peer_create: memset");
    if (s == ISR_OK) s = check_cuda(cudaDeviceSynchronize(), "peer_create: sync");
    memset(handle_out, 0, ISR_PEER_HANDLE_BYTES);
    if (s == ISR_OK && world > 1) {
        cudaIpcMemHandle_t h;
        s = check_cuda(cudaIpcGetMemHandle(&h, p->own), "peer_create: cudaIpcGetMemHandle");
        if (s == ISR_OK) memcpy(handle_out, &h, sizeof(h));
    }
    if (s != ISR_OK) {
        if (p->own) cudaFree(p->own);
        delete p;
        return s;
    }
    p->mapped[rank] = p->own;
    *out = p;
    return ISR_OK;
}

int isr_peer_connect(IsrPeer *p, const unsigned char *handles) {
    using namespace isr;
    ISR_REQUIRE(p != nullptr && (handles != nullptr || p->world == 1), ISR_E_INVALID_ARG,
                "peer_connect: null pointer");
    ISR_REQUIRE(!p->connected, ISR_E_INVALID_ARG, "peer_connect: already connected");
    for (int r = 0; r < p->world; ++r) {
        if (r == p->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * ISR_PEER_HANDLE_BYTES, sizeof(h));
        ISR_TRY(check_cuda(cudaIpcOpenMemHandle(&p->mapped[r], h, cudaIpcMemLazyEnablePeerAccess),
                           "peer_connect: cudaIpcOpenMemHandle"));
    }
    p->connected = true;
    return ISR_OK;
}

int isr_peer_destroy(IsrPeer *p) {
    using namespace isr;
    if (p == nullptr) return ISR_OK;
    int s = check_cuda(cudaDeviceSynchronize(), "peer_destroy: sync");
    for (int r = 0; r < p->world; ++r)
        if (r != p->rank && p->mapped[r] != nullptr) cudaIpcCloseMemHandle(p->mapped[r]);
    if (p->own) cudaFree(p->own);
    delete p;
    return s;
}

size_t isr_icp_workspace_bytes(int64_t ns, int64_t nt, int64_t starts) {
    if (ns < 1 || nt < 1 || starts < 1) return 256;
    return isr::icp_layout(ns, nt, starts).total;
}

int isr_icp_search(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                   const int32_t *src_perm, int64_t ns, const IsrCloud *tgt_cloud, const double *centroid,
                   int32_t *corr_idx, void *workspace, size_t workspace_bytes, void *stream) {
    return isr::icp_search_impl(states, starts, src, src_lo, src_perm, ns, tgt_cloud, centroid, corr_idx,
                                workspace, workspace_bytes, stream, 0);
}

}  // extern "C"

namespace isr {

// `repeat` != 0: this is not the first evaluation of a loop that this library drives
// (isr_icp_run*): the hints hold the previous correspondences (no reset needed) and the
// launch order of the search is still in the workspace.
int icp_search_impl(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                    const int32_t *src_perm, int64_t ns, const IsrCloud *tgt_cloud, const double *centroid,
                    int32_t *corr_idx, void *workspace, size_t workspace_bytes, void *stream, int repeat) {
    ISR_REQUIRE(tgt_cloud != nullptr, ISR_E_INVALID_ARG, "icp: null target descriptor");
    const int64_t nt = tgt_cloud->n;
    ISR_REQUIRE(starts >= 1 && ns >= 1 && nt >= 1, ISR_E_SHAPE,
                "icp: need starts, ns, nt >= 1 (starts=%lld ns=%lld nt=%lld)", (long long)starts,
                (long long)ns, (long long)nt);
    ISR_REQUIRE(states && src && tgt_cloud->soa7 && centroid && corr_idx, ISR_E_INVALID_ARG,
                "icp: null pointer");
    ISR_REQUIRE(tgt_cloud->bstride == 0, ISR_E_SHAPE, "icp: the target cloud is shared by all starts");
    ISR_REQUIRE(starts <= 65535, ISR_E_SHAPE, "icp: starts %lld > 65535", (long long)starts);
    IcpLayout L = icp_layout(ns, nt, starts);
    ISR_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, ISR_E_WORKSPACE,
                "icp: workspace %zu < %zu bytes", workspace_bytes, L.total);
    ISR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ISR_E_ALIGN,
                "icp: workspace not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    float *xs = reinterpret_cast<float *>(ws + L.xs);
    float *d2 = reinterpret_cast<float *>(ws + L.d2);
    const int64_t nsp = isr_soa_padded_len(ns);
    const int32_t *done = &states[0].done;  // device address arithmetic only
    // source: FP64 pose from the device state, centred on the target's centroid, hi/lo planes
    ISR_TRY(isr_prepare_cloud(src, src_lo, src_perm, ns, &states[0].T[0], kStateDoubles, nullptr, 0,
                              centroid, starts, xs, nsp, done, kStateInts, stream));
    // every source point starts the search from its previous correspondence: between two ICP
    // iterations the pose moves little, so that neighbour is already a near-final bound
    int32_t *hint = reinterpret_cast<int32_t *>(ws + L.hint);
    if (!repeat) {
        const long long total = (long long)starts * nsp;
        icp_hint_reset_kernel<<<(unsigned)((total + 1023) / 1024), 256, 0, st>>>(states, nsp, total, hint);
        ISR_TRY(launched("icp_hint_reset_kernel"));
    }
    const IsrCloud src_cloud{xs, ns, nsp, 7 * nsp, nullptr, src_perm, nullptr, hint, nullptr};
    return nn2_search(&src_cloud, tgt_cloud, starts, 1, d2, corr_idx, done, kStateInts, ws + L.nnws,
                      L.total - L.nnws, stream, repeat);
}

}  // namespace isr

extern "C" {

int isr_icp_corr_dist(const IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                      int64_t ns, const float *tgt, const int32_t *corr_idx, double *out_D,
                      void *stream) {
    using namespace isr;
    ISR_REQUIRE(starts >= 1 && starts <= 65535 && ns >= 1, ISR_E_SHAPE, "icp_corr_dist: bad size");
    ISR_REQUIRE(states && src && tgt && corr_idx && out_D, ISR_E_INVALID_ARG, "icp_corr_dist: null pointer");
    dim3 grid((unsigned)acc_blocks(ns), (unsigned)starts);
    icp_corr_dist_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(states, src, src_lo, ns, tgt, corr_idx, out_D);
    return launched("icp_corr_dist_kernel");
}

int isr_icp_accumulate_corr(const IsrIcpState *states, int64_t starts, const float *src,
                            const float *src_lo, int64_t ns, const float *tgt, int64_t nt,
                            const int32_t *corr_idx, double max_dist, double *sums, uint8_t *inlier,
                            void *workspace, size_t workspace_bytes, void *stream) {
    using namespace isr;
    if (max_dist <= 0.0) {
        // upstream: a non-positive distance yields an empty result
        ISR_REQUIRE(starts >= 1 && sums, ISR_E_INVALID_ARG, "icp_accumulate_corr: bad argument");
        return check_cuda(cudaMemsetAsync(sums, 0, (size_t)starts * kNS * 8, (cudaStream_t)stream),
                          "icp memset");
    }
    return accumulate_corr_px(states, starts, src, src_lo, ns, tgt, nt, corr_idx, max_dist, sums, inlier,
                              workspace, workspace_bytes, PeerView{}, stream);
}

int isr_icp_accumulate(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                       const int32_t *src_perm, int64_t ns, const float *tgt,
                       const IsrCloud *tgt_cloud, const double *centroid, double max_dist,
                       double *sums, int32_t *corr_idx, uint8_t *inlier, void *workspace,
                       size_t workspace_bytes, void *stream) {
    using namespace isr;
    ISR_REQUIRE(tgt_cloud != nullptr && tgt != nullptr && sums != nullptr, ISR_E_INVALID_ARG,
                "icp: null pointer");
    if (max_dist > 0.0)
        ISR_TRY(isr_icp_search(states, starts, src, src_lo, src_perm, ns, tgt_cloud, centroid, corr_idx,
                               workspace, workspace_bytes, stream));
    return isr_icp_accumulate_corr(states, starts, src, src_lo, ns, tgt, tgt_cloud->n, corr_idx, max_dist,
                                   sums, inlier, workspace, workspace_bytes, stream);
}

int isr_icp_solve(IsrIcpState *states, int64_t starts, const double *sums, int64_t ns_total,
                  double rel_fitness, double rel_rmse, int final_eval, void *stream) {
    return isr::solve_px(states, starts, sums, ns_total, rel_fitness, rel_rmse, final_eval,
                         isr::PeerView{}, stream);
}

int isr_icp_run(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                const int32_t *src_perm, int64_t ns, const float *tgt, const IsrCloud *tgt_cloud,
                const double *centroid, double max_dist, int max_iteration, double rel_fitness,
                double rel_rmse, double *sums, int32_t *corr_idx, uint8_t *inlier, void *workspace,
                size_t workspace_bytes, void *stream) {
    using namespace isr;
    ISR_REQUIRE(max_iteration >= 0, ISR_E_INVALID_ARG, "icp_run: max_iteration < 0");
    ISR_REQUIRE(tgt_cloud != nullptr && tgt != nullptr && sums != nullptr, ISR_E_INVALID_ARG,
                "icp_run: null pointer");
    for (int k = 0; k <= max_iteration; ++k) {
        if (max_dist > 0.0) {
            ISR_TRY(icp_search_impl(states, starts, src, src_lo, src_perm, ns, tgt_cloud, centroid, corr_idx,
                                    workspace, workspace_bytes, stream, k > 0));
            ISR_TRY(accumulate_corr_px(states, starts, src, src_lo, ns, tgt, tgt_cloud->n, corr_idx, max_dist,
                                       sums, inlier, workspace, workspace_bytes, PeerView{}, stream, k > 0));
        } else {
            ISR_TRY(isr_icp_accumulate_corr(states, starts, src, src_lo, ns, tgt, tgt_cloud->n, corr_idx,
                                            max_dist, sums, inlier, workspace, workspace_bytes, stream));
        }
        ISR_TRY(isr_icp_solve(states, starts, sums, ns, rel_fitness, rel_rmse,
                              k == max_iteration ? 1 : 0, stream));
    }
    return ISR_OK;
}

int isr_icp_run_sharded(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                        const int32_t *src_perm, int64_t ns, int64_t ns_total, const float *tgt,
                        const IsrCloud *tgt_cloud, const double *centroid, double max_dist,
                        int max_iteration, double rel_fitness, double rel_rmse, double *sums,
                        int32_t *corr_idx, uint8_t *inlier, void *workspace, size_t workspace_bytes,
                        IsrPeer *peer, void *stream) {
    using namespace isr;
    ISR_REQUIRE(max_iteration >= 0, ISR_E_INVALID_ARG, "icp_run_sharded: max_iteration < 0");
    ISR_REQUIRE(peer != nullptr && peer->connected, ISR_E_INVALID_ARG,
                "icp_run_sharded: the peer exchange is not connected");
    ISR_REQUIRE(starts >= 1 && starts <= ISR_PEER_MAX_STARTS, ISR_E_SHAPE,
                "icp_run_sharded: starts %lld outside 1..%d", (long long)starts, ISR_PEER_MAX_STARTS);
    ISR_REQUIRE(tgt_cloud != nullptr && tgt != nullptr && sums != nullptr, ISR_E_INVALID_ARG,
                "icp_run_sharded: null pointer");
    ISR_REQUIRE(ns >= 1 && ns_total >= ns, ISR_E_SHAPE, "icp_run_sharded: ns %lld, ns_total %lld",
                (long long)ns, (long long)ns_total);
    for (int k = 0; k <= max_iteration; ++k) {
        if (max_dist > 0.0) {
            ISR_TRY(icp_search_impl(states, starts, src, src_lo, src_perm, ns, tgt_cloud, centroid, corr_idx,
                                    workspace, workspace_bytes, stream, k > 0));
            // the message number advances only once both kernels of the pair are enqueued
            const PeerView px = peer_view(peer, true);
            ISR_TRY(accumulate_corr_px(states, starts, src, src_lo, ns, tgt, tgt_cloud->n, corr_idx, max_dist,
                                       sums, inlier, workspace, workspace_bytes, px, stream, k > 0));
            const int s = solve_px(states, starts, sums, ns_total, rel_fitness, rel_rmse,
                                   k == max_iteration ? 1 : 0, px, stream);
            peer->seq += 1;  // the accumulate kernel has been launched: its message exists
            ISR_TRY(s);
        } else {
            // no correspondences anywhere: every rank's sums are zero, nothing to exchange
            ISR_TRY(check_cuda(cudaMemsetAsync(sums, 0, (size_t)starts * kNS * 8, (cudaStream_t)stream),
                               "icp memset"));
            ISR_TRY(solve_px(states, starts, sums, ns_total, rel_fitness, rel_rmse,
                             k == max_iteration ? 1 : 0, PeerView{}, stream));
        }
    }
    return ISR_OK;
}

}  // extern "C"
