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
// Two forms of one evaluation + update:
//   FUSED (isr_icp_run, isr_icp_run_sharded; the product path): ONE kernel per iteration --
//     nn2_pruned_kernel<FUSED> transforms the source itself, searches, gathers every
//     neighbour's original coordinates (16-byte rows in stored order), forms the 17 sums in a
//     launch-independent fixed order, exchanges them with the peer GPUs and solves Kabsch in
//     the last warp to arrive (icp_device.cuh).  Per iteration nothing but that one launch is
//     enqueued; nothing is read by the host.
//   STEPWISE (isr_icp_search / _corr_dist / _accumulate_corr / _solve; target-sharded ICP, the
//     NCCL exchange, the exhaustive search): K1' (FP64 transform -> hi/lo planes) + K2 +
//     icp_accumulate_kernel (gather: 12 B source + 4 B index + 12 B gathered target + 1 B
//     flag per source point) + icp_solve_kernel.
// Both re-derive every correspondence distance in FP64 from the original source point and the
// FP64 pose, so the inlier test, fitness, rmse and the 17 Kabsch sums carry no FP32 error; only
// the choice of neighbour is made in FP32.  No floating-point atomics.  A finished start sets
// `done`; all later kernels for it exit at once, so the host enqueues max_iteration + 1 passes
// without ever reading the device.
#include <math_constants.h>
#include <string.h>

#include <stdlib.h>

#include "icp_device.cuh"
#include "isr_common.cuh"

namespace isr {

constexpr int kAccThreads = 128;
constexpr int kStateInts = (int)(sizeof(IsrIcpState) / sizeof(int32_t));
constexpr int kStateDoubles = (int)(sizeof(IsrIcpState) / sizeof(double));
static_assert(sizeof(IsrIcpState) % 8 == 0, "IsrIcpState must be a whole number of doubles");

// ---- stepwise accumulate: the same sums, in the same order, as the fused search's epilogue ----
// One warp per block of 256 STORED source points (perm_q: stored position -> original index),
// grid (ceil(nqb / 4), starts) x 128 threads.  Per point: original source row (12 B + 12 B lo),
// its correspondence (4 B, original target index; < 0 = none on this rank), the gathered target
// row (12 B), one flag written.  Row sums, then icp_fused_tail with do_solve = 0: rows -> block ->
// group -> start in index order, so that the 17 sums equal the fused iteration's bit for bit.
__global__ void __launch_bounds__(kAccThreads)
icp_accumulate_rows_kernel(const __grid_constant__ IcpFuse f, const float *__restrict__ src,
                           const float *__restrict__ src_lo, const int32_t *__restrict__ perm_q, int64_t ns,
                           const float *__restrict__ tgt, const int32_t *__restrict__ idx,
                           uint8_t *__restrict__ inlier) {
    __shared__ double scratch_all[kAccThreads / 32][32 * kNS];
    const int b = blockIdx.y;
    if (f.states[b].done != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blk = blockIdx.x * (kAccThreads / 32) + warp;
    if (blk >= f.nqb) return;
    double *scratch = scratch_all[warp];
    double T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = f.states[b].T[k];
    double rs[8];
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
        const int64_t i = (int64_t)blk * 256 + r * 32 + lane;
        double c[kNS];
#pragma unroll
        for (int k = 0; k < kNS; ++k) c[k] = 0.0;
        if (i < ns) {
            const int64_t io = perm_q != nullptr ? perm_q[i] : i;
            const int j = idx[(int64_t)b * ns + io];
            bool in = false;
            if (j >= 0) {
                double px = src[3 * io], py = src[3 * io + 1], pz = src[3 * io + 2];
                if (src_lo != nullptr) {
                    px += (double)src_lo[3 * io]; py += (double)src_lo[3 * io + 1]; pz += (double)src_lo[3 * io + 2];
                }
                double sx, sy, sz;
                icp_apply_pose(T, px, py, pz, sx, sy, sz);
                const double tx = tgt[3ll * j], ty = tgt[3ll * j + 1], tz = tgt[3ll * j + 2];
                const double d2 = icp_dist2(sx, sy, sz, tx, ty, tz);
                in = d2 < f.max_d2;
                if (in) icp_contrib(sx, sy, sz, tx, ty, tz, d2, c);
            }
            if (inlier != nullptr) inlier[(int64_t)b * ns + io] = in ? 1 : 0;
        }
#pragma unroll
        for (int k = 0; k < kNS; ++k) scratch[lane * kNS + k] = c[k];
        __syncwarp();
        rs[r] = icp_row_sum(scratch, lane);
        __syncwarp();
    }
    icp_fused_tail(f, b, blk, 0xFFu, lane, rs);
}

// out_D[start][i] = FP64 squared distance between T.src[i] and tgt[idx[start][i]] (the same
// arithmetic as the accumulate kernels); +inf where idx < 0.  grid (blocks, starts).
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
            double sx, sy, sz;
            icp_apply_pose(T, px, py, pz, sx, sy, sz);
            D = icp_dist2(sx, sy, sz, (double)tgt[3ll * j], (double)tgt[3ll * j + 1], (double)tgt[3ll * j + 2]);
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

// one warp per start; lane 0 carries the (tiny, serial) FP64 solve (icp_device.cuh).
__global__ void __launch_bounds__(32)
icp_solve_kernel(IsrIcpState *__restrict__ states, const double *__restrict__ sums,
                 int64_t ns_total, double rel_fitness, double rel_rmse, int final_eval) {
    IsrIcpState &st = states[blockIdx.x];
    if (st.done != 0 || threadIdx.x != 0) return;
    icp_solve_state(st, sums + (int64_t)blockIdx.x * kNS, ns_total, rel_fitness, rel_rmse, final_eval);
}

// tgt4[i] = original coordinates of the target point stored at position i (perm NULL: i itself);
// padded slots hold zeros (never referenced: the search reports real points only)
__global__ void __launch_bounds__(256)
icp_gather_tgt4_kernel(const float *__restrict__ tgt, const int32_t *__restrict__ perm, int64_t nt,
                       int64_t ntp, float4 *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= ntp) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nt) {
        const int64_t j = perm != nullptr ? perm[i] : i;
        v = make_float4(tgt[3 * j], tgt[3 * j + 1], tgt[3 * j + 2], 0.f);
    }
    out[i] = v;
}

// after the fused loop: correspondences and flags from stored order back to the caller's indexing
__global__ void __launch_bounds__(256)
icp_unpermute_kernel(const int32_t *__restrict__ hint, const uint8_t *__restrict__ inl_stored,
                     const int32_t *__restrict__ perm_q, const int32_t *__restrict__ perm_t, int64_t ns,
                     int64_t nsp, int64_t nt, int32_t *__restrict__ corr_idx, uint8_t *__restrict__ inlier) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t s = blockIdx.y;
    if (i >= ns) return;
    const int64_t io = perm_q != nullptr ? perm_q[i] : i;
    int j = hint[s * nsp + i];
    j = j < 0 ? 0 : (j >= nt ? (int)(nt - 1) : j);
    corr_idx[s * ns + io] = perm_t != nullptr ? perm_t[j] : j;
    if (inlier != nullptr) inlier[s * ns + io] = inl_stored[s * nsp + i];
}

// grid of the exact-distance kernel
static int acc_blocks(int64_t ns) {
    int64_t want = (ns + 256 * 4 - 1) / (256 * 4);
    if (want > 2048) want = 2048;
    if (want < 1) want = 1;
    return (int)want;
}

int icp_search_impl(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                    const int32_t *src_perm, int64_t ns, const IsrCloud *tgt_cloud, const double *centroid,
                    int32_t *corr_idx, void *workspace, size_t workspace_bytes, void *stream, int repeat);

struct IcpLayout {
    size_t xs, d2, hint, nnws;                    // stepwise search (hint, nnws: both forms)
    size_t blocksum, groupsum, ftick;             // the fixed-order reduction of the 17 sums (both)
    size_t src7, tgt4, inl, rowsum;               // fused iteration
    size_t total;
    int nqb, ngroups;
};

static IcpLayout icp_layout(int64_t ns, int64_t nt, int64_t starts) {
    IcpLayout L;
    const int64_t nsp = isr_soa_padded_len(ns), ntp = isr_soa_padded_len(nt);
    size_t off = 0;
    L.xs = off;       off += align256((size_t)starts * 7 * nsp * 4);
    L.d2 = off;       off += align256((size_t)starts * ns * 4);
    L.hint = off;     off += align256((size_t)starts * nsp * 4);
    L.nqb = nn2_query_blocks(ns);
    L.ngroups = (L.nqb + kFuseGroup - 1) / kFuseGroup;
    L.blocksum = off; off += align256((size_t)starts * L.nqb * kNS * 8);
    L.groupsum = off; off += align256((size_t)starts * L.ngroups * kNS * 8);
    L.ftick = off;    off += align256((size_t)starts * (L.nqb + L.ngroups + 1) * 4);
    L.src7 = off;     off += align256((size_t)7 * nsp * 4);
    L.tgt4 = off;     off += align256((size_t)ntp * 16);
    L.inl = off;      off += align256((size_t)starts * nsp);
    L.rowsum = off;   off += align256((size_t)starts * L.nqb * 8 * kNS * 8);
    const size_t w1 = isr_nn_workspace_bytes(ns, nt, starts), w2 = isr_nn2_workspace_bytes(ns, nt, starts);
    L.nnws = off;     off += w1 > w2 ? w1 : w2;
    L.total = off;
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
    long long timeout_cycles = 20000000000ll;  // ~10 s of SM clock
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
    v.timeout_cycles = p->timeout_cycles;
    return v;
}

static int check_workspace(const IcpLayout &L, const void *workspace, size_t workspace_bytes) {
    ISR_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, ISR_E_WORKSPACE,
                "icp: workspace %zu < %zu bytes", workspace_bytes, L.total);
    ISR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ISR_E_ALIGN,
                "icp: workspace not 256-byte aligned");
    return ISR_OK;
}

static int accumulate_corr(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                           const int32_t *src_perm, int64_t ns, const float *tgt, int64_t nt,
                           const int32_t *corr_idx, double max_dist, double *sums, uint8_t *inlier,
                           void *workspace, size_t workspace_bytes, void *stream, int repeat = 0) {
    ISR_REQUIRE(starts >= 1 && starts <= 65535 && ns >= 1 && nt >= 1, ISR_E_SHAPE,
                "icp_accumulate_corr: bad size");
    ISR_REQUIRE(states && src && tgt && corr_idx && sums, ISR_E_INVALID_ARG,
                "icp_accumulate_corr: null pointer");
    IcpLayout L = icp_layout(ns, nt, starts);
    ISR_TRY(check_workspace(L, workspace, workspace_bytes));
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    IcpFuse f{};
    f.states = states;
    f.max_d2 = max_dist * max_dist;
    f.blocksum = reinterpret_cast<double *>(ws + L.blocksum);
    f.groupsum = reinterpret_cast<double *>(ws + L.groupsum);
    f.tickets = reinterpret_cast<unsigned *>(ws + L.ftick);
    f.sums = sums;
    f.nqb = L.nqb;
    f.ngroups = L.ngroups;
    f.do_solve = 0;
    // (the last warp of every reduction level leaves its ticket at zero for the next launch)
    if (!repeat)
        ISR_TRY(check_cuda(cudaMemsetAsync(f.tickets, 0, (size_t)starts * (L.nqb + L.ngroups + 1) * 4, st),
                           "icp memset"));
    constexpr int kWarps = kAccThreads / 32;
    dim3 grid((unsigned)((L.nqb + kWarps - 1) / kWarps), (unsigned)starts);
    ProfScope prof(kProfIcpAcc, st);
    icp_accumulate_rows_kernel<<<grid, kAccThreads, 0, st>>>(f, src, src_lo, src_perm, ns, tgt, corr_idx, inlier);
    return launched("icp_accumulate_rows_kernel");
}

// The loop of isr_icp_run / isr_icp_run_sharded, one fused launch per evaluation.  `peer` NULL:
// no exchange (single GPU).
static int run_fused(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                     const int32_t *src_perm, int64_t ns, int64_t ns_total, const float *tgt,
                     const IsrCloud *tgt_cloud, const double *centroid, double max_dist, int max_iteration,
                     double rel_fitness, double rel_rmse, double *sums, int32_t *corr_idx, uint8_t *inlier,
                     void *workspace, size_t workspace_bytes, IsrPeer *peer, void *stream) {
    const int64_t nt = tgt_cloud->n;
    ISR_REQUIRE(starts >= 1 && starts <= 65535 && ns >= 1 && nt >= 1, ISR_E_SHAPE, "icp_run: bad size");
    ISR_REQUIRE(states && src && centroid && corr_idx, ISR_E_INVALID_ARG, "icp_run: null pointer");
    ISR_REQUIRE(tgt_cloud->bstride == 0, ISR_E_SHAPE, "icp: the target cloud is shared by all starts");
    const IcpLayout L = icp_layout(ns, nt, starts);
    ISR_TRY(check_workspace(L, workspace, workspace_bytes));
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    const int64_t nsp = isr_soa_padded_len(ns), ntp = tgt_cloud->npad;
    float *src7 = reinterpret_cast<float *>(ws + L.src7);
    float4 *tgt4 = reinterpret_cast<float4 *>(ws + L.tgt4);
    int32_t *hint = reinterpret_cast<int32_t *>(ws + L.hint);
    uint8_t *inl = reinterpret_cast<uint8_t *>(ws + L.inl);
    unsigned *ftick = reinterpret_cast<unsigned *>(ws + L.ftick);

    // once per run: the source as stored-order hi/lo planes (identity pose, no centring), the
    // target's original coordinates as 16-byte rows in stored order, empty hints, zero tickets
    ISR_TRY(isr_prepare_cloud(src, src_lo, src_perm, ns, nullptr, 16, nullptr, 16, nullptr, 1, src7, nsp,
                              nullptr, 0, stream));
    icp_gather_tgt4_kernel<<<(unsigned)((ntp + 255) / 256), 256, 0, st>>>(tgt, tgt_cloud->perm, nt, ntp, tgt4);
    ISR_TRY(launched("icp_gather_tgt4_kernel"));
    {
        const long long total = (long long)starts * nsp;
        icp_hint_reset_kernel<<<(unsigned)((total + 1023) / 1024), 256, 0, st>>>(states, nsp, total, hint);
        ISR_TRY(launched("icp_hint_reset_kernel"));
    }
    ISR_TRY(check_cuda(cudaMemsetAsync(ftick, 0, (size_t)starts * (L.nqb + L.ngroups + 1) * 4, st), "icp memset"));

    IcpFuse f{};
    f.states = states;
    f.src7 = src7;
    f.tgt4 = tgt4;
    f.centroid = centroid;
    f.max_d2 = max_dist * max_dist;
    f.rowsum = reinterpret_cast<double *>(ws + L.rowsum);
    f.blocksum = reinterpret_cast<double *>(ws + L.blocksum);
    f.groupsum = reinterpret_cast<double *>(ws + L.groupsum);
    f.tickets = ftick;
    f.sums = sums;
    f.inlier = inl;
    f.nqb = L.nqb;
    f.ngroups = L.ngroups;
    f.ns_total = ns_total;
    f.rel_fitness = rel_fitness;
    f.rel_rmse = rel_rmse;
    f.do_solve = 1;
    // the query "cloud" of the fused search: the original source planes, shared by all starts
    const IsrCloud src_cloud{src7, ns, nsp, 0, nullptr, src_perm, nullptr, hint, nullptr};
    const int32_t *done = &states[0].done;  // device address arithmetic only
    for (int k = 0; k <= max_iteration; ++k) {
        f.final_eval = k == max_iteration ? 1 : 0;
        f.px = peer_view(peer, true);  // world == 0 without a peer
        // (iterations 2 and 4 re-cut the launch list from the costs that the hinted iterations before
        // them measured: nn2.cu, block_rebalance_kernel; iteration 0 is the unhinted search)
        static const int recut_rounds = getenv("ISR_ICP_RECUT_ROUNDS") ? atoi(getenv("ISR_ICP_RECUT_ROUNDS")) : 2;
        // (a grid deeper than one wave is throughput-bound: one re-cut is all it gains from)
        const int rounds = nn2_query_blocks(ns) * starts > (int64_t)sm_count() * 20 ? (recut_rounds < 1 ? recut_rounds : 1) : recut_rounds;
        const int reuse = k == 0 ? 0 : (k % 2 == 0 && k <= 2 * rounds) ? 2 : 1;
        const int s = nn2_search(&src_cloud, tgt_cloud, starts, 1, nullptr, nullptr, done, kStateInts,
                                 ws + L.nnws, L.total - L.nnws, stream, reuse, &f);
        if (peer != nullptr && s == ISR_OK) peer->seq += 1;  // the launch exists: so does its message
        ISR_TRY(s);
    }
    dim3 grid((unsigned)((ns + 255) / 256), (unsigned)starts);
    icp_unpermute_kernel<<<grid, 256, 0, st>>>(hint, inl, src_perm, tgt_cloud->perm, ns, nsp, nt, corr_idx, inlier);
    return launched("icp_unpermute_kernel");
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

int isr_peer_set_timeout(IsrPeer *p, double seconds) {
    using namespace isr;
    ISR_REQUIRE(p != nullptr && seconds > 0.0, ISR_E_INVALID_ARG, "peer_set_timeout: bad argument");
    int khz = 0;
    ISR_TRY(isr_device_info(nullptr, &khz, nullptr));
    const double cycles = seconds * (double)khz * 1e3;
    p->timeout_cycles = cycles > 9e18 ? (long long)9e18 : (long long)cycles;
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

// `repeat` != 0: this is not the first evaluation of a loop that this library drives: the hints
// hold the previous correspondences (no reset needed) and the launch order of the search is
// still in the workspace.
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
    ISR_TRY(check_workspace(L, workspace, workspace_bytes));
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
                      L.total - L.nnws, stream, repeat, nullptr);
}

// The stepwise loop (exhaustive search, or a target without tile spheres): K1' + K2 + accumulate
// + solve per evaluation.
static int run_stepwise(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                        const int32_t *src_perm, int64_t ns, const float *tgt, const IsrCloud *tgt_cloud,
                        const double *centroid, double max_dist, int max_iteration, double rel_fitness,
                        double rel_rmse, double *sums, int32_t *corr_idx, uint8_t *inlier, void *workspace,
                        size_t workspace_bytes, void *stream) {
    for (int k = 0; k <= max_iteration; ++k) {
        ISR_TRY(icp_search_impl(states, starts, src, src_lo, src_perm, ns, tgt_cloud, centroid, corr_idx,
                                workspace, workspace_bytes, stream, k > 0));
        ISR_TRY(accumulate_corr(states, starts, src, src_lo, src_perm, ns, tgt, tgt_cloud->n, corr_idx, max_dist,
                                sums, inlier, workspace, workspace_bytes, stream, k > 0));
        ISR_TRY(isr_icp_solve(states, starts, sums, ns, rel_fitness, rel_rmse, k == max_iteration ? 1 : 0,
                              stream));
    }
    return ISR_OK;
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

int isr_icp_accumulate_corr(IsrIcpState *states, int64_t starts, const float *src, const float *src_lo,
                            const int32_t *src_perm, int64_t ns, const float *tgt, int64_t nt,
                            const int32_t *corr_idx, double max_dist, double *sums, uint8_t *inlier,
                            void *workspace, size_t workspace_bytes, void *stream) {
    using namespace isr;
    if (max_dist <= 0.0) {
        // upstream: a non-positive distance yields an empty result
        ISR_REQUIRE(starts >= 1 && sums, ISR_E_INVALID_ARG, "icp_accumulate_corr: bad argument");
        return check_cuda(cudaMemsetAsync(sums, 0, (size_t)starts * kNS * 8, (cudaStream_t)stream),
                          "icp memset");
    }
    return accumulate_corr(states, starts, src, src_lo, src_perm, ns, tgt, nt, corr_idx, max_dist, sums, inlier,
                           workspace, workspace_bytes, stream);
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
    return isr_icp_accumulate_corr(states, starts, src, src_lo, src_perm, ns, tgt, tgt_cloud->n, corr_idx,
                                   max_dist, sums, inlier, workspace, workspace_bytes, stream);
}

int isr_icp_solve(IsrIcpState *states, int64_t starts, const double *sums, int64_t ns_total,
                  double rel_fitness, double rel_rmse, int final_eval, void *stream) {
    using namespace isr;
    ISR_REQUIRE(starts >= 1 && states && sums, ISR_E_INVALID_ARG, "icp_solve: bad argument");
    ProfScope prof(kProfIcpSolve, (cudaStream_t)stream);
    icp_solve_kernel<<<(unsigned)starts, 32, 0, (cudaStream_t)stream>>>(states, sums, ns_total, rel_fitness,
                                                                        rel_rmse, final_eval);
    return launched("icp_solve_kernel");
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
    if (max_dist <= 0.0) {
        // upstream: a non-positive distance yields an empty correspondence set in every evaluation
        for (int k = 0; k <= max_iteration; ++k) {
            ISR_TRY(isr_icp_accumulate_corr(states, starts, src, src_lo, src_perm, ns, tgt, tgt_cloud->n,
                                            corr_idx, max_dist, sums, inlier, workspace, workspace_bytes,
                                            stream));
            ISR_TRY(isr_icp_solve(states, starts, sums, ns, rel_fitness, rel_rmse,
                                  k == max_iteration ? 1 : 0, stream));
        }
        return ISR_OK;
    }
    if (nn2_fusable(tgt_cloud))
        return run_fused(states, starts, src, src_lo, src_perm, ns, ns, tgt, tgt_cloud, centroid, max_dist,
                         max_iteration, rel_fitness, rel_rmse, sums, corr_idx, inlier, workspace,
                         workspace_bytes, nullptr, stream);
    return run_stepwise(states, starts, src, src_lo, src_perm, ns, tgt, tgt_cloud, centroid, max_dist,
                        max_iteration, rel_fitness, rel_rmse, sums, corr_idx, inlier, workspace,
                        workspace_bytes, stream);
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
    if (max_dist <= 0.0) {
        // no correspondences anywhere: every rank's sums are zero, nothing to exchange
        for (int k = 0; k <= max_iteration; ++k) {
            ISR_TRY(check_cuda(cudaMemsetAsync(sums, 0, (size_t)starts * kNS * 8, (cudaStream_t)stream),
                               "icp memset"));
            ISR_TRY(isr_icp_solve(states, starts, sums, ns_total, rel_fitness, rel_rmse,
                                  k == max_iteration ? 1 : 0, stream));
        }
        return ISR_OK;
    }
    ISR_REQUIRE(nn2_fusable(tgt_cloud), ISR_E_INVALID_ARG,
                "icp_run_sharded: needs the pruned search (target tile spheres, isr_set_nn_pruning(1))");
    return run_fused(states, starts, src, src_lo, src_perm, ns, ns_total, tgt, tgt_cloud, centroid, max_dist,
                     max_iteration, rel_fitness, rel_rmse, sums, corr_idx, inlier, workspace, workspace_bytes,
                     peer, stream);
}

}  // extern "C"
