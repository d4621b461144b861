// Batched candidate-pose verification: the hot loop of verfication.py:61-108 (Chamfer)
// and the ADDS scoring of choosePose.py:20-22,124-138, run for a whole candidate batch.
//
// Per chunk of C candidates:  K1 writes the two transformed clouds as SoA planes
// (C x 2.4 MB at 100k points -- L2-resident by the time K2 reads them), K2 runs both
// search directions batched over the chunk, one CTA per (direction, candidate) reduces
// mean sqrt(d2) in FP64 with a fixed order, and a finalize kernel writes the loss.
// The selection is a single-CTA first-minimum argmin (list.index(min(...)),
// verfication.py:105-106).  No float atomics anywhere: results are run-to-run identical.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "isr_common.cuh"

namespace isr {

__global__ void verify_finalize_kernel(const double *__restrict__ mean_a,
                                       const double *__restrict__ mean_b,
                                       const uint8_t *__restrict__ valid, int64_t offset, int c,
                                       int bidirectional, double *__restrict__ out_loss) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    double l = bidirectional ? (mean_a[i] + mean_b[i]) / 2 : mean_a[i];
    if (valid != nullptr && valid[offset + i] == 0) l = CUDART_INF;
    out_loss[offset + i] = l;
}

// single CTA; first minimum wins; NaN never wins.
__global__ void __launch_bounds__(1024)
argmin_first_kernel(const double *__restrict__ loss, int64_t n, int64_t *__restrict__ out_best) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    double bv = CUDART_INF;
    long long bi = -1;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = loss[i];
        if (v < bv || (bi < 0 && v == bv)) { bv = v; bi = i; }  // ascending i per thread
    }
    auto better = [](double v, long long i, double w, long long j) {
        if (j < 0) return false;
        if (i < 0) return true;
        return (w < v) || (w == v && j < i);
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, bv, o);
        const long long j = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(bv, bi, w, j)) { bv = w; bi = j; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = (blockDim.x + 31) >> 5;
        bv = threadIdx.x < nw ? sv[threadIdx.x] : CUDART_INF;
        bi = threadIdx.x < nw ? si[threadIdx.x] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_xor_sync(0xffffffffu, bv, o);
            const long long j = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(bv, bi, w, j)) { bv = w; bi = j; }
        }
        if (threadIdx.x == 0) {
            out_best[0] = bi;
            out_best[1] = __double_as_longlong(bv);
        }
    }
}

static int verify_chunk(int64_t nq, int64_t nt, int64_t b) {
    // cap the SoA scratch near 4 GiB and the grid z-dimension; at least 1
    const int64_t nqp = isr_soa_padded_len(nq), ntp = isr_soa_padded_len(nt);
    const int64_t per = (nqp + ntp) * 28;
    int64_t c = (int64_t(4) << 30) / per;
    if (c > 256) c = 256;
    if (c > b) c = b;
    if (c < 1) c = 1;
    return (int)c;
}

struct VerifyLayout {
    size_t xs, ys, d2a, d2b, means, centroid, perm_q, perm_t, stage_x, stage_y, sub_x, sub_y, box_x, box_y, sortws, nnws, total;
    int chunk;
};

// 1 = direct-difference K2 (nn.cu, 3 planes), 0 = filtered exact K2 (nn2.cu, 7 planes)
bool nn_mode_direct() {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("ISR_NN_MODE");
        mode = (e != nullptr && strcmp(e, "direct") == 0) ? 1 : 0;
    }
    return mode == 1;
}

static VerifyLayout verify_layout(int64_t nq, int64_t nt, int64_t b, int bidirectional) {
    VerifyLayout L;
    L.chunk = verify_chunk(nq, nt, b);
    const int64_t nqp = isr_soa_padded_len(nq), ntp = isr_soa_padded_len(nt);
    const int planes = 7;  // sized for the larger layout so either mode fits
    const int64_t big = nq > nt ? nq : nt;
    size_t off = 0;
    L.xs = off;   off += align256((size_t)L.chunk * planes * nqp * 4);
    L.ys = off;   off += align256((size_t)L.chunk * planes * ntp * 4);
    L.d2a = off;  off += align256((size_t)L.chunk * nq * 4);
    L.d2b = off;  off += bidirectional ? align256((size_t)L.chunk * nt * 4) : 0;
    L.means = off; off += align256((size_t)2 * L.chunk * 8);
    L.centroid = off; off += 256;
    L.perm_q = off; off += align256((size_t)nq * 4);
    L.perm_t = off; off += align256((size_t)nt * 4);
    L.stage_x = off; off += align256((size_t)L.chunk * isr_stage_sphere_count(nqp) * 16);
    L.stage_y = off; off += align256((size_t)L.chunk * isr_stage_sphere_count(ntp) * 16);
    L.sub_x = off; off += align256((size_t)L.chunk * (nqp / ISR_SUB_TILE) * 16);
    L.sub_y = off; off += align256((size_t)L.chunk * (ntp / ISR_SUB_TILE) * 16);
    L.box_x = off; off += align256((size_t)L.chunk * (nqp / ISR_SUB_TILE) * 4);
    L.box_y = off; off += align256((size_t)L.chunk * (ntp / ISR_SUB_TILE) * 4);
    L.sortws = off; off += isr_spatial_order_workspace_bytes(big);
    const size_t w1 = isr_nn_workspace_bytes(big, big, L.chunk);
    const size_t w2 = isr_nn2_workspace_bytes(big, big, L.chunk);
    L.nnws = off; off += w1 > w2 ? w1 : w2;
    L.total = off;
    return L;
}

}  // namespace isr

extern "C" {

size_t isr_verify_workspace_bytes(int64_t nq, int64_t nt, int64_t b, int bidirectional) {
    if (nq < 1 || nt < 1 || b < 1) return 256;
    return isr::verify_layout(nq, nt, b, bidirectional).total;
}

int isr_verify_poses(const float *cloud_q, int64_t nq, const float *cloud_t, int64_t nt,
                     const double *poses_q, const double *poses_t, const uint8_t *valid, int64_t b,
                     int bidirectional, double *out_loss, int64_t *out_best, void *workspace,
                     size_t workspace_bytes, void *stream) {
    using namespace isr;
    ISR_REQUIRE(nq >= 1 && nt >= 1 && b >= 1, ISR_E_SHAPE,
                "verify: need nq, nt, b >= 1 (nq=%lld nt=%lld b=%lld)", (long long)nq,
                (long long)nt, (long long)b);
    ISR_REQUIRE(cloud_q && cloud_t && poses_q && poses_t && out_loss, ISR_E_INVALID_ARG,
                "verify: null pointer");
    const VerifyLayout L = verify_layout(nq, nt, b, bidirectional);
    ISR_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, ISR_E_WORKSPACE,
                "verify: workspace %zu < %zu bytes", workspace_bytes, L.total);
    ISR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ISR_E_ALIGN,
                "verify: workspace not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    float *xs = reinterpret_cast<float *>(ws + L.xs);
    float *ys = reinterpret_cast<float *>(ws + L.ys);
    float *d2a = reinterpret_cast<float *>(ws + L.d2a);
    float *d2b = reinterpret_cast<float *>(ws + L.d2b);
    double *means = reinterpret_cast<double *>(ws + L.means);
    void *nnws = ws + L.nnws;
    const size_t nnws_bytes = L.total - L.nnws;
    const int64_t nqp = isr_soa_padded_len(nq), ntp = isr_soa_padded_len(nt);

    const bool direct = nn_mode_direct();
    double *centroid = reinterpret_cast<double *>(ws + L.centroid);
    int32_t *perm_q = reinterpret_cast<int32_t *>(ws + L.perm_q);
    int32_t *perm_t = reinterpret_cast<int32_t *>(ws + L.perm_t);
    float *stage_x = reinterpret_cast<float *>(ws + L.stage_x);
    float *stage_y = reinterpret_cast<float *>(ws + L.stage_y);
    float *sub_x = reinterpret_cast<float *>(ws + L.sub_x);
    float *sub_y = reinterpret_cast<float *>(ws + L.sub_y);
    uint32_t *box_x = reinterpret_cast<uint32_t *>(ws + L.box_x);
    uint32_t *box_y = reinterpret_cast<uint32_t *>(ws + L.box_y);
    if (!direct) {
        ISR_TRY(isr_centroid(cloud_t, nt, centroid, stream));
        // Hilbert order once per cloud; every candidate's rigid copy inherits the coherence.
        // The losses are means over all points, so nothing has to be permuted back.
        ISR_TRY(isr_spatial_order(cloud_t, nt, perm_t, ws + L.sortws, L.nnws - L.sortws, stream));
        if (cloud_q == cloud_t && nq == nt) perm_q = perm_t;
        else ISR_TRY(isr_spatial_order(cloud_q, nq, perm_q, ws + L.sortws, L.nnws - L.sortws, stream));
    }
    const IsrCloud cx{xs, nq, nqp, 7 * nqp, stage_x, nullptr, sub_x, nullptr, box_x};
    const IsrCloud cy{ys, nt, ntp, 7 * ntp, stage_y, nullptr, sub_y, nullptr, box_y};
    for (int64_t k0 = 0; k0 < b; k0 += L.chunk) {
        const int c = (int)((b - k0) < L.chunk ? (b - k0) : L.chunk);
        if (direct) {
            ISR_TRY(isr_transform_points_soa(cloud_q, nq, poses_q + k0 * 16, 16, c, xs, nqp, nullptr,
                                             0, stream));
            ISR_TRY(isr_transform_points_soa(cloud_t, nt, poses_t + k0 * 16, 16, c, ys, ntp, nullptr,
                                             0, stream));
            ISR_TRY(isr_nn_soa(xs, nq, nqp, 3 * nqp, ys, nt, ntp, 3 * ntp, c, d2a, nullptr, nullptr,
                               0, nnws, nnws_bytes, stream));
        } else {
            // both clouds of candidate k are centred on c_k = Pt_k . centroid(cloud_t)
            ISR_TRY(isr_prepare_cloud(cloud_q, nullptr, perm_q, nq, poses_q + k0 * 16, 16,
                                      poses_t + k0 * 16, 16, centroid, c, xs, nqp, nullptr, 0, stream));
            ISR_TRY(isr_prepare_cloud(cloud_t, nullptr, perm_t, nt, poses_t + k0 * 16, 16,
                                      poses_t + k0 * 16, 16, centroid, c, ys, ntp, nullptr, 0, stream));
            ISR_TRY(isr_tile_spheres(xs, nq, nqp, 7 * nqp, c, stage_x, sub_x, box_x, stream));
            ISR_TRY(isr_tile_spheres(ys, nt, ntp, 7 * ntp, c, stage_y, sub_y, box_y, stream));
            ISR_TRY(isr_nn2(&cx, &cy, c, 0, d2a, nullptr, nullptr, 0, nnws, nnws_bytes, stream));
        }
        ISR_TRY(isr_mean_sqrt(d2a, nq, c, means, stream));
        if (bidirectional) {
            if (direct)
                ISR_TRY(isr_nn_soa(ys, nt, ntp, 3 * ntp, xs, nq, nqp, 3 * nqp, c, d2b, nullptr,
                                   nullptr, 0, nnws, nnws_bytes, stream));
            else
                ISR_TRY(isr_nn2(&cy, &cx, c, 0, d2b, nullptr, nullptr, 0, nnws, nnws_bytes, stream));
            ISR_TRY(isr_mean_sqrt(d2b, nt, c, means + L.chunk, stream));
        }
        verify_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(means, means + L.chunk, valid, k0, c,
                                                                bidirectional, out_loss);
        ISR_TRY(launched("verify_finalize_kernel"));
    }
    if (out_best != nullptr) {
        argmin_first_kernel<<<1, 1024, 0, st>>>(out_loss, b, out_best);
        ISR_TRY(launched("argmin_first_kernel"));
    }
    return ISR_OK;
}


// ---- ADD-S with the target prepared once (choosePose.py:124-138 for a whole table of pairs) ----
// Candidate k scores M_k . cloud_q against cloud_t ITSELF (one-directional mean 1-NN distance).
// With M_k = Pt_k^-1 Pq_k (isr_rigid_relative) this equals ADDS(verts, gtR, gtT, R, T) of
// choosePose.py:20-22 whenever (R, T) is a rigid motion -- distances do not change when both
// clouds are moved by Pt^-1 -- but the 100k-point surface is sorted, centred, split into hi/lo
// planes and given its tile spheres ONCE instead of once per pose pair; per pair only the (small)
// vertex cloud is transformed.  The target planes and spheres stay L2-resident across the batch.
size_t isr_adds_fixed_target_workspace_bytes(int64_t nq, int64_t nt, int64_t b) {
    if (nq < 1 || nt < 1 || b < 1) return 256;
    return isr::verify_layout(nq, nt, b, 0).total;
}

int isr_adds_fixed_target(const float *cloud_q, int64_t nq, const float *cloud_t, int64_t nt,
                          const double *poses_q, const uint8_t *valid, int64_t b, double *out_loss,
                          int64_t *out_best, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace isr;
    ISR_REQUIRE(nq >= 1 && nt >= 1 && b >= 1, ISR_E_SHAPE,
                "adds_fixed_target: need nq, nt, b >= 1 (nq=%lld nt=%lld b=%lld)", (long long)nq,
                (long long)nt, (long long)b);
    ISR_REQUIRE(cloud_q && cloud_t && poses_q && out_loss, ISR_E_INVALID_ARG, "adds_fixed_target: null pointer");
    const VerifyLayout L = verify_layout(nq, nt, b, 0);
    ISR_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, ISR_E_WORKSPACE,
                "adds_fixed_target: workspace %zu < %zu bytes", workspace_bytes, L.total);
    ISR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ISR_E_ALIGN,
                "adds_fixed_target: workspace not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    float *xs = reinterpret_cast<float *>(ws + L.xs);
    float *ys = reinterpret_cast<float *>(ws + L.ys);
    float *d2a = reinterpret_cast<float *>(ws + L.d2a);
    double *means = reinterpret_cast<double *>(ws + L.means);
    double *centroid = reinterpret_cast<double *>(ws + L.centroid);
    int32_t *perm_q = reinterpret_cast<int32_t *>(ws + L.perm_q);
    int32_t *perm_t = reinterpret_cast<int32_t *>(ws + L.perm_t);
    float *stage_y = reinterpret_cast<float *>(ws + L.stage_y);
    float *sub_y = reinterpret_cast<float *>(ws + L.sub_y);
    uint32_t *box_y = reinterpret_cast<uint32_t *>(ws + L.box_y);
    const int64_t nqp = isr_soa_padded_len(nq), ntp = isr_soa_padded_len(nt);
    // the target, once: centroid, curve order, centred hi/lo planes (identity pose), tile spheres
    ISR_TRY(isr_centroid(cloud_t, nt, centroid, stream));
    ISR_TRY(isr_spatial_order(cloud_t, nt, perm_t, ws + L.sortws, L.nnws - L.sortws, stream));
    ISR_TRY(isr_spatial_order(cloud_q, nq, perm_q, ws + L.sortws, L.nnws - L.sortws, stream));
    ISR_TRY(isr_prepare_cloud(cloud_t, nullptr, perm_t, nt, nullptr, 16, nullptr, 16, centroid, 1, ys, ntp,
                              nullptr, 0, stream));
    ISR_TRY(isr_tile_spheres(ys, nt, ntp, 0, 1, stage_y, sub_y, box_y, stream));
    const IsrCloud cx{xs, nq, nqp, 7 * nqp, nullptr, nullptr, nullptr, nullptr, nullptr};
    const IsrCloud cy{ys, nt, ntp, 0, stage_y, nullptr, sub_y, nullptr, box_y};
    for (int64_t k0 = 0; k0 < b; k0 += L.chunk) {
        const int c = (int)((b - k0) < L.chunk ? (b - k0) : L.chunk);
        // queries: M_k . cloud_q, centred on the (fixed) target centroid
        ISR_TRY(isr_prepare_cloud(cloud_q, nullptr, perm_q, nq, poses_q + k0 * 16, 16, nullptr, 16, centroid, c,
                                  xs, nqp, nullptr, 0, stream));
        ISR_TRY(isr_nn2(&cx, &cy, c, 0, d2a, nullptr, nullptr, 0, ws + L.nnws, L.total - L.nnws, stream));
        ISR_TRY(isr_mean_sqrt(d2a, nq, c, means, stream));
        verify_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(means, means, valid, k0, c, 0, out_loss);
        ISR_TRY(launched("verify_finalize_kernel"));
    }
    if (out_best != nullptr) {
        argmin_first_kernel<<<1, 1024, 0, st>>>(out_loss, b, out_best);
        ISR_TRY(launched("argmin_first_kernel"));
    }
    return ISR_OK;
}

}  // extern "C"
