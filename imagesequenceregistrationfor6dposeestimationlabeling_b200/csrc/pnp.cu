// PnP hypothesis scoring -- HBM/L2-bound.
//
// The consensus test inside cv2.solvePnPRansac as the reference calls it (choosePose.py:23-33,
// 280-300: P3P hypotheses, reprojectionError = 2 px, no distortion): project every 3-D point of
// the 2-D/3-D correspondence set with a hypothesis (R, t) and the camera matrix, and count the
// correspondences whose squared reprojection error is <= reperr^2.  OpenCV's restated
// arithmetic (calib3d, PnPRansacCallback::computeError + RANSACPointSetRegistrator::findInliers;
// pinned against cv2 4.13's own projectPoints output, tests/golden/reference_pnp_cv2.npz): pinhole
// projection in FP64 stored as float32, z == 0 replaced by 1, float32 differences against the
// float32 image points, error = float32(double(du)^2 + double(dv)^2), inlier iff
// double(error) <= reperr * reperr.
//
// One launch scores b hypotheses against the same n correspondences: grid (ceil(n / 1024), b
// / 16); a thread keeps 4 correspondences (20 B each, read once, L2-resident across the grid's
// y) in registers and walks 16 hypotheses staged in shared memory; counts are reduced with
// __syncthreads_count-style ballots and one integer atomicAdd per CTA and hypothesis
// (integer: order-independent).  Algorithmic bytes per launch: n * 20 read + b * 4 written
// (+ b * n optional inlier flags).
#include "isr_common.cuh"

namespace isr {

constexpr int kPnpThreads = 256;
constexpr int kPnpPoses = 16;
constexpr int kPnpPts = 4;

__global__ void __launch_bounds__(kPnpThreads)
pnp_score_kernel(const float *__restrict__ p3d, const float *__restrict__ p2d, int64_t n,
                 const double *__restrict__ cam, const double *__restrict__ poses, int64_t b, double thr2,
                 int32_t *__restrict__ out_count, uint8_t *__restrict__ out_inlier) {
    __shared__ double spose[kPnpPoses][12];
    __shared__ int scount[kPnpPoses];
    const int64_t b0 = (int64_t)blockIdx.y * kPnpPoses;
    const int nb = (int)((b - b0) < kPnpPoses ? (b - b0) : kPnpPoses);
    for (int k = threadIdx.x; k < nb * 12; k += kPnpThreads) spose[k / 12][k % 12] = poses[(b0 + k / 12) * 16 + k % 12];
    if (threadIdx.x < kPnpPoses) scount[threadIdx.x] = 0;
    const double fx = cam[0], sk = cam[1], cx = cam[2], fy = cam[4], cy = cam[5];
    const int64_t i0 = ((int64_t)blockIdx.x * kPnpThreads + threadIdx.x) * kPnpPts;
    double X[kPnpPts], Y[kPnpPts], Z[kPnpPts];
    float u[kPnpPts], v[kPnpPts];
#pragma unroll
    for (int k = 0; k < kPnpPts; ++k) {
        const int64_t i = i0 + k < n ? i0 + k : n - 1;
        X[k] = p3d[3 * i]; Y[k] = p3d[3 * i + 1]; Z[k] = p3d[3 * i + 2];
        u[k] = p2d[2 * i]; v[k] = p2d[2 * i + 1];
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        const double *P = spose[j];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < kPnpPts; ++k) {
            const double x = ((P[0] * X[k] + P[1] * Y[k]) + P[2] * Z[k]) + P[3];
            const double y = ((P[4] * X[k] + P[5] * Y[k]) + P[6] * Z[k]) + P[7];
            double z = ((P[8] * X[k] + P[9] * Y[k]) + P[10] * Z[k]) + P[11];
            z = z != 0.0 ? 1.0 / z : 1.0;
            const double xn = x * z, yn = y * z;
            const double pu = fx * xn + sk * yn + cx, pv = fy * yn + cy;
            const float du = u[k] - (float)pu, dv = v[k] - (float)pv;
            const float e = (float)((double)du * (double)du + (double)dv * (double)dv);
            const bool in = (i0 + k < n) && (double)e <= thr2;
            cnt += in ? 1 : 0;
            if (out_inlier != nullptr && i0 + k < n) out_inlier[(b0 + j) * n + i0 + k] = in ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0 && cnt != 0) atomicAdd(&scount[j], cnt);
    }
    __syncthreads();
    if (threadIdx.x < nb && scount[threadIdx.x] != 0) atomicAdd(&out_count[b0 + threadIdx.x], scount[threadIdx.x]);
}

}  // namespace isr

extern "C" {

int isr_pnp_score(const float *p3d, const float *p2d, int64_t n, const double *cam, const double *poses,
                  int64_t b, double reperr, int32_t *out_count, uint8_t *out_inlier, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0 && b >= 0, ISR_E_SHAPE, "pnp_score: negative size");
    if (b == 0) return ISR_OK;
    ISR_REQUIRE(out_count != nullptr && cam != nullptr && poses != nullptr, ISR_E_INVALID_ARG,
                "pnp_score: null pointer");
    ISR_REQUIRE(reperr >= 0.0, ISR_E_INVALID_ARG, "pnp_score: negative reprojection error");
    cudaStream_t st = (cudaStream_t)stream;
    ISR_TRY(check_cuda(cudaMemsetAsync(out_count, 0, (size_t)b * 4, st), "pnp_score memset"));
    if (n == 0) return ISR_OK;
    ISR_REQUIRE(p3d != nullptr && p2d != nullptr, ISR_E_INVALID_ARG, "pnp_score: null pointer");
    const int64_t gy = (b + kPnpPoses - 1) / kPnpPoses;
    ISR_REQUIRE(gy <= 65535, ISR_E_SHAPE, "pnp_score: more than %d hypotheses per call", 65535 * kPnpPoses);
    dim3 grid((unsigned)((n + kPnpThreads * kPnpPts - 1) / (kPnpThreads * kPnpPts)), (unsigned)gy);
    const double thr2 = reperr * reperr;
    ProfScope prof(kProfTransform, st);
    pnp_score_kernel<<<grid, kPnpThreads, 0, st>>>(p3d, p2d, n, cam, poses, b, thr2, out_count, out_inlier);
    return launched("pnp_score_kernel");
}

}  // extern "C"
