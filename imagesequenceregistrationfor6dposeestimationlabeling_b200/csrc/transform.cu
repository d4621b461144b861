// K1 -- batched rigid transform (HBM-bound).
//
// Replaces the numpy `pc.dot(R.T) + t` of verfication.py:83-85 / icp.py:68 /
// choosePose.py:21-22 and Open3D's PointCloud.transform (icp.py:22,110).  The
// arithmetic is FP64 (the reference's is) with one rounding to float32 on store.
//
// Two output layouts:
//   AoS  [b][n][3]      the public transform_points result; a thread owns 4 consecutive points
//                       so loads and stores are 16-byte vectors (a warp moves 1536 contiguous B);
//   SoA  [b][3][npad]   the planes K2 streams with bulk async copies; padded slots get
//                       ISR_PAD_COORD so they can never win a minimum.
// Algorithmic bytes per launch: n*12 read (L2-resident across the batch) + b*n*12 written.
#include "isr_common.cuh"

namespace isr {

constexpr int kTfThreads = 256;

struct Pose12 {
    double r[9];
    double t[3];
};

__device__ __forceinline__ Pose12 load_pose(const double *p) {
    Pose12 P;
    P.r[0] = p[0]; P.r[1] = p[1]; P.r[2] = p[2];  P.t[0] = p[3];
    P.r[3] = p[4]; P.r[4] = p[5]; P.r[5] = p[6];  P.t[1] = p[7];
    P.r[6] = p[8]; P.r[7] = p[9]; P.r[8] = p[10]; P.t[2] = p[11];
    return P;
}

__device__ __forceinline__ void apply_pose(const Pose12 &P, float x, float y, float z, float &ox,
                                           float &oy, float &oz) {
    const double dx = x, dy = y, dz = z;
    // same evaluation order as numpy's row . column product followed by "+ t"
    ox = (float)(((P.r[0] * dx + P.r[1] * dy) + P.r[2] * dz) + P.t[0]);
    oy = (float)(((P.r[3] * dx + P.r[4] * dy) + P.r[5] * dz) + P.t[1]);
    oz = (float)(((P.r[6] * dx + P.r[7] * dy) + P.r[8] * dz) + P.t[2]);
}

// grid: (ceil(n / 1024), ceil(b / kTfPosesPerCta)).  A thread owns 4 consecutive points
// (three float4 in, three float4 out per pose -- a warp reads/writes 1536 contiguous bytes),
// keeps them in registers and walks 16 poses that were staged in shared memory once.
// No barrier in the pose loop.  n % 4 != 0 or unaligned pointers take the scalar tail path.
constexpr int kTfPosesPerCta = 16;
constexpr int kTfPointsPerThread = 4;

__global__ void __launch_bounds__(kTfThreads)
transform_aos_kernel(const float *__restrict__ pts, int64_t n, const double *__restrict__ poses,
                     int64_t b, float *__restrict__ out, int vec_ok) {
    __shared__ double spose[kTfPosesPerCta][12];  // this CTA's poses, fetched once
    // per-warp transposition buffer: lanes own 12 consecutive floats each, but a store
    // instruction should cover 512 contiguous bytes (lane l writes float4 number k*32+l)
    __shared__ __align__(16) float wbuf[kTfThreads / 32][32 * 12];
    const int tid = threadIdx.x;
    const int64_t b_begin = (int64_t)blockIdx.y * kTfPosesPerCta;
    const int64_t b_end = min(b, b_begin + kTfPosesPerCta);
    {
        const int nb = (int)(b_end - b_begin);
        if (tid < nb * 12) spose[tid / 12][tid % 12] = poses[(b_begin + tid / 12) * 16 + tid % 12];
    }
    __syncthreads();
    const int64_t i0 = ((int64_t)blockIdx.x * kTfThreads + tid) * kTfPointsPerThread;
    if (i0 >= n) return;
    const int cnt = (int)min((int64_t)kTfPointsPerThread, n - i0);
    // every lane of this warp holds 4 live points (warp-uniform: i0 grows with the lane)
    const bool warp_full = __all_sync(__activemask(), cnt == kTfPointsPerThread) &&
                           __activemask() == 0xffffffffu;
    float in[12];
    if (vec_ok && cnt == kTfPointsPerThread) {
        const float4 *s4 = reinterpret_cast<const float4 *>(pts + i0 * 3);
        const float4 a = s4[0], c = s4[1], d = s4[2];
        in[0] = a.x; in[1] = a.y; in[2] = a.z; in[3] = a.w; in[4] = c.x; in[5] = c.y;
        in[6] = c.z; in[7] = c.w; in[8] = d.x; in[9] = d.y; in[10] = d.z; in[11] = d.w;
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) in[k] = (k < cnt * 3) ? pts[i0 * 3 + k] : 0.f;
    }
    for (int64_t bb = b_begin; bb < b_end; ++bb) {
        const Pose12 P = load_pose(spose[bb - b_begin]);
        float o[12];
#pragma unroll
        for (int k = 0; k < kTfPointsPerThread; ++k)
            apply_pose(P, in[3 * k], in[3 * k + 1], in[3 * k + 2], o[3 * k], o[3 * k + 1], o[3 * k + 2]);
        float *dst = out + (bb * n + i0) * 3;
        if (vec_ok && warp_full) {
            float4 *w4 = reinterpret_cast<float4 *>(wbuf[tid >> 5]);
            const int lane = tid & 31;
            w4[lane * 3 + 0] = make_float4(o[0], o[1], o[2], o[3]);
            w4[lane * 3 + 1] = make_float4(o[4], o[5], o[6], o[7]);
            w4[lane * 3 + 2] = make_float4(o[8], o[9], o[10], o[11]);
            __syncwarp();
            float4 *d4 = reinterpret_cast<float4 *>(dst - (int64_t)lane * 12);  // the warp's block
            d4[lane] = w4[lane];
            d4[32 + lane] = w4[32 + lane];
            d4[64 + lane] = w4[64 + lane];
            __syncwarp();
        } else if (vec_ok && cnt == kTfPointsPerThread) {
            float4 *d4 = reinterpret_cast<float4 *>(dst);
            d4[0] = make_float4(o[0], o[1], o[2], o[3]);
            d4[1] = make_float4(o[4], o[5], o[6], o[7]);
            d4[2] = make_float4(o[8], o[9], o[10], o[11]);
        } else {
            for (int k = 0; k < cnt * 3; ++k) dst[k] = o[k];
        }
    }
}

// grid: (npad / 256, b).  poses == nullptr -> plain repack.
__global__ void __launch_bounds__(kTfThreads)
transform_soa_kernel(const float *__restrict__ pts, int64_t n, const double *__restrict__ poses,
                     int64_t pose_stride, float *__restrict__ out, int64_t npad,
                     const int32_t *__restrict__ skip, int64_t skip_stride) {
    __shared__ float stage[kTfThreads * 3];
    const int b = blockIdx.y;
    if (skip != nullptr && skip[(int64_t)b * skip_stride] != 0) return;
    const int64_t p0 = (int64_t)blockIdx.x * kTfThreads;
    const int cnt = (int)max((int64_t)0, min((int64_t)kTfThreads, n - p0));
    const int tid = threadIdx.x;
    const float *src = pts + p0 * 3;
    for (int k = tid; k < cnt * 3; k += kTfThreads) stage[k] = src[k];
    __syncthreads();
    float ox = ISR_PAD_COORD, oy = ISR_PAD_COORD, oz = ISR_PAD_COORD;
    if (tid < cnt) {
        const float x = stage[tid * 3], y = stage[tid * 3 + 1], z = stage[tid * 3 + 2];
        if (poses != nullptr) {
            const Pose12 P = load_pose(poses + (int64_t)b * pose_stride);
            apply_pose(P, x, y, z, ox, oy, oz);
        } else {
            ox = x; oy = y; oz = z;
        }
    }
    float *o = out + (int64_t)b * 3 * npad + p0 + tid;
    o[0] = ox;
    o[npad] = oy;
    o[2 * npad] = oz;
}

// float64 in / float64 out, one pose, in place allowed: Open3D's PointCloud.transform keeps
// double coordinates (icp.py:22,110).  24 B read + 24 B written per point.
__global__ void __launch_bounds__(kTfThreads)
transform_f64_kernel(const double *__restrict__ pts, int64_t n, const double *__restrict__ pose,
                     double *__restrict__ out) {
    const Pose12 P = load_pose(pose);
    for (int64_t i = (int64_t)blockIdx.x * kTfThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kTfThreads) {
        const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        // same evaluation order as numpy's row . column product followed by "+ t"
        const double ox = ((P.r[0] * x + P.r[1] * y) + P.r[2] * z) + P.t[0];
        const double oy = ((P.r[3] * x + P.r[4] * y) + P.r[5] * z) + P.t[1];
        const double oz = ((P.r[6] * x + P.r[7] * y) + P.r[8] * z) + P.t[2];
        out[3 * i] = ox; out[3 * i + 1] = oy; out[3 * i + 2] = oz;
    }
}

}  // namespace isr

extern "C" {

int isr_transform_points_f64(const double *pts, int64_t n, const double *pose, double *out, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0, ISR_E_SHAPE, "transform_points_f64: negative size");
    if (n == 0) return ISR_OK;
    ISR_REQUIRE(pts && pose && out, ISR_E_INVALID_ARG, "transform_points_f64: null pointer");
    int64_t blocks = (n + kTfThreads - 1) / kTfThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ProfScope prof(kProfTransform, (cudaStream_t)stream);
    transform_f64_kernel<<<(unsigned)blocks, kTfThreads, 0, (cudaStream_t)stream>>>(pts, n, pose, out);
    return launched("transform_f64_kernel");
}

int64_t isr_soa_padded_len(int64_t n) {
    return n <= 0 ? ISR_SOA_TILE : isr::round_up(n, ISR_SOA_TILE);
}

int isr_transform_points(const float *pts, int64_t n, const double *poses, int64_t b, float *out,
                         void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0 && b >= 0, ISR_E_SHAPE, "transform_points: negative size");
    if (n == 0 || b == 0) return ISR_OK;
    ISR_REQUIRE(pts && poses && out, ISR_E_INVALID_ARG, "transform_points: null pointer");
    ISR_REQUIRE(b <= 65535 * 16, ISR_E_SHAPE, "transform_points: batch %lld too large", (long long)b);
    const int vec_ok = (n % 4 == 0) && aligned16(out) && aligned16(pts);
    const int64_t per_cta = (int64_t)kTfThreads * kTfPointsPerThread;
    dim3 grid((unsigned)((n + per_cta - 1) / per_cta),
              (unsigned)((b + kTfPosesPerCta - 1) / kTfPosesPerCta));
    ProfScope prof(kProfTransform, (cudaStream_t)stream);
    transform_aos_kernel<<<grid, kTfThreads, 0, (cudaStream_t)stream>>>(pts, n, poses, b, out, vec_ok);
    return launched("transform_aos_kernel");
}

int isr_transform_points_soa(const float *pts, int64_t n, const double *poses, int64_t pose_stride,
                             int64_t b, float *out_soa, int64_t npad, const int32_t *skip,
                             int64_t skip_stride, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0 && b >= 1, ISR_E_SHAPE, "transform_points_soa: bad size");
    ISR_REQUIRE(out_soa && (pts || n == 0), ISR_E_INVALID_ARG, "transform_points_soa: null pointer");
    ISR_REQUIRE(npad >= n && npad % ISR_SOA_TILE == 0 && npad > 0, ISR_E_SHAPE,
                "transform_points_soa: npad %lld must be a positive multiple of %d and >= n",
                (long long)npad, ISR_SOA_TILE);
    ISR_REQUIRE(poses != nullptr || b == 1, ISR_E_INVALID_ARG,
                "transform_points_soa: repack (poses NULL) needs b == 1");
    ISR_REQUIRE(b <= 65535, ISR_E_SHAPE, "transform_points_soa: batch %lld > 65535", (long long)b);
    ISR_REQUIRE(aligned16(out_soa), ISR_E_ALIGN, "transform_points_soa: out not 16-byte aligned");
    dim3 grid((unsigned)(npad / kTfThreads), (unsigned)b);
    ProfScope prof(kProfTransform, (cudaStream_t)stream);
    transform_soa_kernel<<<grid, kTfThreads, 0, (cudaStream_t)stream>>>(
        pts, n, poses, pose_stride, out_soa, npad, skip, skip_stride);
    return launched("transform_soa_kernel");
}

}  // extern "C"
