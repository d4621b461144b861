// K1 -- batched rigid transform (HBM-bound).
//
// Replaces the numpy `pc.dot(R.T) + t` of verfication.py:83-85 / icp.py:68 /
// choosePose.py:21-22 and Open3D's PointCloud.transform (icp.py:22,110).  The
// arithmetic is FP64 (the reference's is) with one rounding to float32 on store.
//
// Two output layouts:
//   AoS  [b][n][3]      the public transform_points result; staged through shared memory
//                       so the global stores are 16-byte vectors (coalesced 128-B lines);
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

// grid: (ceil(n / 256), ceil(b / kTfPosesPerCta)): the CTA's 256 points stay in registers
// while it walks 16 poses; every pose's 3 KB go out as float4 through shared memory.
constexpr int kTfPosesPerCta = 16;

__global__ void __launch_bounds__(kTfThreads)
transform_aos_kernel(const float *__restrict__ pts, int64_t n, const double *__restrict__ poses,
                     int64_t b, float *__restrict__ out, int vec_ok) {
    __shared__ float stage[kTfThreads * 3];
    const int64_t p0 = (int64_t)blockIdx.x * kTfThreads;
    const int cnt = (int)min((int64_t)kTfThreads, n - p0);
    const int tid = threadIdx.x;

    // coalesced read of this block's 256 x 3 floats through shared memory
    const float *src = pts + p0 * 3;
    for (int k = tid; k < cnt * 3; k += kTfThreads) stage[k] = src[k];
    __syncthreads();
    float x = 0.f, y = 0.f, z = 0.f;
    if (tid < cnt) { x = stage[tid * 3]; y = stage[tid * 3 + 1]; z = stage[tid * 3 + 2]; }
    const int64_t b_begin = (int64_t)blockIdx.y * kTfPosesPerCta;
    const int64_t b_end = min(b, b_begin + kTfPosesPerCta);
    for (int64_t bb = b_begin; bb < b_end; ++bb) {
        const Pose12 P = load_pose(poses + bb * 16);
        float ox = 0.f, oy = 0.f, oz = 0.f;
        if (tid < cnt) apply_pose(P, x, y, z, ox, oy, oz);
        __syncthreads();  // previous pose's stores have been read out of `stage`
        if (tid < cnt) {
            stage[tid * 3] = ox;
            stage[tid * 3 + 1] = oy;
            stage[tid * 3 + 2] = oz;
        }
        __syncthreads();
        float *dst = out + (bb * n + p0) * 3;
        if (vec_ok && cnt == kTfThreads) {
            // 768 floats = 192 float4, 16-byte aligned because n % 4 == 0 and p0 % 256 == 0
            if (tid < kTfThreads * 3 / 4)
                reinterpret_cast<float4 *>(dst)[tid] = reinterpret_cast<const float4 *>(stage)[tid];
        } else {
            for (int k = tid; k < cnt * 3; k += kTfThreads) dst[k] = stage[k];
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

}  // namespace isr

extern "C" {

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
    const int vec_ok = (n % 4 == 0) && aligned16(out);
    dim3 grid((unsigned)((n + kTfThreads - 1) / kTfThreads),
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
