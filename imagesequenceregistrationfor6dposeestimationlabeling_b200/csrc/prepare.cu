// K1' -- cloud preparation for the filtered nearest-neighbour kernel (nn2.cu).
//
// Same rigid transform as K1 (FP64 R.p + t; verfication.py:83-85, icp.py:68,110), but the
// result is (a) shifted by a per-batch centre c so that coordinates are small, and (b) kept
// at FP64 accuracy as a float32 hi/lo pair.  Output "SoA7" planes per cloud, each npad long:
//   0,1,2  hi.x hi.y hi.z    float32( R.p + t - c )
//   3      norm              fl32 |hi|^2 = fma(z,z, fma(y,y, x*x))
//   4,5,6  lo.x lo.y lo.z    float32( (R.p + t - c) - hi )
// Distances are translation invariant, so subtracting the same c from both clouds of a pair
// changes nothing mathematically; hi+lo reproduces the FP64 coordinate to ~2^-48.
// Padded slots: hi = ISR_PAD_COORD, norm = 3 ISR_PAD_COORD^2 (finite), lo = 0.
//
// The centre is c = C . m with m the centroid of a cloud (FP64, centroid_kernel) and C the
// "centre pose" of the batch item (NULL = identity): for candidate verification both clouds
// of candidate k use c_k = Pt_k . centroid(cloud_t).
// HBM: reads n*12 B (L2-resident across the batch), writes 7 planes * 4 B per point.
#include "isr_common.cuh"

namespace isr {

constexpr int kPrepThreads = 256;

// single CTA, fixed order => deterministic.  out[0..2] = mean of pts (FP64).
__global__ void __launch_bounds__(1024)
centroid_kernel(const float *__restrict__ pts, int64_t n, double *__restrict__ out) {
    __shared__ double red[3][32];
    double sx = 0, sy = 0, sz = 0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        sx += pts[3 * i];
        sy += pts[3 * i + 1];
        sz += pts[3 * i + 2];
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = sx; red[1][threadIdx.x >> 5] = sy; red[2][threadIdx.x >> 5] = sz;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        sx = warp_sum(red[0][threadIdx.x]);
        sy = warp_sum(red[1][threadIdx.x]);
        sz = warp_sum(red[2][threadIdx.x]);
        if (threadIdx.x == 0) {
            const double inv = n > 0 ? 1.0 / (double)n : 0.0;
            out[0] = sx * inv; out[1] = sy * inv; out[2] = sz * inv;
        }
    }
}

// grid: (npad / 256, ceil(b / kPosesPerCta)).  A CTA keeps its 256 points in registers and
// walks kPosesPerCta poses, so the (gathered) point read and the CTA start-up are paid once
// per 16 x 7 KB of coalesced plane writes.
constexpr int kPosesPerCta = 16;

__global__ void __launch_bounds__(kPrepThreads)
prepare_soa7_kernel(const float *__restrict__ pts, const float *__restrict__ pts_lo,
                    const int32_t *__restrict__ perm, int64_t n, const double *__restrict__ poses,
                    int64_t pose_stride, const double *__restrict__ centre_poses,
                    int64_t centre_pose_stride, const double *__restrict__ centroid, int64_t b,
                    float *__restrict__ out, int64_t npad, const int32_t *__restrict__ skip,
                    int64_t skip_stride) {
    __shared__ float stage[kPrepThreads * 3];
    __shared__ double spose[kPosesPerCta][12];    // pose rows of this CTA's batch items
    __shared__ double scentre[kPosesPerCta][3];   // and their centres c_b
    const int64_t p0 = (int64_t)blockIdx.x * kPrepThreads;
    const int cnt = (int)max((int64_t)0, min((int64_t)kPrepThreads, n - p0));
    const int tid = threadIdx.x;
    {
        const int64_t b0 = (int64_t)blockIdx.y * kPosesPerCta;
        const int nb = (int)min((int64_t)kPosesPerCta, b - b0);
        if (poses != nullptr && tid < nb * 12)
            spose[tid / 12][tid % 12] = poses[(b0 + tid / 12) * pose_stride + tid % 12];
        if (tid < nb) {
            double cx = 0, cy = 0, cz = 0;
            if (centroid != nullptr) {
                cx = centroid[0]; cy = centroid[1]; cz = centroid[2];
                if (centre_poses != nullptr) {
                    const double *C = centre_poses + (b0 + tid) * centre_pose_stride;
                    const double mx = cx, my = cy, mz = cz;
                    cx = ((C[0] * mx + C[1] * my) + C[2] * mz) + C[3];
                    cy = ((C[4] * mx + C[5] * my) + C[6] * mz) + C[7];
                    cz = ((C[8] * mx + C[9] * my) + C[10] * mz) + C[11];
                }
            }
            scentre[tid][0] = cx; scentre[tid][1] = cy; scentre[tid][2] = cz;
        }
    }
    if (perm == nullptr) {
        const float *src = pts + p0 * 3;
        for (int k = tid; k < cnt * 3; k += kPrepThreads) stage[k] = src[k];
    } else if (tid < cnt) {  // stored position p0+tid holds original point perm[p0+tid]
        const float *src = pts + (int64_t)perm[p0 + tid] * 3;
        stage[tid * 3] = src[0]; stage[tid * 3 + 1] = src[1]; stage[tid * 3 + 2] = src[2];
    }
    __syncthreads();
    const bool live = tid < cnt;
    double px = 0, py = 0, pz = 0;
    if (live) {
        px = stage[tid * 3]; py = stage[tid * 3 + 1]; pz = stage[tid * 3 + 2];
        if (pts_lo != nullptr) {  // float64 input carried as a float32 hi/lo pair
            const float *l = pts_lo + (perm != nullptr ? (int64_t)perm[p0 + tid] : p0 + tid) * 3;
            px += (double)l[0]; py += (double)l[1]; pz += (double)l[2];
        }
    }
    const int64_t b_begin = (int64_t)blockIdx.y * kPosesPerCta;
    const int64_t b_end = min(b, b_begin + kPosesPerCta);
    for (int64_t bb = b_begin; bb < b_end; ++bb) {
        if (skip != nullptr && skip[bb * skip_stride] != 0) continue;
        float hx = ISR_PAD_COORD, hy = ISR_PAD_COORD, hz = ISR_PAD_COORD;
        float lx = 0.f, ly = 0.f, lz = 0.f;
        if (live) {
            const int bl = (int)(bb - b_begin);
            const double cx = scentre[bl][0], cy = scentre[bl][1], cz = scentre[bl][2];
            double x = px, y = py, z = pz;
            if (poses != nullptr) {
                const double *P = spose[bl];
                x = ((P[0] * px + P[1] * py) + P[2] * pz) + P[3];
                y = ((P[4] * px + P[5] * py) + P[6] * pz) + P[7];
                z = ((P[8] * px + P[9] * py) + P[10] * pz) + P[11];
            }
            x -= cx; y -= cy; z -= cz;
            hx = (float)x; hy = (float)y; hz = (float)z;
            lx = (float)(x - (double)hx); ly = (float)(y - (double)hy); lz = (float)(z - (double)hz);
        }
        const float nrm = __fmaf_rn(hz, hz, __fmaf_rn(hy, hy, __fmul_rn(hx, hx)));
        float *o = out + bb * 7 * npad + p0 + tid;
        o[0] = hx;
        o[npad] = hy;
        o[2 * npad] = hz;
        o[3 * npad] = nrm;
        o[4 * npad] = lx;
        o[5 * npad] = ly;
        o[6 * npad] = lz;
    }
}

// centroid of the real points of every 1024-point SoA tile ("stage"): grid (stages, batch)
__global__ void __launch_bounds__(256)
stage_centroid_kernel(const float *__restrict__ soa7, int64_t n, int64_t npad, int64_t bstride,
                      float4 *__restrict__ out) {
    __shared__ float red[4][8];
    const int s = blockIdx.x, b = blockIdx.y;
    const float *base = soa7 + (int64_t)b * bstride + (int64_t)s * ISR_SOA_TILE;
    float cx = 0.f, cy = 0.f, cz = 0.f, cn = 0.f;
    for (int k = threadIdx.x; k < ISR_SOA_TILE; k += 256) {
        if ((int64_t)s * ISR_SOA_TILE + k < n) {
            cx += base[k]; cy += base[npad + k]; cz += base[2 * npad + k]; cn += 1.f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx += __shfl_xor_sync(0xffffffffu, cx, o);
        cy += __shfl_xor_sync(0xffffffffu, cy, o);
        cz += __shfl_xor_sync(0xffffffffu, cz, o);
        cn += __shfl_xor_sync(0xffffffffu, cn, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = cx; red[1][threadIdx.x >> 5] = cy;
        red[2][threadIdx.x >> 5] = cz; red[3][threadIdx.x >> 5] = cn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        cx = cy = cz = cn = 0.f;
        for (int w = 0; w < 8; ++w) { cx += red[0][w]; cy += red[1][w]; cz += red[2][w]; cn += red[3][w]; }
        const float inv = cn > 0.f ? 1.f / cn : 0.f;
        // an all-padding tile gets a far-away centroid so it is never chosen first
        out[(int64_t)b * gridDim.x + s] = cn > 0.f ? make_float4(cx * inv, cy * inv, cz * inv, cn)
                                                   : make_float4(1e18f, 1e18f, 1e18f, 0.f);
    }
}

}  // namespace isr

extern "C" {

int isr_centroid(const float *pts, int64_t n, double *out3, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0 && out3 != nullptr && (pts != nullptr || n == 0), ISR_E_INVALID_ARG,
                "centroid: bad argument");
    ProfScope prof(kProfTransform, (cudaStream_t)stream);
    centroid_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pts, n, out3);
    return launched("centroid_kernel");
}

int isr_stage_centroids(const float *soa7, int64_t n, int64_t npad, int64_t bstride, int64_t batch,
                        float *out, void *stream) {
    using namespace isr;
    ISR_REQUIRE(soa7 && out && n >= 0 && npad >= n && npad % ISR_SOA_TILE == 0 && npad > 0 && batch >= 1,
                ISR_E_INVALID_ARG, "stage_centroids: bad argument");
    ISR_REQUIRE(batch <= 65535, ISR_E_SHAPE, "stage_centroids: batch > 65535");
    ISR_REQUIRE(aligned16(out), ISR_E_ALIGN, "stage_centroids: out not 16-byte aligned");
    dim3 grid((unsigned)(npad / ISR_SOA_TILE), (unsigned)batch);
    ProfScope prof(kProfTransform, (cudaStream_t)stream);
    stage_centroid_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(soa7, n, npad, bstride,
                                                                   reinterpret_cast<float4 *>(out));
    return launched("stage_centroid_kernel");
}

int isr_prepare_cloud(const float *pts, const float *pts_lo, const int32_t *perm, int64_t n,
                      const double *poses, int64_t pose_stride,
                      const double *centre_poses, int64_t centre_pose_stride,
                      const double *centroid, int64_t b, float *out_soa7, int64_t npad,
                      const int32_t *skip, int64_t skip_stride, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0 && b >= 1, ISR_E_SHAPE, "prepare_cloud: bad size");
    ISR_REQUIRE(out_soa7 && (pts || n == 0), ISR_E_INVALID_ARG, "prepare_cloud: null pointer");
    ISR_REQUIRE(npad >= n && npad % ISR_SOA_TILE == 0 && npad > 0, ISR_E_SHAPE,
                "prepare_cloud: npad %lld must be a positive multiple of %d and >= n",
                (long long)npad, ISR_SOA_TILE);
    ISR_REQUIRE(poses != nullptr || b == 1, ISR_E_INVALID_ARG,
                "prepare_cloud: repack (poses NULL) needs b == 1");
    ISR_REQUIRE(b <= 65535, ISR_E_SHAPE, "prepare_cloud: batch %lld > 65535", (long long)b);
    ISR_REQUIRE(aligned16(out_soa7), ISR_E_ALIGN, "prepare_cloud: out not 16-byte aligned");
    dim3 grid((unsigned)(npad / kPrepThreads), (unsigned)((b + kPosesPerCta - 1) / kPosesPerCta));
    ProfScope prof(kProfTransform, (cudaStream_t)stream);
    prepare_soa7_kernel<<<grid, kPrepThreads, 0, (cudaStream_t)stream>>>(
        pts, pts_lo, perm, n, poses, pose_stride, centre_poses, centre_pose_stride, centroid, b,
        out_soa7, npad, skip, skip_stride);
    return launched("prepare_soa7_kernel");
}

}  // extern "C"
