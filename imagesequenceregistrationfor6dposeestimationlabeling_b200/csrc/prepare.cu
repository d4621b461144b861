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

// Bounding spheres of the stored tiles, for the bound-pruned scan of nn2.cu: one sphere per
// 1024-point stage and one per 64-point sub-tile (ISR_SUB_TILE), each (cx, cy, cz, r) with the
// centre at the middle of the tile's bounding box and r >= max |p - c| over the tile's real
// points -- inflated so that the bound also holds for the FP64 (hi + lo) coordinates and
// survives the float32 evaluation here.  A tile without real points gets r = -1 and a far
// centre (never scanned first, always pruned).  grid (stages, batch), 256 threads; thread t
// owns points 4t..4t+3 of the stage, so a sub-tile is 16 consecutive lanes.
__global__ void __launch_bounds__(256)
tile_spheres_kernel(const float *__restrict__ soa7, int64_t n, int64_t npad, int64_t bstride,
                    float4 *__restrict__ out_stage, int stage_total, float4 *__restrict__ out_sub,
                    uint32_t *__restrict__ out_box) {
    constexpr int kSubs = ISR_SOA_TILE / ISR_SUB_TILE;  // 16
    __shared__ float red[6][8];
    __shared__ float rmax[8];
    const int s = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const float *base = soa7 + (int64_t)b * bstride + (int64_t)s * ISR_SOA_TILE + 4 * t;
    const float4 X = *reinterpret_cast<const float4 *>(base);
    const float4 Y = *reinterpret_cast<const float4 *>(base + npad);
    const float4 Z = *reinterpret_cast<const float4 *>(base + 2 * npad);
    const float px[4] = {X.x, X.y, X.z, X.w}, py[4] = {Y.x, Y.y, Y.z, Y.w}, pz[4] = {Z.x, Z.y, Z.z, Z.w};
    const int64_t g0 = (int64_t)s * ISR_SOA_TILE + 4 * t;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (g0 + k < n) {
            lo[0] = fminf(lo[0], px[k]); hi[0] = fmaxf(hi[0], px[k]);
            lo[1] = fminf(lo[1], py[k]); hi[1] = fmaxf(hi[1], py[k]);
            lo[2] = fminf(lo[2], pz[k]); hi[2] = fmaxf(hi[2], pz[k]);
        }
    }
    // ---- sub-tile: 16 lanes --------------------------------------------------------------
    float slo[3], shi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        slo[d] = lo[d]; shi[d] = hi[d];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            slo[d] = fminf(slo[d], __shfl_xor_sync(0xffffffffu, slo[d], o));
            shi[d] = fmaxf(shi[d], __shfl_xor_sync(0xffffffffu, shi[d], o));
        }
    }
    auto radius2 = [&](float cx, float cy, float cz) {
        float m = -1.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (g0 + k < n) {
                const float dx = px[k] - cx, dy = py[k] - cy, dz = pz[k] - cz;
                m = fmaxf(m, dx * dx + dy * dy + dz * dz);
            }
        }
        return m;
    };
    // r = sqrt(max d2) rounded up, + the lo parts of the points (|lo| <= 2^-24 |hi| per
    // component) and of the evaluation, relative to the tile's distance from the origin
    auto inflate = [](float m2, float cx, float cy, float cz) {
        const float r = __fsqrt_ru(m2) * 1.00002f;
        const float cn = __fsqrt_ru(cx * cx + cy * cy + cz * cz);
        return r + 1e-6f * (cn + r) + 1e-37f;
    };
    {
        const bool any = slo[0] <= shi[0];
        const float cx = 0.5f * (slo[0] + shi[0]), cy = 0.5f * (slo[1] + shi[1]), cz = 0.5f * (slo[2] + shi[2]);
        float m = any ? radius2(cx, cy, cz) : -1.f;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((t & 15) == 0) {
            const float r = any ? inflate(m, cx, cy, cz) : -1.f;
            out_sub[((int64_t)b * gridDim.x + s) * kSubs + (t >> 4)] =
                any ? make_float4(cx, cy, cz, r)
                    : make_float4(ISR_PAD_COORD, ISR_PAD_COORD, ISR_PAD_COORD, -1.f);
            if (out_box != nullptr) {
                // half-extents of the bounding box about the same centre, inflated like the radius,
                // as 10-bit fractions of it, rounded up (decoded as r * k / 1023, again rounded up)
                unsigned packed = 0x3FFFFFFFu;
                if (any && r > 0.f) {
                    const float slack = r - __fsqrt_rd(fmaxf(m, 0.f));  // what `inflate` added
                    packed = 0;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const float h = 0.5f * (shi[d] - slo[d]) * 1.00002f + slack;
                        int k = (int)ceilf(h / r * 1023.f * 1.00001f);
                        k = k < 1 ? 1 : (k > 1023 ? 1023 : k);
                        packed |= (unsigned)k << (10 * d);
                    }
                }
                out_box[((int64_t)b * gridDim.x + s) * kSubs + (t >> 4)] = packed;
            }
        }
    }
    // ---- stage: whole CTA ----------------------------------------------------------------
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        slo[d] = fminf(slo[d], __shfl_xor_sync(0xffffffffu, slo[d], 16));
        shi[d] = fmaxf(shi[d], __shfl_xor_sync(0xffffffffu, shi[d], 16));
    }
    if ((t & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) { red[d][t >> 5] = slo[d]; red[3 + d][t >> 5] = shi[d]; }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        slo[d] = red[d][0]; shi[d] = red[3 + d][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) { slo[d] = fminf(slo[d], red[d][w]); shi[d] = fmaxf(shi[d], red[3 + d][w]); }
    }
    const bool any = slo[0] <= shi[0];
    const float cx = 0.5f * (slo[0] + shi[0]), cy = 0.5f * (slo[1] + shi[1]), cz = 0.5f * (slo[2] + shi[2]);
    float m = any ? radius2(cx, cy, cz) : -1.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((t & 31) == 0) rmax[t >> 5] = m;
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, rmax[w]);
        out_stage[(int64_t)b * stage_total + s] =
            any ? make_float4(cx, cy, cz, inflate(m, cx, cy, cz))
                : make_float4(ISR_PAD_COORD, ISR_PAD_COORD, ISR_PAD_COORD, -1.f);
    }
}

// one warp per (chunk of 32 stages, batch item): the sphere that bounds the chunk's stage
// spheres -- centre of the box around them, radius max(|c_i - c| + r_i), rounded up
__global__ void __launch_bounds__(32)
chunk_spheres_kernel(float4 *__restrict__ stage, int stages, int total) {
    const int c = blockIdx.x, b = blockIdx.y, lane = threadIdx.x;
    float4 *st = stage + (int64_t)b * total;
    const int s = c * 32 + lane;
    float4 S = make_float4(0.f, 0.f, 0.f, -1.f);
    if (s < stages) S = st[s];
    const bool ok = S.w >= 0.f;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {ok ? S.x - S.w : inf, ok ? S.y - S.w : inf, ok ? S.z - S.w : inf};
    float hi[3] = {ok ? S.x + S.w : -inf, ok ? S.y + S.w : -inf, ok ? S.z + S.w : -inf};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    }
    const bool any = lo[0] <= hi[0];
    const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
    const float dx = S.x - cx, dy = S.y - cy, dz = S.z - cz;
    float r = ok ? __fsqrt_ru(dx * dx + dy * dy + dz * dz) * 1.00001f + S.w : -1.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (lane == 0)
        st[stages + c] = any ? make_float4(cx, cy, cz, r * 1.00001f + 1e-37f)
                             : make_float4(ISR_PAD_COORD, ISR_PAD_COORD, ISR_PAD_COORD, -1.f);
}

}  // namespace isr

extern "C" {

int64_t isr_stage_sphere_count(int64_t npad) {
    const int64_t stages = npad / ISR_SOA_TILE;
    return stages + (stages + 31) / 32;
}

int isr_centroid(const float *pts, int64_t n, double *out3, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0 && out3 != nullptr && (pts != nullptr || n == 0), ISR_E_INVALID_ARG,
                "centroid: bad argument");
    ProfScope prof(kProfTransform, (cudaStream_t)stream);
    centroid_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pts, n, out3);
    return launched("centroid_kernel");
}

int isr_tile_spheres(const float *soa7, int64_t n, int64_t npad, int64_t bstride, int64_t batch,
                     float *out_stage, float *out_sub, uint32_t *out_box, void *stream) {
    using namespace isr;
    ISR_REQUIRE(soa7 && out_stage && out_sub && n >= 0 && npad >= n && npad % ISR_SOA_TILE == 0 &&
                    npad > 0 && batch >= 1,
                ISR_E_INVALID_ARG, "tile_spheres: bad argument");
    ISR_REQUIRE(batch <= 65535, ISR_E_SHAPE, "tile_spheres: batch > 65535");
    ISR_REQUIRE(aligned16(out_stage) && aligned16(out_sub) && aligned16(soa7) && bstride % 4 == 0,
                ISR_E_ALIGN, "tile_spheres: pointers must be 16-byte aligned");
    dim3 grid((unsigned)(npad / ISR_SOA_TILE), (unsigned)batch);
    ProfScope prof(kProfTransform, (cudaStream_t)stream);
    const int stages = (int)(npad / ISR_SOA_TILE);
    const int total = (int)isr_stage_sphere_count(npad);
    tile_spheres_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        soa7, n, npad, bstride, reinterpret_cast<float4 *>(out_stage), total, reinterpret_cast<float4 *>(out_sub),
        out_box);
    ISR_TRY(launched("tile_spheres_kernel"));
    dim3 cgrid((unsigned)(total - stages), (unsigned)batch);
    chunk_spheres_kernel<<<cgrid, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(out_stage), stages,
                                                                total);
    return launched("chunk_spheres_kernel");
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
