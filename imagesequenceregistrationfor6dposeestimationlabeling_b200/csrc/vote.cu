// The callers either side of the batched ADD-S scoring in choosePose.py,
// on the device so that the 1.64 M pose pairs of a 1280-image sequence never visit the host:
//   rel_pose_table_kernel   relative_poses[i][j] = [R_i^T R_j | t_j - t_i]    choosePose.py:43-51,98-107
//   rigid_relative_kernel   M = Pt^-1 Pq: ADD-S of (Pq . verts) against (Pt . surface) equals the
//                           ADD-S of (M . verts) against the surface itself, so the surface cloud
//                           is prepared ONCE for the whole table (verify.cu: isr_adds_fixed_target)
//   vote_rows_kernel        error[i][j] = loss[i][j] < threshold; votes[i] = sum_j error[i][j]
//   first_max_kernel        np.argmax(votes): the first maximum                choosePose.py:135-151
// All tiny and HBM/L2-bound (128 B per pose); integer sums, FP64 pose algebra, no atomics.
#include <math_constants.h>

#include "isr_common.cuh"

namespace isr {

// out[k - pair0] for the flat pair index k = i * n + j, k in [pair0, pair0 + count)
__global__ void __launch_bounds__(128)
rel_pose_table_kernel(const double *__restrict__ R, const double *__restrict__ t, int64_t n, int64_t pair0,
                      int64_t count, double *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (idx >= count) return;
    const int64_t k = pair0 + idx, i = k / n, j = k % n;
    const double *Ri = R + 9 * i, *Rj = R + 9 * j;
    double *o = out + 16 * idx;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int c = 0; c < 3; ++c)  // (R_i^T R_j)[a][c] = sum_b R_i[b][a] R_j[b][c]
            o[4 * a + c] = __fma_rn(Ri[6 + a], Rj[6 + c], __fma_rn(Ri[3 + a], Rj[3 + c], __dmul_rn(Ri[a], Rj[c])));
        o[4 * a + 3] = __dsub_rn(t[3 * j + a], t[3 * i + a]);
    }
    o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
}

// M = [Rt | tt]^-1 [Rq | tq] with the inverse of a ROTATION: M.R = Rt^T Rq, M.t = Rt^T (tq - tt)
__global__ void __launch_bounds__(128)
rigid_relative_kernel(const double *__restrict__ Pq, const double *__restrict__ Pt, int64_t b,
                      double *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (idx >= b) return;
    const double *q = Pq + 16 * idx, *p = Pt + 16 * idx;
    double *o = out + 16 * idx;
    const double dx = q[3] - p[3], dy = q[7] - p[7], dz = q[11] - p[11];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            o[4 * a + c] = __fma_rn(p[8 + a], q[8 + c], __fma_rn(p[4 + a], q[4 + c], __dmul_rn(p[a], q[c])));
        o[4 * a + 3] = __fma_rn(p[8 + a], dz, __fma_rn(p[4 + a], dy, __dmul_rn(p[a], dx)));
    }
    o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
}

// one CTA per row of the loss table
__global__ void __launch_bounds__(256)
vote_rows_kernel(const double *__restrict__ loss, int64_t cols, double thr, uint8_t *__restrict__ error,
                 int32_t *__restrict__ votes) {
    __shared__ int red[8];
    const int64_t i = blockIdx.x;
    int cnt = 0;
    for (int64_t j = threadIdx.x; j < cols; j += 256) {
        const bool e = loss[i * cols + j] < thr;  // NaN / +inf (a failed pose): no vote
        if (error != nullptr) error[i * cols + j] = e ? 1 : 0;
        cnt += e ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += red[w];
        votes[i] = s;
    }
}

// single CTA: out[0] = index of the first maximum of v[0..n), out[1] = that maximum
__global__ void __launch_bounds__(1024)
first_max_kernel(const int32_t *__restrict__ v, int64_t n, int64_t *__restrict__ out) {
    __shared__ long long sk[32];
    // key: larger value first, then smaller index
    long long best = -1;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        const long long key = ((long long)v[i] << 32) | (long long)(0x7FFFFFFF - (int)i);
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) sk[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) best = sk[w] > best ? sk[w] : best;
        out[0] = n > 0 ? (long long)(0x7FFFFFFF - (int)(best & 0xFFFFFFFFll)) : -1;
        out[1] = n > 0 ? (best >> 32) : 0;
    }
}

}  // namespace isr

extern "C" {

int isr_rel_pose_table(const double *R, const double *t, int64_t n, int64_t pair0, int64_t count,
                       double *out, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 1 && pair0 >= 0 && count >= 0 && pair0 + count <= n * n, ISR_E_SHAPE,
                "rel_pose_table: pairs [%lld, %lld) outside the %lld x %lld table", (long long)pair0,
                (long long)(pair0 + count), (long long)n, (long long)n);
    if (count == 0) return ISR_OK;
    ISR_REQUIRE(R && t && out, ISR_E_INVALID_ARG, "rel_pose_table: null pointer");
    rel_pose_table_kernel<<<(unsigned)((count + 127) / 128), 128, 0, (cudaStream_t)stream>>>(R, t, n, pair0, count, out);
    return launched("rel_pose_table_kernel");
}

int isr_rigid_relative(const double *poses_q, const double *poses_t, int64_t b, double *out, void *stream) {
    using namespace isr;
    ISR_REQUIRE(b >= 0, ISR_E_SHAPE, "rigid_relative: b < 0");
    if (b == 0) return ISR_OK;
    ISR_REQUIRE(poses_q && poses_t && out, ISR_E_INVALID_ARG, "rigid_relative: null pointer");
    rigid_relative_kernel<<<(unsigned)((b + 127) / 128), 128, 0, (cudaStream_t)stream>>>(poses_q, poses_t, b, out);
    return launched("rigid_relative_kernel");
}

int isr_first_max(const int32_t *v, int64_t n, int64_t *out2, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 1 && n <= 0x7FFFFFFF && v && out2, ISR_E_INVALID_ARG, "first_max: bad argument");
    first_max_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(v, n, out2);
    return launched("first_max_kernel");
}

int isr_vote(const double *loss, int64_t rows, int64_t cols, double threshold, uint8_t *out_error,
             int32_t *out_votes, int64_t *out_best, void *stream) {
    using namespace isr;
    ISR_REQUIRE(rows >= 1 && cols >= 1 && rows <= 0x7FFFFFFF, ISR_E_SHAPE, "vote: bad table size");
    ISR_REQUIRE(loss && out_votes, ISR_E_INVALID_ARG, "vote: null pointer");
    vote_rows_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(loss, cols, threshold, out_error, out_votes);
    ISR_TRY(launched("vote_rows_kernel"));
    if (out_best != nullptr) {
        first_max_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(out_votes, rows, out_best);
        ISR_TRY(launched("first_max_kernel"));
    }
    return ISR_OK;
}

}  // extern "C"

namespace isr {

// ---- ADD-S bounds from the target's tile spheres alone (no point is touched) -------------------
// The vote of choosePose.py:135 only asks whether ADDS < 0.1 x diameter.  For a query q and a
// non-empty target tile with bounding sphere (c, r):   |q - c| - r  <=  d(q, tile)  <=  |q - c| + r,
// so with the spheres of a FIXED, prepared target (isr_tile_spheres: 1024-point stages and their
// 64-point sub-tiles)
//     U(q) = min over examined sub-tiles (|q - c| + r)              >= the 1-NN distance of q
//     L(q) = min( min over examined sub-tiles max(0, |q - c| - r),
//                 min over the stages NOT examined max(0, |q - c_s| - r_s) )   <= it
// and mean L <= ADDS <= mean U.  A pose pair whose mean U is below the threshold votes, one whose
// mean L reaches it does not, and only the pairs in between need the exact search: good
// predictions sit at 1-3 mm against a 12 mm threshold, failed ones at tens of mm.  One CTA per
// pose pair; the spheres live in shared memory; the stage spheres are tested once per WARP of 32
// consecutive (curve-ordered) vertices against the warp's bounding sphere, the 16 sub-tiles of the
// kExamine nearest stages once per query.  cloud_q should be in curve order (isr_spatial_order):
// any order gives valid bounds, a coherent one tight bounds.  FP32 with
// outward rounding margins; the means are accumulated in FP64 in a fixed order.
constexpr int kExamine = 3;
constexpr int kBoundThreads = 256;

// sqrt.approx (a MUFU op, <= 2 ulp): the callers widen their bounds by 1e-5 relative
__device__ __forceinline__ float approx_sqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__global__ void __launch_bounds__(kBoundThreads)
adds_bounds_kernel(const float *__restrict__ verts, int64_t nv, const double *__restrict__ poses,
                   const double *__restrict__ centroid, const float4 *__restrict__ stage_c, int stages,
                   const float4 *__restrict__ sub_c, double *__restrict__ out_lower,
                   double *__restrict__ out_upper) {
    extern __shared__ float4 sph[];  // [stages] stage spheres, then [stages * 16] sub-tile spheres
    __shared__ double red[2][kBoundThreads / 32];
    constexpr int kSubs = ISR_SOA_TILE / ISR_SUB_TILE;
    const int64_t k = blockIdx.x;
    for (int i = threadIdx.x; i < stages; i += kBoundThreads) sph[i] = stage_c[i];
    for (int i = threadIdx.x; i < stages * kSubs; i += kBoundThreads) sph[stages + i] = sub_c[i];
    double P[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) P[j] = poses[16 * k + j];
    // (the prepared target is centred on its centroid: fold the centre into the translation)
    P[3] -= centroid[0]; P[7] -= centroid[1]; P[11] -= centroid[2];
    __syncthreads();
    double sumL = 0.0, sumU = 0.0;
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    // (whole warps iterate together: the trip count is rounded up to the CTA's stride)
    for (int64_t i0 = 0; i0 < nv; i0 += kBoundThreads) {
        const int64_t i = i0 + threadIdx.x;
        const bool live = i < nv;
        const int64_t ii = live ? i : nv - 1;
        const double vx = verts[3 * ii], vy = verts[3 * ii + 1], vz = verts[3 * ii + 2];
        const float qx = (float)(P[0] * vx + P[1] * vy + P[2] * vz + P[3]);
        const float qy = (float)(P[4] * vx + P[5] * vy + P[6] * vz + P[7]);
        const float qz = (float)(P[8] * vx + P[9] * vy + P[10] * vz + P[11]);
        // pass 1, once per WARP: its 32 queries are consecutive points of the (curve-ordered)
        // vertex cloud, i.e. one small patch with bounding sphere (cw, rho).  For every query of
        // the warp and every stage s:  |q - c_s| - r_s >= |cw - c_s| - r_s - rho =: lw_s.  The
        // lanes share the stages, keep the kExamine smallest lw_s, and the smallest of the others
        // (`restw`) bounds every point of every stage that is not examined.
        float cx = qx, cy = qy, cz = qz;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cx += __shfl_xor_sync(full, cx, o); cy += __shfl_xor_sync(full, cy, o); cz += __shfl_xor_sync(full, cz, o);
        }
        cx *= (1.f / 32.f); cy *= (1.f / 32.f); cz *= (1.f / 32.f);
        float rho;
        {
            const float dx = qx - cx, dy = qy - cy, dz = qz - cz;
            rho = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rho = fmaxf(rho, __shfl_xor_sync(full, rho, o));
            rho = approx_sqrt(rho) * 1.00001f + 1e-5f;
        }
        float lw[kExamine];   // this lane's smallest lw_s over its share of the stages, ascending
        int id[kExamine];
#pragma unroll
        for (int e = 0; e < kExamine; ++e) { lw[e] = CUDART_INF_F; id[e] = -1; }
        float restw = CUDART_INF_F;
        for (int s = lane; s < stages; s += 32) {
            const float4 S = sph[s];
            const float dx = cx - S.x, dy = cy - S.y, dz = cz - S.z;
            float d = S.w < 0.f ? CUDART_INF_F
                                : approx_sqrt(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) * 0.99999f - S.w - rho;
            int sid = s;
#pragma unroll
            for (int e = 0; e < kExamine; ++e) {  // insertion into the sorted list; the loser moves on
                const bool lt = d < lw[e];
                const float tl = lt ? lw[e] : d;
                const int ti = lt ? id[e] : sid;
                lw[e] = lt ? d : lw[e];
                id[e] = lt ? sid : id[e];
                d = tl;
                sid = ti;
            }
            restw = fminf(restw, d);
        }
        // merge the lanes' lists: kExamine rounds of "smallest head wins" (order-preserving keys)
        int sel[kExamine];
#pragma unroll
        for (int e = 0; e < kExamine; ++e) {
            const unsigned bits = __float_as_uint(lw[0]);
            const unsigned key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
            const unsigned m = __reduce_min_sync(full, key);
            const int owner = __ffs(__ballot_sync(full, key == m)) - 1;
            sel[e] = __shfl_sync(full, id[0], owner);
            if (lane == owner) {
#pragma unroll
                for (int j = 0; j + 1 < kExamine; ++j) { lw[j] = lw[j + 1]; id[j] = id[j + 1]; }
                lw[kExamine - 1] = CUDART_INF_F; id[kExamine - 1] = -1;
            }
        }
        restw = fminf(restw, lw[0]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) restw = fminf(restw, __shfl_xor_sync(full, restw, o));
        float L = fmaxf(restw, 0.f), U = CUDART_INF_F;
        // pass 2, per query: the sub-tiles of the examined stages
#pragma unroll
        for (int e = 0; e < kExamine; ++e) {
            if (sel[e] < 0) continue;  // warp-uniform
            const float4 *sub = sph + stages + sel[e] * kSubs;
#pragma unroll 4
            for (int t = 0; t < kSubs; ++t) {
                const float4 S = sub[t];
                const float dx = qx - S.x, dy = qy - S.y, dz = qz - S.z;
                const float dist = approx_sqrt(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                const bool ok = S.w >= 0.f;
                L = fminf(L, ok ? fmaxf(dist * 0.99999f - S.w, 0.f) : CUDART_INF_F);
                U = fminf(U, ok ? dist * 1.00001f + S.w : CUDART_INF_F);
            }
        }
        if (!live) continue;
        // outward margins for the float32 evaluation (coordinates are object-sized: centred)
        sumL += (double)fmaxf(L * 0.99999f - 1e-4f, 0.f);
        sumU += (double)(U * 1.00001f + 1e-4f);
    }
    sumL = warp_sum(sumL);
    sumU = warp_sum(sumU);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sumL; red[1][threadIdx.x >> 5] = sumU; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < kBoundThreads / 32; ++w) { a += red[0][w]; c += red[1][w]; }
        out_lower[k] = a / (double)nv;
        out_upper[k] = c / (double)nv;
    }
}

}  // namespace isr

extern "C" {

int isr_adds_bounds(const float *cloud_q, int64_t nq, const double *poses_q, int64_t b,
                    const double *centroid, const float *stage_c, int64_t stages, const float *sub_c,
                    double *out_lower, double *out_upper, void *stream) {
    using namespace isr;
    ISR_REQUIRE(nq >= 1 && b >= 0 && stages >= 1, ISR_E_SHAPE, "adds_bounds: bad size");
    if (b == 0) return ISR_OK;
    ISR_REQUIRE(cloud_q && poses_q && centroid && stage_c && sub_c && out_lower && out_upper, ISR_E_INVALID_ARG,
                "adds_bounds: null pointer");
    ISR_REQUIRE(aligned16(stage_c) && aligned16(sub_c), ISR_E_ALIGN, "adds_bounds: spheres must be 16-byte aligned");
    const size_t smem = (size_t)stages * (1 + ISR_SOA_TILE / ISR_SUB_TILE) * sizeof(float4);
    ISR_REQUIRE(smem <= 200 * 1024, ISR_E_SHAPE, "adds_bounds: %lld target stages do not fit shared memory",
                (long long)stages);
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        ISR_TRY(check_cuda(cudaFuncSetAttribute(adds_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                200 * 1024),
                           "adds_bounds smem attr"));
        configured_dev = dev;
    }
    adds_bounds_kernel<<<(unsigned)b, kBoundThreads, smem, (cudaStream_t)stream>>>(
        cloud_q, nq, poses_q, centroid, reinterpret_cast<const float4 *>(stage_c), (int)stages,
        reinterpret_cast<const float4 *>(sub_c), out_lower, out_upper);
    return launched("adds_bounds_kernel");
}

}  // extern "C"
