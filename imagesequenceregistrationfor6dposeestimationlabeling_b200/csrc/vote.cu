// The callers either side of the batched ADD-S scoring in choosePose.py,
// on the device so that the 1.64 M pose pairs of a 1280-image sequence never visit the host:
//   rel_pose_table_kernel   relative_poses[i][j] = [R_i^T R_j | t_j - t_i]    choosePose.py:43-51,98-107
//   rigid_relative_kernel   M = Pt^-1 Pq: ADD-S of (Pq . verts) against (Pt . surface) equals the
//                           ADD-S of (M . verts) against the surface itself, so the surface cloud
//                           is prepared ONCE for the whole table (verify.cu: isr_adds_fixed_target)
//   vote_rows_kernel        error[i][j] = loss[i][j] < threshold; votes[i] = sum_j error[i][j]
//   first_max_kernel        np.argmax(votes): the first maximum                choosePose.py:135-151
// All tiny and HBM/L2-bound (128 B per pose); integer sums, FP64 pose algebra, no atomics.
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
