// Point-cloud normals from k-nearest-neighbour PCA (second half):
// pytorch3d.ops.estimate_pointcloud_normals(points, neighborhood_size=k, disambiguate_directions=True)
// as generateCors.py:200-215 calls it on the 1000 farthest-point samples of the NeRF cloud with
// k = 400 (the reference then negates the result).  Upstream semantics restated:
//   neighbourhood  the k nearest points of the SAME cloud, the point itself included;
//   covariance     mean over the neighbourhood of (p - mean)(p - mean)^T;
//   normal         eigenvector of the smallest eigenvalue (symmetric 3x3 eigen-decomposition);
//   direction      flipped when fewer than k / 2 neighbours have a positive projection
//                  (p_j - p_i) . n  -- which makes the sign a function of the data, not of the
//                  eigen-solver.
// One CTA per point: all n squared distances in shared memory (FP32 direct differences: the
// selection only needs their ORDER, exact ties go to the lower index), the k-th smallest by a
// 4-pass radix select over the float bits, FP64 sums over the selected points in a fixed order,
// a cyclic Jacobi eigen-solver on one thread, a second pass for the direction.  O(n^2) on
// purpose: the reference's n is 1000 (2 MB of distances in total), a tree or the tile machinery
// of nn2.cu would cost more than it saves.  n <= kMaxPoints (shared memory).
#include <math_constants.h>

#include "isr_common.cuh"

namespace isr {

constexpr int kNrmThreads = 256;
constexpr int kNrmMaxPoints = 48 * 1024;  // 192 KB of float distances

// eigenvector of the smallest eigenvalue of the symmetric matrix C (cyclic Jacobi, FP64)
__device__ void smallest_eigenvector(const double Cin[3][3], double nrm[3]) {
    double A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { A[i][j] = Cin[i][j]; V[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-18 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    const double len = sqrt(V[0][m] * V[0][m] + V[1][m] * V[1][m] + V[2][m] * V[2][m]);
    for (int k = 0; k < 3; ++k) nrm[k] = V[k][m] / len;
}

__global__ void __launch_bounds__(kNrmThreads)
knn_normals_kernel(const float *__restrict__ pts, int n, int k, int disambiguate, float *__restrict__ out) {
    extern __shared__ float d2[];  // [n]
    __shared__ unsigned hist[256];
    __shared__ unsigned sel_prefix, sel_remaining;
    __shared__ double red[kNrmThreads / 32][9];
    __shared__ double tot[9];
    __shared__ int red_i[kNrmThreads / 32];
    __shared__ float nrm_s[3];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float qx = pts[3 * i], qy = pts[3 * i + 1], qz = pts[3 * i + 2];
    for (int j = tid; j < n; j += kNrmThreads) {
        const float dx = pts[3 * j] - qx, dy = pts[3 * j + 1] - qy, dz = pts[3 * j + 2] - qz;
        d2[j] = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    }
    // ---- k-th smallest squared distance: radix select over the (non-negative) float bits --------
    if (tid == 0) { sel_prefix = 0; sel_remaining = (unsigned)k; }
    __syncthreads();
    for (int pass = 3; pass >= 0; --pass) {
        hist[tid] = 0;  // kNrmThreads == 256 bins
        __syncthreads();
        const unsigned prefix = sel_prefix;
        const unsigned mask_hi = pass == 3 ? 0u : (0xFFFFFFFFu << (8 * (pass + 1)));
        for (int j = tid; j < n; j += kNrmThreads) {
            const unsigned b = __float_as_uint(d2[j]);
            if ((b & mask_hi) == prefix) atomicAdd(&hist[(b >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned rem = sel_remaining, bin = 0;
            for (; bin < 256; ++bin) {
                if (hist[bin] >= rem) break;
                rem -= hist[bin];
            }
            sel_prefix = prefix | (bin << (8 * pass));
            sel_remaining = rem;  // how many of the values equal to the final threshold are taken
        }
        __syncthreads();
    }
    const unsigned thr = sel_prefix;      // bits of the k-th smallest d2
    const unsigned take_eq = sel_remaining;  // ties at the threshold: the first `take_eq` by index
    // rank of every tied point among the ties (ascending index): a block-wide exclusive count
    // -- done by thread 0 for the (rare, short) tie list to stay deterministic and simple
    __shared__ int eq_limit;  // points with d2 == thr are taken iff their index <= eq_limit
    if (tid == 0) {
        unsigned seen = 0;
        int lim = -1;
        for (int j = 0; j < n && seen < take_eq; ++j)
            if (__float_as_uint(d2[j]) == thr) { ++seen; lim = j; }
        eq_limit = lim;
    }
    __syncthreads();
    auto selected = [&](int j) {
        const unsigned b = __float_as_uint(d2[j]);
        return b < thr || (b == thr && j <= eq_limit);
    };
    // ---- neighbourhood mean and covariance (FP64, fixed order) ------------------------------------
    double acc[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] = 0.0;
    for (int j = tid; j < n; j += kNrmThreads) {
        if (!selected(j)) continue;
        // (relative to the query point: the covariance is translation invariant, the sums stay small)
        const double x = (double)pts[3 * j] - (double)qx, y = (double)pts[3 * j + 1] - (double)qy,
                     z = (double)pts[3 * j + 2] - (double)qz;
        acc[0] += x; acc[1] += y; acc[2] += z;
        acc[3] += x * x; acc[4] += x * y; acc[5] += x * z;
        acc[6] += y * y; acc[7] += y * z; acc[8] += z * z;
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 9; ++c) red[warp][c] = acc[c];
    }
    __syncthreads();
    if (tid < 9) {
        double s = 0.0;
        for (int w = 0; w < kNrmThreads / 32; ++w) s += red[w][tid];
        tot[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        const double inv = 1.0 / (double)k;
        const double mx = tot[0] * inv, my = tot[1] * inv, mz = tot[2] * inv;
        double C[3][3];
        C[0][0] = tot[3] * inv - mx * mx; C[0][1] = tot[4] * inv - mx * my; C[0][2] = tot[5] * inv - mx * mz;
        C[1][1] = tot[6] * inv - my * my; C[1][2] = tot[7] * inv - my * mz; C[2][2] = tot[8] * inv - mz * mz;
        C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
        double nv[3];
        smallest_eigenvector(C, nv);
        nrm_s[0] = (float)nv[0]; nrm_s[1] = (float)nv[1]; nrm_s[2] = (float)nv[2];
    }
    __syncthreads();
    float nx = nrm_s[0], ny = nrm_s[1], nz = nrm_s[2];
    if (disambiguate) {
        int pos = 0;
        for (int j = tid; j < n; j += kNrmThreads) {
            if (!selected(j)) continue;
            const float px = pts[3 * j] - qx, py = pts[3 * j + 1] - qy, pz = pts[3 * j + 2] - qz;
            pos += (px * nx + py * ny + pz * nz) > 0.f ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
        if (lane == 0) red_i[warp] = pos;
        __syncthreads();
        if (tid == 0) {
            int s = 0;
            for (int w = 0; w < kNrmThreads / 32; ++w) s += red_i[w];
            if ((float)s < 0.5f * (float)k) { nx = -nx; ny = -ny; nz = -nz; }
            out[3 * i] = nx; out[3 * i + 1] = ny; out[3 * i + 2] = nz;
        }
    } else if (tid == 0) {
        out[3 * i] = nx; out[3 * i + 1] = ny; out[3 * i + 2] = nz;
    }
}

}  // namespace isr

extern "C" {

int isr_knn_normals(const float *pts, int64_t n, int64_t k, int disambiguate, float *out_normals, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 1 && n <= kNrmMaxPoints, ISR_E_SHAPE, "knn_normals: 1 <= n <= %d points (n=%lld)", kNrmMaxPoints,
                (long long)n);
    ISR_REQUIRE(k >= 1 && k <= n, ISR_E_SHAPE, "knn_normals: neighbourhood %lld outside 1..n", (long long)k);
    ISR_REQUIRE(pts && out_normals, ISR_E_INVALID_ARG, "knn_normals: null pointer");
    const size_t smem = (size_t)n * sizeof(float);
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_dev != dev) {
        ISR_TRY(check_cuda(cudaFuncSetAttribute(knn_normals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                kNrmMaxPoints * (int)sizeof(float)),
                           "knn_normals smem attr"));
        configured_dev = dev;
    }
    knn_normals_kernel<<<(unsigned)n, kNrmThreads, smem, (cudaStream_t)stream>>>(pts, (int)n, (int)k, disambiguate,
                                                                                out_normals);
    return launched("knn_normals_kernel");
}

}  // extern "C"
