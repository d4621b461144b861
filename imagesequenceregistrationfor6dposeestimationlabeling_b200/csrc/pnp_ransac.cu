// PnP-RANSAC hypothesis GENERATION on the device (second half): what
// cv2.solvePnPRansac(h3d, h2d, cam, iterationsCount, reprojectionError, flags=SOLVEPNP_P3P) does
// between its inputs and the consensus test of pnp.cu (choosePose.py:23-33, 280-300).
//
//   p3p_hypotheses_kernel  one thread = one RANSAC iteration: 4 distinct correspondences from a
//                          counter-based generator (splitmix64 of seed / iteration / draw -- any
//                          iteration can be reproduced on its own), Grunert's P3P on the first three
//                          (quartic in the depth ratio, all real roots, FP64), the fourth point picks
//                          the solution with the smallest reprojection error (OpenCV's p3p::solve
//                          with 4 points does the same) -> one pose per iteration, NaN when the
//                          sample is degenerate (it then reprojects nothing and counts 0 inliers)
//   (isr_pnp_score)        consensus of all hypotheses in one launch
//   (first_max_kernel)     the first hypothesis with the largest consensus
//   pnp_refine_kernel      single CTA: Gauss-Newton on the reprojection error over the winner's
//                          inliers (OpenCV refits the winner on its inliers, too: EPnP), inliers
//                          re-evaluated between rounds (locally optimised RANSAC)
// OpenCV's own random stream is not reproduced (it is an implementation detail of cv::RNG and of
// the order RANSAC consumes it); parity is stated on the OUTCOME: the consensus of the returned
// pose is at least that of cv2's on the committed cv2 fixtures (tests/golden/reference_pnp_cv2.npz),
// and the minimal solver reproduces cv2.solveP3P's solutions on the committed 3-point sets.
// Work is tiny (thousands of hypotheses x a few hundred FP64 operations): latency-bound, two
// launches plus the scoring.
#include <math_constants.h>

#include "isr_common.cuh"

extern "C" int isr_first_max(const int32_t *v, int64_t n, int64_t *out2, void *stream);

namespace isr {

struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    const double d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// real roots of c[4] x^4 + ... + c[0] (Durand-Kerner on the monic polynomial, then Newton polish
// of the real parts); returns their number
__device__ int quartic_real_roots(const double c[5], double roots[4]) {
    if (!(fabs(c[4]) > 1e-300)) return 0;
    const double m[4] = {c[0] / c[4], c[1] / c[4], c[2] / c[4], c[3] / c[4]};  // x^4 + m3 x^3 + m2 x^2 + m1 x + m0
    double rad = 0.0;
    for (int k = 0; k < 4; ++k) rad = fmax(rad, fabs(m[k]));
    rad = 1.0 + rad;  // Cauchy bound
    if (!(rad < 1e150)) return 0;
    cplx z[4];
    {
        cplx w = {rad * 0.4, rad * 0.9 * 0.5};
        const cplx step = {0.4, 0.9};
        for (int k = 0; k < 4; ++k) {
            z[k] = w;
            w = cmul(w, step);
        }
    }
    for (int it = 0; it < 60; ++it) {
        double moved = 0.0;
        for (int i = 0; i < 4; ++i) {
            cplx p = {1.0, 0.0};
            p = cmul(p, z[i]); p.re += m[3];
            p = cmul(p, z[i]); p.re += m[2];
            p = cmul(p, z[i]); p.re += m[1];
            p = cmul(p, z[i]); p.re += m[0];
            cplx q = {1.0, 0.0};
            for (int j = 0; j < 4; ++j)
                if (j != i) q = cmul(q, csub(z[i], z[j]));
            if (q.re * q.re + q.im * q.im < 1e-300) continue;
            const cplx d = cdiv(p, q);
            z[i] = csub(z[i], d);
            moved = fmax(moved, fabs(d.re) + fabs(d.im));
        }
        if (moved < 1e-14 * rad) break;
    }
    int n = 0;
    for (int i = 0; i < 4; ++i) {
        if (fabs(z[i].im) > 1e-6 * fmax(1.0, fabs(z[i].re))) continue;
        double x = z[i].re;
        for (int k = 0; k < 3; ++k) {  // Newton polish on the real line
            const double p = (((x + m[3]) * x + m[2]) * x + m[1]) * x + m[0];
            const double dp = ((4.0 * x + 3.0 * m[3]) * x + 2.0 * m[2]) * x + m[1];
            if (fabs(dp) < 1e-300) break;
            x -= p / dp;
        }
        roots[n++] = x;
    }
    return n;
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ bool normalize3(double a[3]) {
    const double n = sqrt(dot3(a, a));
    if (!(n > 1e-300)) return false;
    a[0] /= n; a[1] /= n; a[2] /= n;
    return true;
}
// orthonormal frame of a triangle: e1 along X1 - X0, e3 its normal, e2 = e3 x e1 (columns of F)
__device__ bool tri_frame(const double X[3][3], double F[3][3]) {
    double e1[3] = {X[1][0] - X[0][0], X[1][1] - X[0][1], X[1][2] - X[0][2]};
    double w[3] = {X[2][0] - X[0][0], X[2][1] - X[0][1], X[2][2] - X[0][2]};
    double e3[3], e2[3];
    if (!normalize3(e1)) return false;
    cross3(e1, w, e3);
    if (!normalize3(e3)) return false;
    cross3(e3, e1, e2);
    for (int r = 0; r < 3; ++r) { F[r][0] = e1[r]; F[r][1] = e2[r]; F[r][2] = e3[r]; }
    return true;
}

// Grunert's solution of the perspective three-point problem (as in Haralick et al. 1994):
// object points P[3][3], unit bearing vectors f[3][3] in the camera frame -> up to 4 poses
// (row-major 3x4 [R | t], camera = R object + t).  Returns their number.
__device__ int p3p_grunert(const double P[3][3], const double f[3][3], double out[4][12]) {
    double d12[3], d02[3], d01[3];
    for (int k = 0; k < 3; ++k) {
        d12[k] = P[1][k] - P[2][k];
        d02[k] = P[0][k] - P[2][k];
        d01[k] = P[0][k] - P[1][k];
    }
    const double a2 = dot3(d12, d12), b2 = dot3(d02, d02), c2 = dot3(d01, d01);
    if (!(a2 > 1e-300 && b2 > 1e-300 && c2 > 1e-300)) return 0;
    const double ca = dot3(f[1], f[2]), cb = dot3(f[0], f[2]), cg = dot3(f[0], f[1]);
    const double q1 = (a2 - c2) / b2, q2 = (a2 + c2) / b2, q3 = (b2 - c2) / b2, q4 = (b2 - a2) / b2;
    double c[5];
    c[4] = (q1 - 1.0) * (q1 - 1.0) - 4.0 * c2 / b2 * ca * ca;
    c[3] = 4.0 * (q1 * (1.0 - q1) * cb - (1.0 - q2) * ca * cg + 2.0 * c2 / b2 * ca * ca * cb);
    c[2] = 2.0 * (q1 * q1 - 1.0 + 2.0 * q1 * q1 * cb * cb + 2.0 * q3 * ca * ca - 4.0 * q2 * ca * cb * cg +
                  2.0 * q4 * cg * cg);
    c[1] = 4.0 * (-q1 * (1.0 + q1) * cb + 2.0 * a2 / b2 * cg * cg * cb - (1.0 - q2) * ca * cg);
    c[0] = (1.0 + q1) * (1.0 + q1) - 4.0 * a2 / b2 * cg * cg;
    double roots[4];
    const int nr = quartic_real_roots(c, roots);
    double Fp[3][3];
    if (!tri_frame(P, Fp)) return 0;
    int n = 0;
    for (int r = 0; r < nr; ++r) {
        const double v = roots[r];
        if (!(v > 0.0)) continue;
        const double den = 2.0 * (cg - v * ca);
        if (!(fabs(den) > 1e-12)) continue;
        const double u = ((q1 - 1.0) * v * v - 2.0 * q1 * cb * v + 1.0 + q1) / den;
        if (!(u > 0.0)) continue;
        const double s1sq = b2 / (1.0 + v * v - 2.0 * v * cb);
        if (!(s1sq > 0.0)) continue;
        const double s1 = sqrt(s1sq), s2 = u * s1, s3 = v * s1;
        double Q[3][3], Fq[3][3];
        for (int k = 0; k < 3; ++k) { Q[0][k] = s1 * f[0][k]; Q[1][k] = s2 * f[1][k]; Q[2][k] = s3 * f[2][k]; }
        if (!tri_frame(Q, Fq)) continue;
        double *o = out[n];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j)  // R = Fq Fp^T
                o[4 * i + j] = Fq[i][0] * Fp[j][0] + Fq[i][1] * Fp[j][1] + Fq[i][2] * Fp[j][2];
        }
        for (int i = 0; i < 3; ++i)
            o[4 * i + 3] = Q[0][i] - (o[4 * i] * P[0][0] + o[4 * i + 1] * P[0][1] + o[4 * i + 2] * P[0][2]);
        bool ok = true;
        for (int k = 0; k < 12; ++k) ok = ok && isfinite(o[k]);
        if (ok) ++n;
    }
    return n;
}

// bearing of pixel (u, v) for the camera matrix K = [fx s cx; 0 fy cy; 0 0 1]
__device__ __forceinline__ void bearing(const double K[9], double u, double v, double f[3]) {
    const double y = (v - K[5]) / K[4];
    const double x = (u - K[2] - K[1] * y) / K[0];
    const double n = sqrt(x * x + y * y + 1.0);
    f[0] = x / n; f[1] = y / n; f[2] = 1.0 / n;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(128)
p3p_hypotheses_kernel(const float *__restrict__ p3d, const float *__restrict__ p2d, int64_t n,
                      const double *__restrict__ cam, int64_t iterations, unsigned long long seed,
                      double *__restrict__ out_poses) {
    const int64_t it = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (it >= iterations) return;
    double K[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) K[k] = cam[k];
    // four distinct correspondences
    long long id[4];
    unsigned long long ctr = 0;
    for (int k = 0; k < 4; ++k) {
        for (;;) {
            const unsigned long long h = splitmix64(splitmix64(seed ^ ((unsigned long long)it * 0xD1342543DE82EF95ull)) + ctr++);
            const long long cand = (long long)(h % (unsigned long long)n);
            bool dup = false;
            for (int j = 0; j < k; ++j) dup = dup || id[j] == cand;
            if (!dup || ctr > 64) { id[k] = cand; break; }
        }
    }
    double P[3][3], f[3][3];
    for (int k = 0; k < 3; ++k) {
        P[k][0] = p3d[3 * id[k]]; P[k][1] = p3d[3 * id[k] + 1]; P[k][2] = p3d[3 * id[k] + 2];
        bearing(K, p2d[2 * id[k]], p2d[2 * id[k] + 1], f[k]);
    }
    double sol[4][12];
    const int ns = p3p_grunert(P, f, sol);
    // the fourth correspondence picks the solution (smallest squared reprojection error)
    const double X = p3d[3 * id[3]], Y = p3d[3 * id[3] + 1], Z = p3d[3 * id[3] + 2];
    const double u4 = p2d[2 * id[3]], v4 = p2d[2 * id[3] + 1];
    int best = -1;
    double best_e = CUDART_INF;
    for (int s = 0; s < ns; ++s) {
        const double *o = sol[s];
        const double x = o[0] * X + o[1] * Y + o[2] * Z + o[3];
        const double y = o[4] * X + o[5] * Y + o[6] * Z + o[7];
        const double z = o[8] * X + o[9] * Y + o[10] * Z + o[11];
        if (!(z > 0.0)) continue;  // (in front of the camera)
        const double xn = x / z, yn = y / z;
        const double du = K[0] * xn + K[1] * yn + K[2] - u4, dv = K[4] * yn + K[5] - v4;
        const double e = du * du + dv * dv;
        if (e < best_e) { best_e = e; best = s; }
    }
    double *o = out_poses + 16 * it;
    for (int k = 0; k < 12; ++k) o[k] = best >= 0 ? sol[best][k] : CUDART_NAN;
    o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 1.0;
}

// all P3P solutions of explicit 3-point sets (tests: against cv2.solveP3P)
__global__ void __launch_bounds__(64)
p3p_solve_kernel(const double *__restrict__ pts, const double *__restrict__ uv, const double *__restrict__ cam,
                 int64_t b, double *__restrict__ out_poses, int32_t *__restrict__ out_n) {
    const int64_t i = (int64_t)blockIdx.x * 64 + threadIdx.x;
    if (i >= b) return;
    double K[9], P[3][3], f[3][3];
    for (int k = 0; k < 9; ++k) K[k] = cam[k];
    for (int k = 0; k < 3; ++k) {
        for (int c = 0; c < 3; ++c) P[k][c] = pts[9 * i + 3 * k + c];
        bearing(K, uv[6 * i + 2 * k], uv[6 * i + 2 * k + 1], f[k]);
    }
    double sol[4][12];
    const int ns = p3p_grunert(P, f, sol);
    out_n[i] = ns;
    for (int s = 0; s < 4; ++s)
        for (int k = 0; k < 16; ++k)
            out_poses[(4 * i + s) * 16 + k] = s < ns ? (k < 12 ? sol[s][k] : (k == 15 ? 1.0 : 0.0)) : CUDART_NAN;
}

// ---- refit of the winner on its inliers: Gauss-Newton on the reprojection error ----------------
constexpr int kRefThreads = 256;
constexpr int kRefSums = 27;  // upper triangle of J^T J (21) + J^T r (6)

__device__ __forceinline__ void rodrigues(const double w[3], double R[9]) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double th = sqrt(th2);
    double a, b;  // R = I + a [w]x + b [w]x^2
    if (th < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; }
    const double x = w[0], y = w[1], z = w[2];
    R[0] = 1.0 - b * (y * y + z * z); R[1] = -a * z + b * x * y;        R[2] = a * y + b * x * z;
    R[3] = a * z + b * x * y;         R[4] = 1.0 - b * (x * x + z * z); R[5] = -a * x + b * y * z;
    R[6] = -a * y + b * x * z;        R[7] = a * x + b * y * z;         R[8] = 1.0 - b * (x * x + y * y);
}

// single CTA.  pose: in/out 4x4; `best` (device): index of the winning hypothesis in `hyps`.
// out_inlier (uint8 [n]): the consensus set of the WINNING HYPOTHESIS (what cv2 returns as
// `inliers`); out_counts[0] = its size, out_counts[1] = the consensus of the refined pose.
__global__ void __launch_bounds__(kRefThreads)
pnp_refine_kernel(const float *__restrict__ p3d, const float *__restrict__ p2d, int64_t n,
                  const double *__restrict__ cam, const double *__restrict__ hyps,
                  const int64_t *__restrict__ best, double thr2, int rounds, int gn_iters,
                  double *__restrict__ out_pose, uint8_t *__restrict__ out_inlier,
                  uint8_t *__restrict__ work_flag, int32_t *__restrict__ out_counts) {
    __shared__ double T[12];
    __shared__ double red[kRefThreads / 32][kRefSums];
    __shared__ double tot[kRefSums];
    __shared__ int cnt_s[kRefThreads / 32];
    __shared__ int cnt_total;
    __shared__ int stop;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double fx = cam[0], sk = cam[1], cx = cam[2], fy = cam[4], cy = cam[5];
    if (tid < 12) T[tid] = hyps[16 * best[0] + tid];
    if (tid == 0) stop = 0;
    __syncthreads();

    // the consensus test of pnp_score_kernel (OpenCV's arithmetic) for pose T -> flags, count
    auto consensus = [&](uint8_t *flags) {
        int cnt = 0;
        for (int64_t i = tid; i < n; i += kRefThreads) {
            const double X = p3d[3 * i], Y = p3d[3 * i + 1], Z = p3d[3 * i + 2];
            const double x = ((T[0] * X + T[1] * Y) + T[2] * Z) + T[3];
            const double y = ((T[4] * X + T[5] * Y) + T[6] * Z) + T[7];
            double z = ((T[8] * X + T[9] * Y) + T[10] * Z) + T[11];
            z = z != 0.0 ? 1.0 / z : 1.0;
            const double xn = x * z, yn = y * z;
            const double pu = fx * xn + sk * yn + cx, pv = fy * yn + cy;
            const float du = p2d[2 * i] - (float)pu, dv = p2d[2 * i + 1] - (float)pv;
            const float e = (float)((double)du * (double)du + (double)dv * (double)dv);
            const bool in = (double)e <= thr2;
            flags[i] = in ? 1 : 0;
            cnt += in ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) cnt_s[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            int s = 0;
            for (int w = 0; w < kRefThreads / 32; ++w) s += cnt_s[w];
            cnt_total = s;
        }
        __syncthreads();
        return cnt_total;
    };

    const bool valid = isfinite(T[0]);
    if (tid < 16) out_pose[tid] = tid < 12 ? T[tid] : (tid == 15 ? 1.0 : 0.0);  // the winner itself, unless a refit beats it
    int c0 = valid ? consensus(out_inlier) : 0;
    if (!valid) {
        for (int64_t i = tid; i < n; i += kRefThreads) out_inlier[i] = 0;
    }
    if (tid == 0) { out_counts[0] = c0; out_counts[1] = c0; }
    for (int64_t i = tid; i < n; i += kRefThreads) work_flag[i] = out_inlier[i];
    __syncthreads();
    int cur = c0;
    for (int round = 0; round < rounds && cur >= 4; ++round) {
        for (int it = 0; it < gn_iters; ++it) {
            double acc[kRefSums];
#pragma unroll
            for (int k = 0; k < kRefSums; ++k) acc[k] = 0.0;
            for (int64_t i = tid; i < n; i += kRefThreads) {
                if (!work_flag[i]) continue;
                const double X = p3d[3 * i], Y = p3d[3 * i + 1], Z = p3d[3 * i + 2];
                const double x = T[0] * X + T[1] * Y + T[2] * Z + T[3];
                const double y = T[4] * X + T[5] * Y + T[6] * Z + T[7];
                const double z = T[8] * X + T[9] * Y + T[10] * Z + T[11];
                if (!(fabs(z) > 1e-12)) continue;
                const double iz = 1.0 / z, xn = x * iz, yn = y * iz;
                const double ru = fx * xn + sk * yn + cx - (double)p2d[2 * i];
                const double rv = fy * yn + cy - (double)p2d[2 * i + 1];
                // d(u, v) / d(p_c): rows A (u), B (v)
                const double A[3] = {fx * iz, sk * iz, -(fx * xn + sk * yn) * iz};
                const double B[3] = {0.0, fy * iz, -fy * yn * iz};
                // p_c' = exp(w) p_c + d  =>  d p_c = w x p_c + d = -[p_c]x w + d
                const double Ju[6] = {A[1] * (-z) - A[2] * (-y) , A[2] * (-x) - A[0] * (-z), A[0] * (-y) - A[1] * (-x), A[0], A[1], A[2]};
                const double Jv[6] = {B[1] * (-z) - B[2] * (-y) , B[2] * (-x) - B[0] * (-z), B[0] * (-y) - B[1] * (-x), B[0], B[1], B[2]};
                int k = 0;
#pragma unroll
                for (int a = 0; a < 6; ++a)
#pragma unroll
                    for (int b = a; b < 6; ++b) acc[k++] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
#pragma unroll
                for (int a = 0; a < 6; ++a) acc[21 + a] += Ju[a] * ru + Jv[a] * rv;
            }
#pragma unroll
            for (int k = 0; k < kRefSums; ++k) acc[k] = warp_sum(acc[k]);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < kRefSums; ++k) red[warp][k] = acc[k];
            }
            __syncthreads();
            if (tid < kRefSums) {
                double s = 0.0;
                for (int w = 0; w < kRefThreads / 32; ++w) s += red[w][tid];
                tot[tid] = s;
            }
            __syncthreads();
            if (tid == 0) {
                // solve (H + lambda diag H) x = -g by Cholesky
                double H[6][6], g[6], x[6];
                int k = 0;
                for (int a = 0; a < 6; ++a)
                    for (int b = a; b < 6; ++b) { H[a][b] = tot[k]; H[b][a] = tot[k]; ++k; }
                for (int a = 0; a < 6; ++a) { g[a] = -tot[21 + a]; H[a][a] *= 1.0 + 1e-9; }
                bool ok = true;
                for (int a = 0; a < 6 && ok; ++a) {
                    for (int b = 0; b <= a; ++b) {
                        double s = H[a][b];
                        for (int c = 0; c < b; ++c) s -= H[a][c] * H[b][c];
                        if (a == b) {
                            if (!(s > 0.0)) { ok = false; break; }
                            H[a][a] = sqrt(s);
                        } else {
                            H[a][b] = s / H[b][b];
                        }
                    }
                }
                if (ok) {
                    for (int a = 0; a < 6; ++a) {
                        double s = g[a];
                        for (int c = 0; c < a; ++c) s -= H[a][c] * x[c];
                        x[a] = s / H[a][a];
                    }
                    for (int a = 5; a >= 0; --a) {
                        double s = x[a];
                        for (int c = a + 1; c < 6; ++c) s -= H[c][a] * x[c];
                        x[a] = s / H[a][a];
                    }
                    double Rw[9];
                    rodrigues(x, Rw);
                    double Tn[12];
                    for (int i = 0; i < 3; ++i) {
                        for (int j = 0; j < 4; ++j)
                            Tn[4 * i + j] = Rw[3 * i] * T[j] + Rw[3 * i + 1] * T[4 + j] + Rw[3 * i + 2] * T[8 + j];
                        Tn[4 * i + 3] += x[3 + i];
                    }
                    bool fin = true;
                    for (int q = 0; q < 12; ++q) fin = fin && isfinite(Tn[q]);
                    if (fin) {
                        for (int q = 0; q < 12; ++q) T[q] = Tn[q];
                    } else {
                        stop = 1;
                    }
                } else {
                    stop = 1;
                }
            }
            __syncthreads();
            if (stop) break;
        }
        if (stop) break;
        // re-evaluate the consensus with the refined pose; keep refining while it grows
        const int c1 = consensus(work_flag);
        if (tid == 0 && c1 > out_counts[1]) {
            out_counts[1] = c1;
            for (int q = 0; q < 12; ++q) out_pose[q] = T[q];
            out_pose[12] = 0.0; out_pose[13] = 0.0; out_pose[14] = 0.0; out_pose[15] = 1.0;
        }
        __syncthreads();
        if (c1 <= cur) break;
        cur = c1;
    }
}

}  // namespace isr

extern "C" {

size_t isr_pnp_ransac_workspace_bytes(int64_t n, int64_t iterations) {
    if (n < 0 || iterations < 1) return 256;
    return isr::align256((size_t)iterations * 16 * 8) + isr::align256((size_t)iterations * 4) +
           isr::align256((size_t)(n > 0 ? n : 1)) + 256;
}

int isr_p3p_solve(const double *pts, const double *uv, const double *cam, int64_t b, double *out_poses,
                  int32_t *out_n, void *stream) {
    using namespace isr;
    ISR_REQUIRE(b >= 0, ISR_E_SHAPE, "p3p_solve: b < 0");
    if (b == 0) return ISR_OK;
    ISR_REQUIRE(pts && uv && cam && out_poses && out_n, ISR_E_INVALID_ARG, "p3p_solve: null pointer");
    p3p_solve_kernel<<<(unsigned)((b + 63) / 64), 64, 0, (cudaStream_t)stream>>>(pts, uv, cam, b, out_poses, out_n);
    return launched("p3p_solve_kernel");
}

int isr_pnp_ransac(const float *p3d, const float *p2d, int64_t n, const double *cam, int64_t iterations,
                   uint64_t seed, double reperr, int refine_rounds, double *out_pose, int32_t *out_counts,
                   uint8_t *out_inlier, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 4 && iterations >= 1 && iterations <= 65535ll * 16, ISR_E_SHAPE,
                "pnp_ransac: need n >= 4 correspondences and 1 <= iterations <= %d (n=%lld iterations=%lld)",
                65535 * 16, (long long)n, (long long)iterations);
    ISR_REQUIRE(p3d && p2d && cam && out_pose && out_counts && out_inlier, ISR_E_INVALID_ARG,
                "pnp_ransac: null pointer");
    ISR_REQUIRE(reperr >= 0.0 && refine_rounds >= 0, ISR_E_INVALID_ARG, "pnp_ransac: bad argument");
    ISR_REQUIRE(workspace != nullptr && workspace_bytes >= isr_pnp_ransac_workspace_bytes(n, iterations),
                ISR_E_WORKSPACE, "pnp_ransac: workspace too small");
    ISR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ISR_E_ALIGN, "pnp_ransac: workspace not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(workspace);
    double *hyps = reinterpret_cast<double *>(ws);
    int32_t *counts = reinterpret_cast<int32_t *>(ws + align256((size_t)iterations * 16 * 8));
    uint8_t *flags = reinterpret_cast<uint8_t *>(ws + align256((size_t)iterations * 16 * 8) + align256((size_t)iterations * 4));
    int64_t *best = reinterpret_cast<int64_t *>(ws + align256((size_t)iterations * 16 * 8) +
                                                align256((size_t)iterations * 4) + align256((size_t)n));
    p3p_hypotheses_kernel<<<(unsigned)((iterations + 127) / 128), 128, 0, st>>>(p3d, p2d, n, cam, iterations,
                                                                              (unsigned long long)seed, hyps);
    ISR_TRY(launched("p3p_hypotheses_kernel"));
    ISR_TRY(isr_pnp_score(p3d, p2d, n, cam, hyps, iterations, reperr, counts, nullptr, stream));
    ISR_TRY(isr_first_max(counts, iterations, best, stream));
    pnp_refine_kernel<<<1, kRefThreads, 0, st>>>(p3d, p2d, n, cam, hyps, best, reperr * reperr, refine_rounds, 5,
                                                out_pose, out_inlier, flags, out_counts);
    ISR_TRY(launched("pnp_refine_kernel"));
    return ISR_OK;
}

}  // extern "C"
