/* CPU oracle, C part.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py header).
 *
 * Brute-force restatements of the 1-NN search that the reference delegates to
 * Open3D's KD-tree (verfication.py:97,99; icp.py:97-103,113,115) and sklearn's
 * KDTree (choosePose.py:21-22).  Brute force is the definition those trees
 * accelerate, so it pins the scipy cKDTree stand-in used by oracle.py, and it
 * fixes the tie rule the GPU path promises: lowest target index wins.
 *
 *   oracle_nn_f64        exact double search (callers thread over query chunks)
 *   oracle_nn_f32_fma    float32 direct-difference search with the same rounding
 *                        sequence as the CUDA kernel:
 *                        d2 = fma(dz,dz, fma(dy,dy, dx*dx)),  dx = q.x - p.x
 *                        -> the GPU's d2 bits and indices must equal this exactly.
 *
 * Parity unpinned w.r.t. Open3D itself (not installable here).
 */
#include <math.h>
#include <stdint.h>

void oracle_nn_f64(const double *q, int64_t nq, const double *t, int64_t nt,
                   double *out_d2, int64_t *out_idx) {
    for (int64_t i = 0; i < nq; ++i) {
        const double qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
        double best = INFINITY;
        int64_t bi = -1;
        for (int64_t j = 0; j < nt; ++j) {
            const double dx = qx - t[3 * j], dy = qy - t[3 * j + 1], dz = qz - t[3 * j + 2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best) { best = d2; bi = j; }
        }
        out_d2[i] = best;
        out_idx[i] = bi;
    }
}

void oracle_nn_f32_fma(const float *q, int64_t nq, const float *t, int64_t nt,
                       float *out_d2, int32_t *out_idx) {
    for (int64_t i = 0; i < nq; ++i) {
        const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
        float best = INFINITY;
        int32_t bi = -1;
        for (int64_t j = 0; j < nt; ++j) {
            const float dx = qx - t[3 * j], dy = qy - t[3 * j + 1], dz = qz - t[3 * j + 2];
            float d2 = dx * dx;          /* mul.rn   */
            d2 = fmaf(dy, dy, d2);       /* fma.rn   */
            d2 = fmaf(dz, dz, d2);       /* fma.rn   */
            if (d2 < best) { best = d2; bi = (int32_t)j; }
        }
        out_d2[i] = best;
        out_idx[i] = bi;
    }
}
