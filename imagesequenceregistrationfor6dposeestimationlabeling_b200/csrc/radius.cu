// Radius neighbour count -- the arithmetic of Open3D's PointCloud.remove_radius_outlier
// (generateCors.py:254-258, trainPose.py:343-347: the filter that cleans the NeRF surface
// clouds before they enter verification / ICP).
//
// Upstream semantics restated: for every point a KD-tree radius search (nanoflann, float64,
// STRICT d^2 < radius^2, the point itself included) counts its neighbours; the point is kept
// iff count > nb_points.  This kernel returns the counts; the selection is host-side.
//
// Same building blocks as the pruned nearest-neighbour search (nn2.cu), with a fixed bound
// instead of a shrinking one: clouds stored along the Hilbert curve, one warp per 256
// consecutive queries, stage spheres tested against the warp's query sphere, 64-target
// sub-tile spheres tested against every query (radius + sphere radius), surviving sub-tiles
// evaluated pair by pair on the FP32 CUDA cores in the 3-FMA form v = |p|^2 - 2 q.p + |q|^2.
// v carries a bounded cancellation error E, so it only decides pairs outside the band
// [r^2 - E, r^2 + E]; the rare pair inside the band is decided in FP64 from the hi/lo
// coordinates.  Counts therefore equal the float64 brute-force counts.
#include <math_constants.h>

#include "isr_common.cuh"

namespace isr {

struct RadiusParams {
    const float *q;   // SoA7 [7][nq_pad]
    int nq, nq_pad;
    const float *t;   // SoA7 [7][nt_pad]
    int nt, nt_pad;
    const float4 *stage_c, *sub_c;
    int stages;
    const int *perm_q;
    double r2;        // radius^2
    float r;          // radius, rounded up
    int *out;
};

constexpr int kRQ = 8;  // queries per lane

__global__ void __launch_bounds__(128) radius_count_kernel(const RadiusParams p) {
    constexpr int SUBS = ISR_SOA_TILE / ISR_SUB_TILE;
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
    const int q0 = warp * (32 * kRQ) + lane;
    if (q0 - lane >= p.nq) return;
    const float *__restrict__ gq = p.q;
    const float *__restrict__ gt = p.t;

    float q2x[kRQ], q2y[kRQ], q2z[kRQ], nq2[kRQ], thr_in[kRQ], thr_out[kRQ];
    int cnt[kRQ];
    float reach = 0.f;  // per-lane bound for the sphere tests: radius (+ slack), 0 for dead lanes
    const float u = 5.9604645e-8f;
#pragma unroll
    for (int r = 0; r < kRQ; ++r) {
        const int i = min(q0 + r * 32, p.nq_pad - 1);
        const float x = gq[i], y = gq[p.nq_pad + i], z = gq[2ll * p.nq_pad + i];
        q2x[r] = -2.0f * x; q2y[r] = -2.0f * y; q2z[r] = -2.0f * z;
        nq2[r] = __fmaf_rn(z, z, __fmaf_rn(y, y, x * x));
        cnt[r] = 0;
        const bool live = q0 + r * 32 < p.nq;
        // |v - d^2| <= 13 u (|q| + d)^2 + rounding of |q|^2 and of the final add; the band
        // below covers every target within 1.5 r, and targets beyond 1.5 r land above it as
        // long as |q| + 1.5 r < 600 r.  Otherwise (a radius tiny against the coordinates)
        // every pair of a surviving tile is decided in FP64.
        const float qn = sqrtf(nq2[r]);
        const float R = qn + 1.5f * p.r;
        const float E = 32.f * u * R * R;
        const bool wide = !(R < 600.f * p.r);
        const float r2f = (float)p.r2;
        thr_in[r] = !live ? -CUDART_INF_F : (wide ? -CUDART_INF_F : __fsub_rd(r2f, E) * 0.999999f);
        thr_out[r] = !live ? -CUDART_INF_F : (wide ? CUDART_INF_F : __fadd_ru(r2f, E) * 1.000001f);
        if (live) reach = fmaxf(reach, p.r * 1.0001f + 1e-6f * qn);
    }
    // the warp's query sphere
    float cWx, cWy, cWz, rW;
    {
        float cx = 0.f, cy = 0.f, cz = 0.f, cn = 0.f;
#pragma unroll
        for (int r = 0; r < kRQ; ++r)
            if (q0 + r * 32 < p.nq) { cx += q2x[r]; cy += q2y[r]; cz += q2z[r]; cn += 1.f; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cx += __shfl_xor_sync(0xffffffffu, cx, o);
            cy += __shfl_xor_sync(0xffffffffu, cy, o);
            cz += __shfl_xor_sync(0xffffffffu, cz, o);
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
        }
        const float inv = -0.5f / cn;
        cWx = cx * inv; cWy = cy * inv; cWz = cz * inv;
        float m = 0.f;
#pragma unroll
        for (int r = 0; r < kRQ; ++r) {
            if (q0 + r * 32 < p.nq) {
                const float dx = fmaf(q2x[r], -0.5f, -cWx), dy = fmaf(q2y[r], -0.5f, -cWy),
                            dz = fmaf(q2z[r], -0.5f, -cWz);
                m = fmaxf(m, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        rW = __fsqrt_ru(m) * 1.00002f;
    }
    float reachW = reach;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) reachW = fmaxf(reachW, __shfl_xor_sync(0xffffffffu, reachW, o));

    for (int base = 0; base < p.stages; base += 32) {
        const int s = base + lane;
        bool cand = false;
        if (s < p.stages) {
            const float4 S = p.stage_c[s];
            const float dx = S.x - cWx, dy = S.y - cWy, dz = S.z - cWz;
            const float rr = (reachW + rW + S.w) * 1.0001f;
            cand = S.w >= 0.f && !(fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > rr * rr);
        }
        unsigned mask = __ballot_sync(0xffffffffu, cand);
        while (mask != 0) {
            const int st = base + __ffs(mask) - 1;
            mask &= mask - 1;
#pragma unroll 1
            for (int sub = 0; sub < SUBS; ++sub) {
                const float4 S = p.sub_c[(long long)st * SUBS + sub];
                // skip the sub-tile iff every query is farther than radius + sphere radius
                bool out = S.w < 0.f;
                if (!out) {
                    float m = CUDART_INF_F;
#pragma unroll
                    for (int r = 0; r < kRQ; ++r) {
                        const float dx = fmaf(q2x[r], -0.5f, -S.x), dy = fmaf(q2y[r], -0.5f, -S.y),
                                    dz = fmaf(q2z[r], -0.5f, -S.z);
                        m = fminf(m, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                    }
                    const float rr = (reach + S.w) * 1.0001f;
                    out = m > rr * rr;
                }
                if (__all_sync(0xffffffffu, out)) continue;
                const long long tbase = ((long long)st * SUBS + sub) * ISR_SUB_TILE;
#pragma unroll 1
                for (int g = 0; g < ISR_SUB_TILE / 4; ++g) {
                    const float4 X = *reinterpret_cast<const float4 *>(gt + tbase + 4 * g);
                    const float4 Y = *reinterpret_cast<const float4 *>(gt + p.nt_pad + tbase + 4 * g);
                    const float4 Z = *reinterpret_cast<const float4 *>(gt + 2ll * p.nt_pad + tbase + 4 * g);
                    const float4 N = *reinterpret_cast<const float4 *>(gt + 3ll * p.nt_pad + tbase + 4 * g);
                    const float px[4] = {X.x, X.y, X.z, X.w}, py[4] = {Y.x, Y.y, Y.z, Y.w},
                                pz[4] = {Z.x, Z.y, Z.z, Z.w}, pn[4] = {N.x, N.y, N.z, N.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int r = 0; r < kRQ; ++r) {
                            const float v = __fmaf_rn(q2x[r], px[k], __fmaf_rn(q2y[r], py[k],
                                                      __fmaf_rn(q2z[r], pz[k], pn[k]))) + nq2[r];
                            if (v < thr_in[r]) {
                                ++cnt[r];
                            } else if (v <= thr_out[r]) {
                                // inside the error band: decide in FP64 from hi + lo
                                const long long j = tbase + 4 * g + k;
                                const int i = min(q0 + r * 32, p.nq_pad - 1);
                                if (j < p.nt) {
                                    const double dx = ((double)gq[i] - (double)px[k]) +
                                                      ((double)gq[4ll * p.nq_pad + i] - (double)gt[4ll * p.nt_pad + j]);
                                    const double dy = ((double)gq[p.nq_pad + i] - (double)py[k]) +
                                                      ((double)gq[5ll * p.nq_pad + i] - (double)gt[5ll * p.nt_pad + j]);
                                    const double dz = ((double)gq[2ll * p.nq_pad + i] - (double)pz[k]) +
                                                      ((double)gq[6ll * p.nq_pad + i] - (double)gt[6ll * p.nt_pad + j]);
                                    if (fma(dz, dz, fma(dy, dy, dx * dx)) < p.r2) ++cnt[r];
                                }
                            }
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRQ; ++r) {
        const int i = q0 + r * 32;
        if (i < p.nq) p.out[p.perm_q != nullptr ? p.perm_q[i] : i] = cnt[r];
    }
}

}  // namespace isr

extern "C" {

int isr_radius_count(const IsrCloud *q, const IsrCloud *t, double radius, int32_t *out_count,
                     void *stream) {
    using namespace isr;
    ISR_REQUIRE(q != nullptr && t != nullptr && out_count != nullptr, ISR_E_INVALID_ARG,
                "radius_count: null pointer");
    ISR_REQUIRE(q->n >= 0 && t->n >= 1, ISR_E_SHAPE, "radius_count: need nq >= 0, nt >= 1");
    ISR_REQUIRE(radius > 0.0 && radius < 1e18, ISR_E_INVALID_ARG, "radius_count: radius must be positive");
    if (q->n == 0) return ISR_OK;
    ISR_REQUIRE(q->soa7 && t->soa7 && t->stage_c && t->sub_c, ISR_E_INVALID_ARG,
                "radius_count: the target needs its tile spheres (isr_tile_spheres)");
    ISR_REQUIRE(q->bstride == 0 && t->bstride == 0, ISR_E_SHAPE, "radius_count: single clouds only");
    ISR_REQUIRE(q->npad % ISR_SOA_TILE == 0 && t->npad % ISR_SOA_TILE == 0 && q->npad >= q->n &&
                    t->npad >= t->n && t->npad < (1ll << 31) - 2048 && q->npad < (1ll << 31) - 2048,
                ISR_E_SHAPE, "radius_count: bad padded length");
    ISR_REQUIRE(aligned16(t->soa7) && aligned16(t->stage_c) && aligned16(t->sub_c), ISR_E_ALIGN,
                "radius_count: target planes and spheres must be 16-byte aligned");
    RadiusParams p;
    p.q = q->soa7; p.nq = (int)q->n; p.nq_pad = (int)q->npad;
    p.t = t->soa7; p.nt = (int)t->n; p.nt_pad = (int)t->npad;
    p.stage_c = reinterpret_cast<const float4 *>(t->stage_c);
    p.sub_c = reinterpret_cast<const float4 *>(t->sub_c);
    p.stages = (int)(t->npad / ISR_SOA_TILE);
    p.perm_q = q->perm;
    p.r2 = radius * radius;
    p.r = __builtin_nextafterf((float)radius, 3.0e38f);
    p.out = out_count;
    const long long warps = (q->n + 32 * kRQ - 1) / (32 * kRQ);
    radius_count_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, (cudaStream_t)stream>>>(p);
    return launched("radius_count_kernel");
}

}  // extern "C"
