// K2 -- brute-force nearest neighbour (FP32 CUDA cores, direct-difference form).
//
// Replaces the KD-tree 1-NN that the reference reaches through Open3D
// (PointCloud.compute_point_cloud_distance, verfication.py:97,99 / icp.py:113,115;
// registration_icp / evaluate_registration, icp.py:97-103) and scikit-learn
// (KDTree.query(k=1), choosePose.py:21-22).
//
// Design (B200, sm_100a):
//  * Each thread keeps Q query points in registers; the CTA streams the target cloud
//    through shared memory in stages of STAGE points per coordinate plane.  A stage is
//    fetched by three 1-D bulk async copies (TMA engine, UBLKCP) that complete on an
//    mbarrier; NSTAGES stages are in flight, so global latency never reaches the math.
//  * Inner loop: one broadcast LDS.128 per plane delivers 4 targets; per query and per
//    PAIR of targets the distance is 3 FADD2 + 1 FMUL2 + 2 FFMA2 (packed f32x2: two IEEE
//    round-to-nearest operations per issue slot) and ONE FMNMX3 folds both results into
//    the running minimum.  Nothing else is issued per pair: the argmin is resolved
//    lazily -- only the id of the last SUB-point sub-tile that lowered the minimum is
//    tracked (one compare-select per query per sub-tile), and that single sub-tile is
//    re-scanned from L2 at the end with bit-identical arithmetic.  Strict "<" in tile
//    order and in the re-scan gives the lowest index on exact ties.
//  * d2 = fma(dz,dz, fma(dy,dy, dx*dx)), dx = q.x - p.x.  The |p|^2 - 2 q.p expansion is
//    NOT used: it loses the index at the 1e-4 level, and with it go the
//    tensor cores.  Roofline: FP32 CUDA-core FLOP/s; 8 algorithmic flop per pair, of
//    which 6 FP32-pipe lane-operations are executed (cap = 8/12 of FMA peak).
//  * Optional split of the target range over blockIdx.y (small query counts / tail
//    balance): partial results meet in a packed (d2 bits << 32 | idx) atomicMin, which
//    is order-independent and keeps the lowest-index tie rule.
#include <math_constants.h>
#include <stdlib.h>

#include "isr_common.cuh"

namespace isr {

struct NNParams {
    const float *q;
    long long q_bstride;
    int nq;
    int nq_pad;
    const float *t;
    long long t_bstride;
    int nt_pad;
    float *out_d2;
    int *out_idx;
    u64 *out_packed;
    const int *skip;
    long long skip_stride;
    int stages_total;
    int stages_per_split;
};

__device__ __forceinline__ float dist2_scalar(float qx, float qy, float qz, float px, float py,
                                              float pz) {
    // identical rounding sequence to the packed inner loop
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

template <int Q, int THREADS, int STAGE, int NSTAGES, int SUB, bool WITH_IDX, int MINB, int UNR>
__global__ void __launch_bounds__(THREADS, MINB) nn_kernel(const NNParams p) {
    static_assert(STAGE % SUB == 0 && SUB % 8 == 0, "tile shapes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sbuf = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSTAGES * 3 * STAGE * 4);

    const int b = blockIdx.z;
    if (p.skip != nullptr && p.skip[(long long)b * p.skip_stride] != 0) return;
    const int tid = threadIdx.x;
    const float *__restrict__ gq = p.q + (long long)b * p.q_bstride;
    const float *__restrict__ gt = p.t + (long long)b * p.t_bstride;
    const int s_begin = blockIdx.y * p.stages_per_split;
    const int nst = min(p.stages_per_split, p.stages_total - s_begin);

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NSTAGES; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int sl) {
        const int slot = sl % NSTAGES;
        float *dst = sbuf + (size_t)slot * 3 * STAGE;
        const float *src = gt + (long long)(s_begin + sl) * STAGE;
        mbar_expect_tx(&full[slot], 3u * STAGE * 4u);
        bulk_g2s(dst, src, STAGE * 4u, &full[slot]);
        bulk_g2s(dst + STAGE, src + p.nt_pad, STAGE * 4u, &full[slot]);
        bulk_g2s(dst + 2 * STAGE, src + 2ll * p.nt_pad, STAGE * 4u, &full[slot]);
    };
    if (tid == 0) {
        for (int i = 0; i < NSTAGES - 1 && i < nst; ++i) issue(i);
    }

    float qx[Q], qy[Q], qz[Q], m[Q];
    int bt[Q];
    const int q0 = blockIdx.x * (THREADS * Q);
#pragma unroll
    for (int r = 0; r < Q; ++r) {
        const int i = min(q0 + r * THREADS + tid, p.nq_pad - 1);
        qx[r] = gq[i];
        qy[r] = gq[p.nq_pad + i];
        qz[r] = gq[2ll * p.nq_pad + i];
        m[r] = CUDART_INF_F;
        bt[r] = s_begin * (STAGE / SUB);
    }

    for (int sl = 0; sl < nst; ++sl) {
        if (tid == 0 && sl + NSTAGES - 1 < nst) issue(sl + NSTAGES - 1);
        const int slot = sl % NSTAGES;
        mbar_wait(&full[slot], (sl / NSTAGES) & 1);
        const float4 *sx = reinterpret_cast<const float4 *>(sbuf + (size_t)slot * 3 * STAGE);
        const float4 *sy = sx + STAGE / 4;
        const float4 *sz = sy + STAGE / 4;
#pragma unroll 1
        for (int sub = 0; sub < STAGE / SUB; ++sub) {
            float mo[Q];
#pragma unroll
            for (int r = 0; r < Q; ++r) mo[r] = m[r];
#pragma unroll UNR
            for (int g = 0; g < SUB / 4; ++g) {
                const float4 X = sx[sub * (SUB / 4) + g];
                const float4 Y = sy[sub * (SUB / 4) + g];
                const float4 Z = sz[sub * (SUB / 4) + g];
                const u64 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
                const u64 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
                const u64 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
#pragma unroll
                for (int r = 0; r < Q; ++r) {
                    const u64 qxx = pack2(qx[r], qx[r]);
                    const u64 qyy = pack2(qy[r], qy[r]);
                    const u64 qzz = pack2(qz[r], qz[r]);
                    u64 dx = sub2(qxx, x01), dy = sub2(qyy, y01), dz = sub2(qzz, z01);
                    u64 d = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float d0, d1;
                    unpack2(d, d0, d1);
                    m[r] = min3(m[r], d0, d1);
                    dx = sub2(qxx, x23);
                    dy = sub2(qyy, y23);
                    dz = sub2(qzz, z23);
                    d = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    unpack2(d, d0, d1);
                    m[r] = min3(m[r], d0, d1);
                }
            }
            if (WITH_IDX) {
                const int tile = (s_begin + sl) * (STAGE / SUB) + sub;
#pragma unroll
                for (int r = 0; r < Q; ++r) bt[r] = (m[r] < mo[r]) ? tile : bt[r];
            }
        }
        __syncthreads();  // every warp is done with this slot before it is refilled
    }

    int bi[Q];
    if (WITH_IDX) {
        // resolve the argmin: re-scan the one sub-tile that produced the minimum
#pragma unroll
        for (int r = 0; r < Q; ++r) {
            const int base = bt[r] * SUB;
            const float4 *gx = reinterpret_cast<const float4 *>(gt + base);
            const float4 *gy = reinterpret_cast<const float4 *>(gt + p.nt_pad + base);
            const float4 *gz = reinterpret_cast<const float4 *>(gt + 2ll * p.nt_pad + base);
            float best = CUDART_INF_F;
            int besti = base;
#pragma unroll 2
            for (int j = 0; j < SUB / 4; ++j) {
                const float4 X = __ldg(gx + j), Y = __ldg(gy + j), Z = __ldg(gz + j);
                float d;
                d = dist2_scalar(qx[r], qy[r], qz[r], X.x, Y.x, Z.x);
                if (d < best) { best = d; besti = base + 4 * j; }
                d = dist2_scalar(qx[r], qy[r], qz[r], X.y, Y.y, Z.y);
                if (d < best) { best = d; besti = base + 4 * j + 1; }
                d = dist2_scalar(qx[r], qy[r], qz[r], X.z, Y.z, Z.z);
                if (d < best) { best = d; besti = base + 4 * j + 2; }
                d = dist2_scalar(qx[r], qy[r], qz[r], X.w, Y.w, Z.w);
                if (d < best) { best = d; besti = base + 4 * j + 3; }
            }
            bi[r] = besti;
            m[r] = best;
        }
    }

#pragma unroll
    for (int r = 0; r < Q; ++r) {
        const int i = q0 + r * THREADS + tid;
        if (i < p.nq) {
            const long long o = (long long)b * p.nq + i;
            if (p.out_packed != nullptr) {
                const u64 key = ((u64)__float_as_uint(m[r]) << 32) | (u64)(unsigned)(WITH_IDX ? bi[r] : 0);
                atomicMin(p.out_packed + o, key);
            } else {
                p.out_d2[o] = m[r];
                if (WITH_IDX) p.out_idx[o] = bi[r];
            }
        }
    }
}

__global__ void nn_unpack_kernel(const u64 *__restrict__ packed, long long n,
                                 float *__restrict__ out_d2, int *__restrict__ out_idx,
                                 const int *__restrict__ skip, long long skip_stride,
                                 long long per_batch) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (skip != nullptr && skip[(i / per_batch) * skip_stride] != 0) return;
    const u64 k = packed[i];
    out_d2[i] = __uint_as_float((unsigned)(k >> 32));
    if (out_idx != nullptr) out_idx[i] = (int)(unsigned)(k & 0xffffffffu);
}

// mean of sqrt(d2) per batch, FP64, one CTA per batch => fixed summation order.
constexpr int kMeanThreads = 1024;
__global__ void __launch_bounds__(kMeanThreads)
mean_sqrt_kernel(const float *__restrict__ d2, long long n, double *__restrict__ out) {
    __shared__ double part[kMeanThreads / 32];
    const float *row = d2 + (long long)blockIdx.x * n;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += kMeanThreads) acc += sqrt((double)row[i]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = part[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) out[blockIdx.x] = (n > 0) ? v / (double)n : 0.0;
    }
}

// ---- host-side launch ----------------------------------------------------------------
template <int Q, int THREADS, int STAGE, int NSTAGES, int SUB, int MINB, int UNR = 2>
struct NNVariant {
    static constexpr int kQueriesPerCta = Q * THREADS;
    static constexpr size_t kSmem = (size_t)NSTAGES * 3 * STAGE * 4 + NSTAGES * 8;

    template <bool WITH_IDX>
    static int launch(const NNParams &p, dim3 grid, cudaStream_t st) {
        auto kern = nn_kernel<Q, THREADS, STAGE, NSTAGES, SUB, WITH_IDX, MINB, UNR>;
        static thread_local int configured_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_dev != dev) {
            ISR_TRY(check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)kSmem),
                               "nn smem attr"));
            configured_dev = dev;
        }
        ProfScope prof(kProfNN, st);
        kern<<<grid, THREADS, kSmem, st>>>(p);
        return launched("nn_kernel");
    }

    static int ctas_per_sm() {
        int n = 0;
        auto kern = nn_kernel<Q, THREADS, STAGE, NSTAGES, SUB, true, MINB, UNR>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, THREADS, kSmem) != cudaSuccess ||
            n < 1)
            n = 1;
        return n;
    }
};

// The production shape: 8 queries/thread x 128 threads = 1024 queries per CTA
// (= ISR_SOA_TILE), 1024-target stages, 3 in flight (36 KB), 256-target sub-tiles.
using NNMain = NNVariant<8, 128, 1024, 3, 256, 4>;
static_assert(ISR_SOA_TILE % 1024 == 0, "stage must divide the SoA padding");

struct NNCall {
    const float *q_soa; int64_t nq, nq_pad, q_bstride;
    const float *t_soa; int64_t nt, nt_pad, t_bstride;
    int64_t batch;
    float *out_d2; int32_t *out_idx;
    const int32_t *skip; int64_t skip_stride;
    void *workspace; size_t workspace_bytes;
    cudaStream_t st;
};

template <class V, int STAGE>
static int nn_dispatch(const NNCall &c) {
    const int nqb = (int)((c.nq + V::kQueriesPerCta - 1) / V::kQueriesPerCta);
    const int stages = (int)(c.nt_pad / STAGE);
    // target split: fill the machine when there are few query blocks, and cut the tail
    // when the grid is only a few waves deep.
    static thread_local int slots = 0;
    if (slots == 0) slots = sm_count() * V::ctas_per_sm();
    int splits = 1;
    const long long ctas = (long long)nqb * c.batch;
    if (ctas < 6ll * slots) {
        long long want = (8ll * slots + ctas - 1) / ctas;
        int max_splits = stages / 4 > 0 ? stages / 4 : 1;  // >= 4 stages per split
        splits = (int)(want < max_splits ? want : max_splits);
        if (splits < 1) splits = 1;
    }
    int per = (stages + splits - 1) / splits;
    splits = (stages + per - 1) / per;  // no empty split
    ISR_REQUIRE(splits <= 65535, ISR_E_SHAPE, "nn: too many target splits");

    NNParams p;
    p.q = c.q_soa; p.q_bstride = c.q_bstride; p.nq = (int)c.nq; p.nq_pad = (int)c.nq_pad;
    p.t = c.t_soa; p.t_bstride = c.t_bstride; p.nt_pad = (int)c.nt_pad;
    p.out_d2 = c.out_d2; p.out_idx = c.out_idx; p.out_packed = nullptr;
    p.skip = c.skip; p.skip_stride = c.skip_stride;
    p.stages_total = stages; p.stages_per_split = per;
    dim3 grid((unsigned)nqb, (unsigned)splits, (unsigned)c.batch);

    if (splits == 1) {
        if (c.out_idx != nullptr) return V::template launch<true>(p, grid, c.st);
        return V::template launch<false>(p, grid, c.st);
    }
    const size_t need = (size_t)c.nq * (size_t)c.batch * sizeof(u64);
    ISR_REQUIRE(c.workspace != nullptr && c.workspace_bytes >= need, ISR_E_WORKSPACE,
                "nn: workspace %zu < %zu bytes", c.workspace_bytes, need);
    p.out_packed = reinterpret_cast<u64 *>(c.workspace);
    // skipped batches keep their previous outputs: the unpack kernel honours `skip` too
    ISR_TRY(check_cuda(cudaMemsetAsync(c.workspace, 0xff, need, c.st), "nn memset"));
    ISR_TRY(V::template launch<true>(p, grid, c.st));
    const long long total = (long long)c.nq * c.batch;
    nn_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.st>>>(
        p.out_packed, total, c.out_d2, c.out_idx, c.skip, c.skip_stride, (long long)c.nq);
    return launched("nn_unpack_kernel");
}

#ifdef ISR_NN_TUNING
// Developer-only shapes selected with ISR_NN_VARIANT (built with -DISR_NN_TUNING).
using NNT1 = NNVariant<8, 256, 1024, 3, 256, 2>;
using NNT2 = NNVariant<4, 128, 1024, 3, 256, 7>;
using NNT3 = NNVariant<4, 256, 1024, 3, 256, 3>;
using NNT4 = NNVariant<6, 128, 1024, 3, 256, 5>;
using NNT5 = NNVariant<8, 128, 1024, 3, 256, 4, 1>;
using NNT6 = NNVariant<8, 128, 1024, 3, 256, 4, 4>;
using NNT7 = NNVariant<8, 128, 2048, 3, 256, 3>;
using NNT8 = NNVariant<8, 128, 512, 4, 256, 4>;
using NNT9 = NNVariant<8, 64, 1024, 3, 256, 8>;
using NNT10 = NNVariant<4, 128, 1024, 3, 256, 8, 4>;
using NNT11 = NNVariant<6, 128, 1024, 3, 256, 5, 1>;
using NNT12 = NNVariant<8, 128, 1024, 3, 256, 5, 1>;
#endif

}  // namespace isr

extern "C" {

size_t isr_nn_workspace_bytes(int64_t nq, int64_t nt, int64_t batch) {
    (void)nt;
    if (nq <= 0 || batch <= 0) return 256;
    return isr::align256((size_t)nq * (size_t)batch * sizeof(unsigned long long));
}

int isr_nn_soa(const float *q_soa, int64_t nq, int64_t nq_pad, int64_t q_bstride,
               const float *t_soa, int64_t nt, int64_t nt_pad, int64_t t_bstride, int64_t batch,
               float *out_d2, int32_t *out_idx, const int32_t *skip, int64_t skip_stride,
               void *workspace, size_t workspace_bytes, void *stream) {
    using namespace isr;
    ISR_REQUIRE(nq >= 0 && nt >= 1 && batch >= 0, ISR_E_SHAPE,
                "nn: need nq >= 0, nt >= 1, batch >= 0 (nq=%lld nt=%lld batch=%lld)",
                (long long)nq, (long long)nt, (long long)batch);
    if (nq == 0 || batch == 0) return ISR_OK;
    ISR_REQUIRE(q_soa && t_soa && out_d2, ISR_E_INVALID_ARG, "nn: null pointer");
    ISR_REQUIRE(nq_pad >= nq && nq_pad % ISR_SOA_TILE == 0 && nt_pad >= nt &&
                    nt_pad % ISR_SOA_TILE == 0,
                ISR_E_SHAPE, "nn: padded lengths must be multiples of %d covering n", ISR_SOA_TILE);
    ISR_REQUIRE(nq_pad < (1ll << 31) - 2048 && nt_pad < (1ll << 31) - 2048, ISR_E_SHAPE,
                "nn: clouds beyond int32 indexing");
    ISR_REQUIRE(batch <= 65535, ISR_E_SHAPE, "nn: batch %lld > 65535", (long long)batch);
    ISR_REQUIRE(aligned16(t_soa) && (t_bstride % 4 == 0), ISR_E_ALIGN,
                "nn: target planes must be 16-byte aligned");
    const NNCall c{q_soa, nq, nq_pad, q_bstride, t_soa, nt, nt_pad, t_bstride, batch, out_d2, out_idx,
                   skip, skip_stride, workspace, workspace_bytes, (cudaStream_t)stream};
#ifdef ISR_NN_TUNING
    static int variant = -1;
    if (variant < 0) {
        const char *e = getenv("ISR_NN_VARIANT");
        variant = e ? atoi(e) : 0;
    }
    switch (variant) {
        case 1: return nn_dispatch<NNT1, 1024>(c);
        case 2: return nn_dispatch<NNT2, 1024>(c);
        case 3: return nn_dispatch<NNT3, 1024>(c);
        case 4: return nn_dispatch<NNT4, 1024>(c);
        case 5: return nn_dispatch<NNT5, 1024>(c);
        case 6: return nn_dispatch<NNT6, 1024>(c);
        case 7: return nn_dispatch<NNT7, 2048>(c);
        case 8: return nn_dispatch<NNT8, 512>(c);
        case 9: return nn_dispatch<NNT9, 1024>(c);
        case 10: return nn_dispatch<NNT10, 1024>(c);
        case 11: return nn_dispatch<NNT11, 1024>(c);
        case 12: return nn_dispatch<NNT12, 1024>(c);
        default: break;
    }
#endif
    return nn_dispatch<NNMain, 1024>(c);
}

int isr_mean_sqrt(const float *d2, int64_t n, int64_t batch, double *out_mean, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0 && batch >= 0, ISR_E_SHAPE, "mean_sqrt: negative size");
    if (batch == 0) return ISR_OK;
    ISR_REQUIRE(out_mean && (d2 || n == 0), ISR_E_INVALID_ARG, "mean_sqrt: null pointer");
    ProfScope prof(kProfReduce, (cudaStream_t)stream);
    mean_sqrt_kernel<<<(unsigned)batch, kMeanThreads, 0, (cudaStream_t)stream>>>(d2, n, out_mean);
    return launched("mean_sqrt_kernel");
}

}  // extern "C"
