// Spatial ordering of a cloud (Hilbert curve) -- preprocessing for the filtered
// nearest-neighbour kernel (nn2.cu), once per cloud.
//
// The brute-force scan visits every pair whatever the order, but how often its rare
// exact-resolve path fires depends on the order: when consecutive points are spatial
// neighbours, the queries of a warp improve their running minimum in the same few target
// sub-tiles and stay silent everywhere else.  NeRF surface clouds come out of farthest-point
// sampling (genFeat.py:199-202), i.e. in effectively random order, so the library orders
// them itself.  Results are always reported in the caller's original indexing.
//
//   bbox_kernel        min/max of the cloud (single CTA)
//   morton_keys_kernel key = (30-bit Hilbert index of the cubic-cell quantised point) << 32 | index
//   bitonic_*          sort of the 64-bit keys (unique => deterministic, stable in the index)
//   extract_perm       perm[i] = original index of the i-th point along the curve
#include <math_constants.h>

#include "isr_common.cuh"

namespace isr {

__global__ void __launch_bounds__(1024)
bbox_kernel(const float *__restrict__ pts, int64_t n, float *__restrict__ out6) {
    __shared__ float red[6][32];
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
    float mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = pts[3 * i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if ((threadIdx.x & 31) == 0) {
            red[c][threadIdx.x >> 5] = mn[c];
            red[3 + c][threadIdx.x >> 5] = mx[c];
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = red[c][threadIdx.x], b = red[3 + c][threadIdx.x];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
                b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
            }
            if (threadIdx.x == 0) {
                out6[c] = a;
                out6[3 + c] = b;
            }
        }
    }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void morton_keys_kernel(const float *__restrict__ pts, int64_t n, int64_t npow2,
                                   const float *__restrict__ bbox, u64 *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npow2) return;
    if (i >= n) {
        keys[i] = ~0ull;  // padding sorts last
        return;
    }
    const float ext = fmaxf(fmaxf(bbox[3] - bbox[0], bbox[4] - bbox[1]), fmaxf(bbox[5] - bbox[2], 1e-30f));
    const float sc = 1023.0f / ext;
    unsigned q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = (pts[3 * i + c] - bbox[c]) * sc;
        q[c] = (unsigned)fminf(fmaxf(v, 0.0f), 1023.0f);
    }
    // Hilbert index of the cell (Skilling's transpose form, 10 bits per axis): unlike the
    // Z-curve it never jumps, so every run of consecutive points -- a query row, a sub-tile, a
    // stage -- is one connected patch and its bounding sphere stays tight.
    constexpr unsigned M = 1u << 9;
    for (unsigned Q = M; Q > 1; Q >>= 1) {
        const unsigned P = Q - 1;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (q[c] & Q) {
                q[0] ^= P;
            } else {
                const unsigned t = (q[0] ^ q[c]) & P;
                q[0] ^= t;
                q[c] ^= t;
            }
        }
    }
    q[1] ^= q[0];
    q[2] ^= q[1];
    unsigned t = 0;
    for (unsigned Q = M; Q > 1; Q >>= 1)
        if (q[2] & Q) t ^= Q - 1;
    q[0] ^= t; q[1] ^= t; q[2] ^= t;
    const unsigned code = (spread10(q[0]) << 2) | (spread10(q[1]) << 1) | spread10(q[2]);
    keys[i] = ((u64)code << 32) | (u64)(unsigned)i;
}

constexpr int kSortBlock = 2048;  // elements per CTA in the shared-memory phases

__device__ __forceinline__ void cmpx(u64 &a, u64 &b, bool ascending) {
    if ((a > b) == ascending) {
        const u64 t = a;
        a = b;
        b = t;
    }
}

// sorts each 2048-element block completely (all k <= 2048), direction by global index
__global__ void __launch_bounds__(kSortBlock / 2) bitonic_local_sort_kernel(u64 *__restrict__ keys) {
    __shared__ u64 s[kSortBlock];
    const int64_t base = (int64_t)blockIdx.x * kSortBlock;
    const int t = threadIdx.x;
    s[t] = keys[base + t];
    s[t + kSortBlock / 2] = keys[base + t + kSortBlock / 2];
    __syncthreads();
    for (int k = 2; k <= kSortBlock; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int i = 2 * t - (t & (j - 1));  // lower index of the pair
            const bool asc = (((base + i) & k) == 0);
            cmpx(s[i], s[i + j], asc);
            __syncthreads();
        }
    }
    keys[base + t] = s[t];
    keys[base + t + kSortBlock / 2] = s[t + kSortBlock / 2];
}

// one compare-exchange step with distance j >= kSortBlock
__global__ void bitonic_global_step_kernel(u64 *__restrict__ keys, int64_t half, int64_t k, int64_t j) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= half) return;
    const int64_t i = 2 * t - (t & (j - 1));
    const bool asc = ((i & k) == 0);
    u64 a = keys[i], b = keys[i + j];
    if ((a > b) == asc) {
        keys[i] = b;
        keys[i + j] = a;
    }
}

// finishes phase k: all steps j = kSortBlock/2 .. 1 inside shared memory
__global__ void __launch_bounds__(kSortBlock / 2)
bitonic_local_merge_kernel(u64 *__restrict__ keys, int64_t k) {
    __shared__ u64 s[kSortBlock];
    const int64_t base = (int64_t)blockIdx.x * kSortBlock;
    const int t = threadIdx.x;
    s[t] = keys[base + t];
    s[t + kSortBlock / 2] = keys[base + t + kSortBlock / 2];
    __syncthreads();
    for (int j = kSortBlock >> 1; j > 0; j >>= 1) {
        const int i = 2 * t - (t & (j - 1));
        const bool asc = (((base + i) & k) == 0);
        cmpx(s[i], s[i + j], asc);
        __syncthreads();
    }
    keys[base + t] = s[t];
    keys[base + t + kSortBlock / 2] = s[t + kSortBlock / 2];
}

__global__ void extract_perm_kernel(const u64 *__restrict__ keys, int64_t n, int32_t *__restrict__ perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) perm[i] = (int32_t)(unsigned)(keys[i] & 0xffffffffull);
}

static int64_t pow2_at_least(int64_t n, int64_t floor_) {
    int64_t p = floor_;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace isr

extern "C" {

size_t isr_spatial_order_workspace_bytes(int64_t n) {
    using namespace isr;
    if (n <= 0) return 256;
    return align256((size_t)pow2_at_least(n, kSortBlock) * 8) + 256;
}

int isr_spatial_order(const float *pts, int64_t n, int32_t *perm, void *workspace,
                      size_t workspace_bytes, void *stream) {
    using namespace isr;
    ISR_REQUIRE(n >= 0, ISR_E_SHAPE, "spatial_order: negative size");
    if (n == 0) return ISR_OK;
    ISR_REQUIRE(pts && perm, ISR_E_INVALID_ARG, "spatial_order: null pointer");
    ISR_REQUIRE(n < (1ll << 31) - 2048, ISR_E_SHAPE, "spatial_order: cloud beyond int32 indexing");
    const int64_t np2 = pow2_at_least(n, kSortBlock);
    const size_t need = align256((size_t)np2 * 8) + 256;
    ISR_REQUIRE(workspace != nullptr && workspace_bytes >= need, ISR_E_WORKSPACE,
                "spatial_order: workspace %zu < %zu bytes", workspace_bytes, need);
    ISR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, ISR_E_ALIGN,
                "spatial_order: workspace not 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    u64 *keys = reinterpret_cast<u64 *>(workspace);
    float *bbox = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + align256((size_t)np2 * 8));

    bbox_kernel<<<1, 1024, 0, st>>>(pts, n, bbox);
    ISR_TRY(launched("bbox_kernel"));
    morton_keys_kernel<<<(unsigned)((np2 + 255) / 256), 256, 0, st>>>(pts, n, np2, bbox, keys);
    ISR_TRY(launched("morton_keys_kernel"));
    const unsigned nblk = (unsigned)(np2 / kSortBlock);
    bitonic_local_sort_kernel<<<nblk, kSortBlock / 2, 0, st>>>(keys);
    ISR_TRY(launched("bitonic_local_sort_kernel"));
    const int64_t half = np2 / 2;
    for (int64_t k = 2 * kSortBlock; k <= np2; k <<= 1) {
        for (int64_t j = k >> 1; j >= kSortBlock; j >>= 1) {
            bitonic_global_step_kernel<<<(unsigned)((half + 255) / 256), 256, 0, st>>>(keys, half, k, j);
            ISR_TRY(launched("bitonic_global_step_kernel"));
        }
        bitonic_local_merge_kernel<<<nblk, kSortBlock / 2, 0, st>>>(keys, k);
        ISR_TRY(launched("bitonic_local_merge_kernel"));
    }
    extract_perm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, n, perm);
    return launched("extract_perm_kernel");
}

}  // extern "C"
