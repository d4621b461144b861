// Shared host/device helpers for libisr (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "isr.h"

namespace isr {

// ---- error reporting (thread-local message, status codes; no exceptions) -------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return ISR_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ISR_E_CUDA;
}

// Call right after a kernel launch: counts it and surfaces launch errors.
inline int launched(const char *name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return check_cuda(cudaGetLastError(), name);
}

// Optional per-kernel device timing (isr_profile_*): when enabled, a launch site brackets
// its kernel with a pair of CUDA events on the launching stream.
enum ProfKind { kProfTransform = 0, kProfNN = 1, kProfReduce = 2, kProfIcpAcc = 3, kProfIcpSolve = 4,
                kProfKinds = 5 };
bool prof_enabled();
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);
struct ProfScope {
    int kind;
    cudaStream_t st;
    bool on;
    ProfScope(int k, cudaStream_t s) : kind(k), st(s), on(prof_enabled()) {
        if (on) prof_begin(kind, st);
    }
    ~ProfScope() {
        if (on) prof_end(kind, st);
    }
};

#define ISR_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != ISR_OK) return _s;  \
    } while (0)

#define ISR_REQUIRE(cond, code, ...)  \
    do {                              \
        if (!(cond)) {                \
            set_error(__VA_ARGS__);   \
            return (code);            \
        }                             \
    } while (0)

// isr_nn2 with one more knob for callers that repeat the same search on slowly moving clouds
// (the ICP loop): reuse_order != 0 keeps the launch order of the pruned kernel's query blocks
// that the previous call left in `workspace` (block weights are invariant under the rigid
// motion between two iterations) instead of recomputing it.
// `fuse` (icp_device.cuh), when not NULL, turns the search into one whole ICP evaluation + update:
// q->soa7 then holds the ORIGINAL source planes (the kernel applies the pose itself) and nothing
// is written to out_d2 / out_idx.  nn2_fusable: the pruned search would be used for this target.
struct IcpFuse;
int nn2_search(const IsrCloud *q, const IsrCloud *t, int64_t batch, int use_lo, float *out_d2,
               int32_t *out_idx, const int32_t *skip, int64_t skip_stride, void *workspace,
               size_t workspace_bytes, void *stream, int reuse_order, const IcpFuse *fuse);
bool nn2_fusable(const IsrCloud *t);
int nn2_query_blocks(int64_t nq);  // 256-query blocks of the pruned search

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int sm_count();

// ---- device helpers ---------------------------------------------------------------
#ifdef __CUDACC__
typedef unsigned long long u64;

// Packed FP32 pairs: one FADD2/FMUL2/FFMA2 issue slot carries two IEEE-rn operations.
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// three-input minimum (FMNMX3 on sm_100)
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// mbarrier + 1-D bulk async copy (TMA engine, UBLKCP in SASS)
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// programmatic dependent launch: a grid launched with the stream-serialization attribute may start
// while its predecessor in the stream still runs; pdl_wait() returns once that grid has completed
// and its writes are visible (at once for a normally launched grid), pdl_launch_dependents() lets
// the successor's CTAs take the slots this grid frees
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif  // __CUDACC__

}  // namespace isr
