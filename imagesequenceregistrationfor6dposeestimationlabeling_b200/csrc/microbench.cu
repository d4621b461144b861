// FFMA-chain microbenchmark: the measured FP32 CUDA-core peak that bench.py reports
// next to the nominal 148 SM x 128 lanes x 2 flop x clock.
#include "isr_common.cuh"

namespace isr {

template <bool PACKED>
__global__ void __launch_bounds__(256) ffma_chain_kernel(int iters, float *sink) {
    const float b = 1.0000001f, c = 1.0e-9f;
    if (PACKED) {
        u64 a[16];
        const u64 bb = pack2(b, b), cc = pack2(c, c);
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = pack2(1.0f + threadIdx.x * 1e-6f + k, 2.0f + k);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = fma2(a[k], bb, cc);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) { float lo, hi; unpack2(a[k], lo, hi); s += lo + hi; }
        if (s == 123.456f) sink[0] = s;
    } else {
        float a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = 1.0f + threadIdx.x * 1e-6f + k;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = __fmaf_rn(a[k], b, c);
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += a[k];
        if (s == 123.456f) sink[0] = s;
    }
}

}  // namespace isr

extern "C" int isr_bench_ffma(int blocks, int iters, int packed, float *sink, double *flops_out_host,
                              void *stream) {
    using namespace isr;
    ISR_REQUIRE(blocks > 0 && iters > 0 && sink != nullptr, ISR_E_INVALID_ARG, "bench_ffma: bad arg");
    if (packed)
        ffma_chain_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    else
        ffma_chain_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    if (flops_out_host)
        *flops_out_host = (double)blocks * 256.0 * (double)iters * 16.0 * (packed ? 4.0 : 2.0);
    return launched("ffma_chain_kernel");
}
