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

// FFMA2 operand-pattern probes (developer tuning): 16 accumulator pairs per thread.
//  mode 2: acc = fma2(S_k.F32 scalar, bb, acc)   -- scalar-broadcast A operand (the scan's form)
//  mode 3: acc = fma2(S_k pair, bb, acc)         -- distinct A pairs
//  mode 4: mode 2 plus one FMNMX3 per two FFMA2  -- the scan's instruction mix
template <int MODE>
__global__ void __launch_bounds__(256) ffma2_pattern_kernel(int iters, float *sink, const float *src) {
    u64 a[16];
    float sc[16];
    u64 sp[16];
    const u64 bb = pack2(src[0], src[1]);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        a[k] = pack2(1.0f + threadIdx.x * 1e-6f + k, 2.0f + k);
        sc[k] = src[2 + k];
        sp[k] = pack2(src[2 + k], src[3 + k]);
    }
    float m0 = 1e30f, m1 = 1e30f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == 3) a[k] = fma2(sp[k], bb, a[k]);
            else a[k] = fma2(pack2(sc[k], sc[k]), bb, a[k]);
        }
        if (MODE == 4) {
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                float lo, hi;
                unpack2(a[k], lo, hi);
                m0 = min3(m0, lo, hi);
                unpack2(a[k + 1], lo, hi);
                m1 = min3(m1, lo, hi);
            }
        }
        if (MODE == 5) {  // two 2-input FMNMX per pair
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                float lo, hi;
                unpack2(a[k], lo, hi);
                m0 = fminf(m0, lo); m1 = fminf(m1, hi);
                unpack2(a[k + 1], lo, hi);
                m0 = fminf(m0, lo); m1 = fminf(m1, hi);
            }
        }
        if (MODE == 6) {  // three-input signed integer min on the bit patterns
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float lo, hi;
                unpack2(a[k], lo, hi);
                m0 = __int_as_float(min(min(__float_as_int(m0), __float_as_int(lo)), __float_as_int(hi)));
            }
        }
        if (MODE == 7) {  // LOP3: OR-accumulate both halves (sign-bit flagging)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float lo, hi;
                unpack2(a[k], lo, hi);
                m0 = __int_as_float(__float_as_int(m0) | __float_as_int(lo) | __float_as_int(hi));
            }
        }
        if (MODE == 9) {  // 8 extra scalar FFMA next to 16 FFMA2: does fmalite add capacity?
#pragma unroll
            for (int k = 0; k < 8; ++k) sc[k] = __fmaf_rn(sc[k], 1.0000001f, 1e-9f);
        }
        if (MODE == 8) {  // one FMNMX3 per FOUR values... (half the mins) reference point
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                float lo, hi;
                unpack2(a[k], lo, hi);
                m0 = min3(m0, lo, hi);
            }
        }
    }
    float s = m0 + m1;
#pragma unroll
    for (int k = 0; k < 16; ++k) { float lo, hi; unpack2(a[k], lo, hi); s += lo + hi + sc[k]; }
    if (s == 123.456f) sink[0] = s;
}

}  // namespace isr

extern "C" int isr_bench_ffma2_pattern(int blocks, int iters, int mode, float *sink, const float *src,
                                       double *flops_out_host, void *stream) {
    using namespace isr;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 2) ffma2_pattern_kernel<2><<<blocks, 256, 0, st>>>(iters, sink, src);
    else if (mode == 3) ffma2_pattern_kernel<3><<<blocks, 256, 0, st>>>(iters, sink, src);
    else if (mode == 5) ffma2_pattern_kernel<5><<<blocks, 256, 0, st>>>(iters, sink, src);
    else if (mode == 6) ffma2_pattern_kernel<6><<<blocks, 256, 0, st>>>(iters, sink, src);
    else if (mode == 7) ffma2_pattern_kernel<7><<<blocks, 256, 0, st>>>(iters, sink, src);
    else if (mode == 8) ffma2_pattern_kernel<8><<<blocks, 256, 0, st>>>(iters, sink, src);
    else if (mode == 9) ffma2_pattern_kernel<9><<<blocks, 256, 0, st>>>(iters, sink, src);
    else ffma2_pattern_kernel<4><<<blocks, 256, 0, st>>>(iters, sink, src);
    if (flops_out_host) *flops_out_host = (double)blocks * 256.0 * (double)iters * 16.0 * 4.0;
    return launched("ffma2_pattern_kernel");
}

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
