// Library-level entry points: version, error string, device info, launch counter.
#include <stdarg.h>
#include <string.h>

#include "isr_common.cuh"

namespace isr {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace isr

extern "C" {

int isr_version(void) { return ISR_VERSION; }

const char *isr_last_error(void) { return isr::t_err; }

int isr_device_info(int *sm_count, int *sm_clock_khz, int *smem_per_sm) {
    int dev = 0;
    ISR_TRY(isr::check_cuda(cudaGetDevice(&dev), "cudaGetDevice"));
    int v = 0;
    if (sm_count) {
        ISR_TRY(isr::check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev),
                                "attr sm count"));
        *sm_count = v;
    }
    if (sm_clock_khz) {
        ISR_TRY(isr::check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev),
                                "attr clock"));
        *sm_clock_khz = v;
    }
    if (smem_per_sm) {
        ISR_TRY(isr::check_cuda(
            cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev),
            "attr smem"));
        *smem_per_sm = v;
    }
    return ISR_OK;
}

uint64_t isr_launch_count(void) { return isr::g_launches.load(); }
void isr_reset_launch_count(void) { isr::g_launches.store(0); }

}  // extern "C"
