// Library-level entry points: version, error string, device info, launch counter.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <utility>
#include <vector>

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

// ---- per-kernel device timing ---------------------------------------------------------
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events[kProfKinds];
static thread_local cudaEvent_t t_prof_open[kProfKinds];

bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }

void prof_begin(int kind, cudaStream_t st) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) { t_prof_open[kind] = nullptr; return; }
    cudaEventRecord(e, st);
    t_prof_open[kind] = e;
}

void prof_end(int kind, cudaStream_t st) {
    cudaEvent_t b = t_prof_open[kind], e = nullptr;
    if (b == nullptr) return;
    if (cudaEventCreate(&e) != cudaSuccess) { cudaEventDestroy(b); return; }
    cudaEventRecord(e, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_events[kind].push_back({b, e});
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

int isr_profile_enable(int on) {
    isr::g_prof_on.store(on != 0);
    return ISR_OK;
}

int isr_profile_collect(double *ms_by_kind, uint64_t *launches_by_kind) {
    using namespace isr;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int status = ISR_OK;
    for (int k = 0; k < kProfKinds; ++k) {
        double total = 0.0;
        for (auto &pr : g_prof_events[k]) {
            float ms = 0.f;
            cudaError_t e = cudaEventSynchronize(pr.second);
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, pr.first, pr.second);
            if (e != cudaSuccess) status = check_cuda(e, "profile_collect");
            total += ms;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        if (ms_by_kind) ms_by_kind[k] = total;
        if (launches_by_kind) launches_by_kind[k] = g_prof_events[k].size();
        g_prof_events[k].clear();
    }
    return status;
}

uint64_t isr_launch_count(void) { return isr::g_launches.load(); }
void isr_reset_launch_count(void) { isr::g_launches.store(0); }

}  // extern "C"
