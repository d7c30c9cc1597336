// Library-level pieces of the C-ABI: version, error string, device queries.
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace leida {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct DevInfo {
    int sms = 0;
    int smem_optin = 0;
};

static DevInfo dev_info() {
    static thread_local int cached_dev = -1;
    static thread_local DevInfo info;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return DevInfo{148, 227 * 1024};
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&info.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&info.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cached_dev = dev;
    }
    return info;
}

int sm_count() { int s = dev_info().sms; return s > 0 ? s : 148; }
int max_smem_optin() { int s = dev_info().smem_optin; return s > 0 ? s : 227 * 1024; }

}  // namespace leida

extern "C" int leida_version(void) { return 100; }
extern "C" const char* leida_last_error(void) { return leida::g_err; }
extern "C" int leida_device_sm_count(void) { return leida::sm_count(); }
extern "C" int64_t leida_launch_count(void) { return leida::g_launches.load(); }
