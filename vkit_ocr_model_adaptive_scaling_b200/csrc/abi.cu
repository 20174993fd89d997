// C-ABI runtime basics: version, thread-local error text, device check.  See include/vkocr_b200.h.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

#define VKOCR_ABI_VERSION 1

static thread_local char g_err[512] = "";

void vkocr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void vkocr_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int vkocr_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}

extern "C" {

int vkocr_abi_version(void) { return VKOCR_ABI_VERSION; }

const char* vkocr_last_error(void) { return g_err; }

// Number of kernels this library has launched in this process so far (bench.py reports the per-step delta).
long long vkocr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// 0 when `device` is a compute-capability 10.x GPU (the only target of this library), negative otherwise.
int vkocr_device_check(int device) {
    int major = 0, minor = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (e != cudaSuccess) VK_FAIL(VKOCR_CUDA_ERROR, "device_check: %s", cudaGetErrorString(e));
    if (major != 10) VK_FAIL(VKOCR_CUDA_ERROR, "device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
    return VKOCR_OK;
}

}  // extern "C"
