// Error reporting, device checks and the driver entry point used for TMA descriptors.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ebsd {

static thread_local char g_error[512] = "";
static unsigned long long g_launches = 0;

void note_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_device_arch() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s", cudaGetErrorString(e));
        return EBSD_ERR_CUDA;
    }
    static thread_local int cached_dev = -1, cached_major = 0;
    if (cached_dev != dev) {
        int major = 0;
        e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        if (e != cudaSuccess) {
            set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
            return EBSD_ERR_CUDA;
        }
        cached_dev = dev;
        cached_major = major;
    }
    if (cached_major != 10) {
        set_error("libebsd_b200 is built for sm_100a only; device %d has compute capability major %d", dev,
                  cached_major);
        return EBSD_ERR_ARCH;
    }
    return EBSD_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached;
    if (dev != cached_dev) {
        int n = 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

tensormap_encode_fn get_tensormap_encode() {
    static tensormap_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (tensormap_encode_fn)p;
    }
    return fn;
}

}  // namespace ebsd

extern "C" {
int ebsd_abi_version(void) { return 1; }
const char *ebsd_last_error(void) { return ebsd::g_error; }
uint64_t ebsd_launch_count(void) { return __atomic_load_n(&ebsd::g_launches, __ATOMIC_RELAXED); }
}
