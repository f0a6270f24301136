// Library-wide state of libtvs_b200.so: error message, launch counter, device checks.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include <stdlib.h>

#include "common.cuh"
#include "tvs_b200.h"

namespace tvs {

static thread_local char g_error[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// Called right after every kernel launch: counts it and surfaces launch-configuration errors
// (asynchronous faults show up at the caller's next synchronisation, as with any CUDA library).
int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return -3;
    }
    return 0;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// Programmatic dependent launch, see common.cuh.  Off unless TVS_PDL=1: measured on the MaPLe step (B200, one CUDA
// graph of ~640 kernel nodes) it changes nothing - 11.41 / 11.29 ms with, 11.35 / 11.32 ms without - so the inter-kernel
// gap is not what the step is waiting for; the plumbing stays for kernels with longer prologues.
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("TVS_PDL"); return e && (e[0] == '1' || e[0] == '2'); }();
    return on;
}
// (round 2, persistent attention kernels in place: TVS_PDL=1 9.63 vs 9.35 ms, TVS_PDL=2 9.44-9.56 vs 9.43 ms - still no gain.)
// TVS_PDL=2: only launches of at most this many CTAs (the text tower's and the decoder's small kernels - a persistent GEMM has one
// CTA per SM) are launched programmatically; 0 = no limit (TVS_PDL=1)
int pdl_max_ctas() {
    static const int n = [] { const char* e = getenv("TVS_PDL"); return (e && e[0] == '2') ? 128 : 0; }();
    return n;
}

}  // namespace tvs

extern "C" __attribute__((visibility("default"))) int tvs_version(void) { return TVS_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) const char* tvs_last_error(void) { return tvs::g_error; }
extern "C" __attribute__((visibility("default"))) int64_t tvs_launch_count(void) { return tvs::g_launches.load(std::memory_order_relaxed); }

extern "C" __attribute__((visibility("default"))) int tvs_device_check(void) {
    int dev = 0, major = 0, minor = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        tvs::set_error("no CUDA device: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return -2;
    }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        tvs::set_error("libtvs_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
        return -1;
    }
    return 0;
}
