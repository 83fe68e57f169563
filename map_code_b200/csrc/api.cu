// Error reporting + misc entry points of the C ABI (include/map_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mapb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MAP_ECUDA;
    }
    return MAP_OK;
}
__global__ void timestamp_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    *slot = t;
}
}  // namespace mapb

extern "C" int map_timestamp_ns(unsigned long long* slot, map_stream_t stream) {
    MAP_REQUIRE(slot != nullptr, "map_timestamp_ns: null pointer");
    mapb::timestamp_kernel<<<1, 1, 0, mapb::as_stream(stream)>>>(slot);
    return mapb::check_launch("map_timestamp_ns");
}

extern "C" int map_abi_version(void) { return MAP_B200_ABI_VERSION; }
extern "C" const char* map_last_error(void) { return mapb::g_err; }
extern "C" int map_sm_count(int* out) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        mapb::set_error("map_sm_count: no CUDA device");
        cudaGetLastError();
        return MAP_ECUDA;
    }
    *out = n;
    return MAP_OK;
}
