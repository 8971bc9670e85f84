// libpepper_b200: error state, version, device probe.
#include "common.cuh"
#include <stdarg.h>

namespace pb {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace pb

extern "C" const char *pb_last_error(void) { return pb::g_err; }
extern "C" int pb_version(void) { return 100; }
extern "C" int pb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
