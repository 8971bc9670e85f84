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

// device -> host copy of a library-owned buffer (tests / read-backs of views such as pb_bam_fetch_device's)
extern "C" int pb_memcpy_to_host(void *h_dst, const void *d_src, int64_t bytes) {
    if (bytes < 0 || (bytes && (!h_dst || !d_src))) { pb::set_error("pb_memcpy_to_host: bad argument"); return PB_ERR_ARG; }
    if (bytes) PB_CUDA(cudaMemcpy(h_dst, d_src, (size_t) bytes, cudaMemcpyDeviceToHost));
    return PB_OK;
}
