// Library-level entry points: error string, ABI version, device probe.
#include <stdarg.h>

#include "common.cuh"

namespace tdr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace tdr

extern "C" TDR_API int tdr_abi_version(void) { return 2; }  // 2: per-call kNN options, persistent run entry points, row-local LargeVis step

extern "C" TDR_API const char* tdr_last_error(void) { return tdr::g_err; }

extern "C" TDR_API int tdr_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    TDR_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    TDR_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (p.major != 10) {
        tdr::set_error("libtdrb200 is built for sm_100a only; device is sm_%d%d", p.major, p.minor);
        return TDR_E_UNSUPPORTED;
    }
    return TDR_OK;
}
