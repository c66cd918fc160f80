// ABI bookkeeping entry points.
#include "common.cuh"

namespace rvc {
char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
static int g_grid_limit = 0;
int grid_limit() { return g_grid_limit; }
}  // namespace rvc

// Persistent-kernel grid cap (0 = one CTA per SM): rvc_conv1d_tc and rvc_resblock_pair_tc launch at most `n` CTAs while it is set.  The engine sets
// it around the launches (or the graph capture) of the generator when the encoder of the NEXT chunk runs next to it on a second stream, so that the
// generator's persistent CTAs leave SMs to the latency-bound short-sequence kernels (rvc_maker_b200/engine.py, overlap mode).  Process-wide, not thread-safe:
// set it from the thread that issues the launches.  Returns the previous value.
RVC_API int rvc_set_grid_limit(int n) {
    const int prev = rvc::g_grid_limit;
    rvc::g_grid_limit = n > 0 ? n : 0;
    return prev;
}

RVC_API int rvc_abi_version(void) { return RVC_ABI_VERSION; }
RVC_API const char* rvc_last_error(void) { return rvc::last_error_buf(); }

RVC_API int rvc_device_info(int device, int* out4) {
    if (!out4) return rvc::fail(RVC_EINVAL, "rvc_device_info: null output");
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return rvc::fail((int)e, "rvc_device_info: %s", cudaGetErrorString(e));
    out4[0] = p.multiProcessorCount;
    out4[1] = p.major;
    out4[2] = p.minor;
    out4[3] = (int)p.sharedMemPerBlockOptin;
    return 0;
}
