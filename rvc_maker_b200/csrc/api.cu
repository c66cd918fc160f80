// ABI bookkeeping entry points.
#include "common.cuh"

namespace rvc {
char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace rvc

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
