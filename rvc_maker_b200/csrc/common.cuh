// Shared helpers for librvc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/rvc_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librvc_b200 is written for sm_100a only"
#endif

namespace rvc {

char* last_error_buf();  // thread-local, 512 bytes (api.cu)
int grid_limit();        // rvc_set_grid_limit (api.cu): cap on the CTAs of a persistent launch, 0 = none
#define RVC_API extern "C" __attribute__((visibility("default")))

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.f / (1.f + expf(-v)); }

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch: kernels call pdl_prologue() first (lets the next kernel in the stream be scheduled early, then
// waits until every prerequisite grid has completed and its writes are visible); hosts launch through launch_pdl().
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // measured on B200 (bench.py, same box): PDL on these small / many-block kernels is 3 % SLOWER end to end (their early-scheduled
    // blocks occupy SMs while they wait), so it is off unless RVC_B200_PDL=1; the tcgen05 conv kernel always uses it (+3 %).
    static const bool enabled = [] { const char* e = getenv("RVC_B200_PDL"); return e && e[0] == '1'; }();
    cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace rvc
