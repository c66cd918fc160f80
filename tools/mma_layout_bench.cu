// Microbenchmark: tcgen05.mma (M=128, K=16, fp16) cycles per instruction for different shared-memory operand layouts.
// Operands are whatever is in shared memory (values do not matter for timing).   nvcc -arch=sm_100a -O3 -o mma_bench tools/mma_layout_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../rvc_maker_b200/csrc/tc_common.cuh"

__device__ __forceinline__ uint64_t desc_sw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    const uint32_t lo = ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    return ((uint64_t)hi << 32) | lo;
}

template <int NT>
__global__ void __launch_bounds__(128, 1) bench(int variant, int iters, int rows_a, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&slot), 512);
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x == 32) {
        const uint32_t a0 = tc::smem_u32(smem), b0 = a0 + 96 * 1024;
        const uint32_t idesc = tc::make_idesc(0u, NT);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int ks = i & 7;   // 8 k-steps per "tap"
            const int tap = (i >> 3) % 5;
            uint64_t da, db;
            if (variant == 0) {        // interleaved, rows 16 B apart (shift-friendly): LBO = rows*16, SBO = 128
                da = desc_sw(a0 + (2 * ks * rows_a + tap * 3) * 16, rows_a * 16, 128, 0);
                db = desc_sw(b0 + (2 * ks * NT) * 16, NT * 16, 128, 0);
            } else if (variant == 1) { // interleaved canonical dense: K chunks adjacent (LBO = 128), SBO = 256
                da = desc_sw(a0 + ks * 128 * 32, 128, 256, 0);
                db = desc_sw(b0 + ks * NT * 32, 128, 256, 0);
            } else if (variant == 2) { // 128B swizzle, rows of 128 B: SBO = 1024, k-step = +32 B, row shift = +128 B * tap
                da = desc_sw(a0 + ks * 32 + tap * 3 * 128, 16, 1024, 2);
                db = desc_sw(b0 + ks * 32, 16, 1024, 2);
            } else {                   // variant 0 for A, but accumulators alternate between 4 TMEM tiles
                da = desc_sw(a0 + (2 * ks * rows_a + tap * 3) * 16, rows_a * 16, 128, 0);
                db = desc_sw(b0 + (2 * ks * NT) * 16, NT * 16, 128, 0);
            }
            const uint32_t d = tmem + ((variant == 3) ? (uint32_t)((i & 3) * NT) : 0u);
            tc::umma<false>(d, da, db, idesc, 1u);
        }
        tc::umma_commit(tc::smem_u32(&bar));
        tc::mbar_wait(tc::smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) { tc::fence_after_sync(); tc::tmem_dealloc(tmem, 512); }
}

template <int NT> void run(long long* d_out) {
    cudaFuncSetAttribute(bench<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[4] = {"interleave rows@16B (current)", "interleave dense (LBO=128)", "swizzle-128B rows@128B", "current, 4 accumulators"};
    for (int v = 0; v < 4; ++v) {
        for (int grid : {1, 148}) {
            long long h = 0;
            const int iters = 4000;
            bench<NT><<<grid, 128, 200 * 1024>>>(v, iters, 306, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
            printf("N=%3d  %-32s grid=%3d  %7.1f cycles/MMA  (%s)\n", NT, names[v], grid, (double)h / iters, cudaGetErrorString(e));
        }
    }
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 8);
    run<32>(d_out); run<64>(d_out); run<128>(d_out); run<256>(d_out);
    return 0;
}
