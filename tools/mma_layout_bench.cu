// Microbenchmark: tcgen05.mma (K=16, fp16, SS mode) cycles per instruction vs. N, M, number of issuing warps and operand layout.
// Operands are whatever is in shared memory (values do not matter for timing).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_layout_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../rvc_maker_b200/csrc/tc_common.cuh"

__host__ __device__ constexpr uint32_t idesc_mn(uint32_t m, uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24); }

// variant 0: interleaved rows@16B (LBO = rows*16, SBO = 128);  1: 128B swizzle (SBO = 1024)
__global__ void __launch_bounds__(128, 1) bench(int variant, int M, int N, int nissue, int iters, long long* out, int commit_every = 0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint64_t cbar[8];
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) {
        tc::mbar_init(tc::smem_u32(&bar[0]), 1); tc::mbar_init(tc::smem_u32(&bar[1]), 1);
        for (int i = 0; i < 8; ++i) tc::mbar_init(tc::smem_u32(&cbar[i]), 1u << 19);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&slot), 512);
    tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
    const uint32_t tmem = slot;
    if (warp >= 1 && warp <= nissue) {
        const int w = warp - 1;
        const uint32_t a16 = tc::smem_u32(smem) >> 4, b16 = (tc::smem_u32(smem) + 96 * 1024) >> 4;
        const uint32_t idesc = idesc_mn(M, N);
        const uint32_t hi = variant == 0 ? ((128u >> 4) | (1u << 14)) : ((1024u >> 4) | (1u << 14) | (2u << 29));
        const uint32_t lo_hi = variant == 0 ? (306u << 16) : (1u << 16);
        const uint32_t d = tmem + (uint32_t)(w * 256);
        long long t0 = clock64();
        uint32_t off = 0;
        for (int i = 0; i < iters; ++i) {
            const uint64_t da = ((uint64_t)hi << 32) | (lo_hi | ((a16 + off) & 0x3FFFu));
            const uint64_t db = ((uint64_t)hi << 32) | (lo_hi | ((b16 + off) & 0x3FFFu));
            if (tc::elect_one()) tc::umma<false>(d, da, db, idesc, 1u);
            off = (off + 2) & 63;
            if (commit_every && (i % commit_every) == commit_every - 1 && tc::elect_one()) tc::umma_commit(tc::smem_u32(&cbar[(i / commit_every) & 7]));
        }
        if (tc::elect_one()) tc::umma_commit(tc::smem_u32(&bar[w]));
        tc::mbar_wait(tc::smem_u32(&bar[w]), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 32) out[0] = t1 - t0;
    }
    tc::fence_before_sync(); __syncthreads();
    if (warp == 0) { tc::fence_after_sync(); tc::tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 4000;
    for (int variant = 0; variant < 2; ++variant)
        for (int M : {128, 64})
            for (int N : {32, 64, 128, 256})
                for (int nissue : {1, 2}) {
                    long long h = 0;
                    bench<<<148, 128, 200 * 1024>>>(variant, M, N, nissue, iters, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
                    printf("%-12s M=%3d N=%3d issuers=%d  %6.1f cycles per MMA per issuer, %6.1f per MMA overall (%s)\n",
                           variant ? "swizzle128" : "interleave", M, N, nissue, (double)h / iters, (double)h / iters / nissue, cudaGetErrorString(e));
                }
    printf("--- commit cost: M=128 N=128, 1 issuer, commit to an mbarrier every n MMAs ---\n");
    for (int n : {0, 32, 16, 8, 4, 2, 1}) {
        long long h = 0;
        bench<<<148, 128, 200 * 1024>>>(0, 128, 128, 1, iters, d_out, n);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("commit every %2d MMAs: %6.1f cycles per MMA  (%s)\n", n, (double)h / iters, cudaGetErrorString(e));
    }
    return 0;
}
