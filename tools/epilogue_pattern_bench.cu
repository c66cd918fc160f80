// Memory-pattern experiment behind the conv2 epilogue layout decision (not part of the library).
// Emulates the epilogue's global traffic only: per work item a warp reads a residual tile (32 positions x 32 channels fp32),
// writes the same shape back and writes the packed fp16 copy (32 positions x 4 x 16 B).
//   mode 0: NCL fp32, one position per lane, one LDG.32/STG.32 per channel (what the kernel does)
//   mode 1: channel-blocked fp32 [c/8][L][8], 2 x 128-bit per 8-channel chunk
//   mode 2: channel-blocked fp32, one 256-bit access per chunk
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/epb tools/epilogue_pattern_bench.cu && /tmp/epb
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld256(const float* p) {
    f8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(float* p, const f8& r) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(384) k(const float* __restrict__ res, float* __restrict__ y, uint4* __restrict__ ya, int C, int L, int Lp, int njobs,
                                         int do_ld, int do_st, int do_pk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = warp >> 2, q = warp & 3;
    const int lt = (L + 255) / 256, cchunks = C / 32;
    const int ntiles = njobs * lt;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int job = t / lt, l0 = (t % lt) * 256;
        const float* rj = res + (size_t)job * C * L;
        float* yj = y + (size_t)job * C * L;
        uint4* yaj = ya + (size_t)job * (C / 8) * Lp;
        for (int it = wg; it < 2 * cchunks; it += 3) {
            const int m = it / cchunks, c0 = (it % cchunks) * 32;
            const int l = l0 + m * 128 + q * 32 + lane;
            if (l >= L) continue;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 1.f;
            if (do_ld) {
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = rj[(size_t)(c0 + i) * L + l];
                } else if (MODE == 1) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4* p = reinterpret_cast<const float4*>(rj + ((size_t)(c0 / 8 + g) * L + l) * 8);
                        float4 a = p[0], b = p[1];
                        v[8 * g] = a.x; v[8 * g + 1] = a.y; v[8 * g + 2] = a.z; v[8 * g + 3] = a.w;
                        v[8 * g + 4] = b.x; v[8 * g + 5] = b.y; v[8 * g + 6] = b.z; v[8 * g + 7] = b.w;
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        f8 a = ld256(rj + ((size_t)(c0 / 8 + g) * L + l) * 8);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[8 * g + i] = a.v[i];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v[i] * 1.0001f + 0.5f;
            if (do_st) {
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) yj[(size_t)(c0 + i) * L + l] = v[i];
                } else if (MODE == 1) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float4* p = reinterpret_cast<float4*>(yj + ((size_t)(c0 / 8 + g) * L + l) * 8);
                        p[0] = make_float4(v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3]);
                        p[1] = make_float4(v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]);
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        f8 a;
#pragma unroll
                        for (int i = 0; i < 8; ++i) a.v[i] = v[8 * g + i];
                        st256(yj + ((size_t)(c0 / 8 + g) * L + l) * 8, a);
                    }
                }
            }
            if (do_pk) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 pk = make_uint4(__float_as_uint(v[8 * g]), __float_as_uint(v[8 * g + 2]), __float_as_uint(v[8 * g + 4]), __float_as_uint(v[8 * g + 6]));
                    yaj[(size_t)(c0 / 8 + g) * Lp + 64 + l] = pk;
                }
            }
        }
    }
}

int main() {
    const int njobs = 3;
    struct { int C, L; } cases[] = {{128, 120000}, {32, 480000}, {256, 12000}};
    for (auto cs : cases) {
        const int C = cs.C, L = cs.L, Lp = 64 + (L + 255) / 256 * 256 + 64;
        float *res, *y; uint4* ya;
        cudaMalloc(&res, (size_t)njobs * C * L * 4); cudaMalloc(&y, (size_t)njobs * C * L * 4); cudaMalloc(&ya, (size_t)njobs * (C / 8) * Lp * 16);
        cudaMemset(res, 0, (size_t)njobs * C * L * 4);
        char* flush; cudaMalloc(&flush, 256 << 20);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        printf("C=%d L=%d (res %.0f MB, y %.0f MB, packed %.0f MB)\n", C, L, njobs * C * L * 4e-6, njobs * C * L * 4e-6, njobs * C * L * 2e-6);
        for (int mode = 0; mode < 3; ++mode)
            for (int sel = 0; sel < 4; ++sel) {
                const int do_ld = sel == 0 || sel == 1, do_st = sel == 0 || sel == 2, do_pk = sel == 0 || sel == 3;
                float best = 1e9;
                for (int rep = 0; rep < 6; ++rep) {
                    cudaMemsetAsync(flush, rep, 256 << 20);
                    cudaEventRecord(e0);
                    if (mode == 0) k<0><<<148, 384>>>(res, y, ya, C, L, Lp, njobs, do_ld, do_st, do_pk);
                    if (mode == 1) k<1><<<148, 384>>>(res, y, ya, C, L, Lp, njobs, do_ld, do_st, do_pk);
                    if (mode == 2) k<2><<<148, 384>>>(res, y, ya, C, L, Lp, njobs, do_ld, do_st, do_pk);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (rep && ms < best) best = ms;
                }
                const double bytes = (double)njobs * C * L * (do_ld * 4 + do_st * 4 + do_pk * 2);
                printf("  mode %d %-14s %7.1f us  %6.2f TB/s\n", mode, sel == 0 ? "ld+st+packed" : sel == 1 ? "ld only" : sel == 2 ? "st only" : "packed only",
                       best * 1e3, bytes / best * 1e-9);
            }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaFree(res); cudaFree(y); cudaFree(ya); cudaFree(flush);
    }
    return 0;
}
