// ConvTranspose1d (+ fused pre-leaky_relu, bias and NSF noise-conv injection) and conv_post, fp32 CUDA-core path.
//
// Phase decomposition: output l = q*u + r - p (p = (k-u)/2) receives x[ci, q-m] * w[ci, co, r + m*u] for the
// M = ceil(k/u) taps m.  A thread owns ONE phase r, 4 input positions q and 8 output channels, so the weight row it
// needs is shared by its 32 accumulators; thread index t = r + u*qg makes consecutive threads write consecutive l.
#include "common.cuh"

namespace {

constexpr int CI_T = 8;
constexpr int CGN = 2;            // channel groups per CTA
constexpr int CO_T = 8 * CGN;     // 16 output channels per CTA
constexpr int WPAD = CO_T + 4;    // padded row: different phases hit different banks

__global__ void __launch_bounds__(256) convtr1d_f32_kernel(const rvc_convtr1d_desc d, const int QG, const int M) {
    extern __shared__ float smem[];
    const int u = d.u, k = d.k;
    const int TQ = u * QG;          // threads along t
    const int QT = 4 * QG;          // q positions per CTA
    const int XW = QT + M - 1;
    float* x_s = smem;                        // [CI_T][XW]
    float* w_s = smem + CI_T * XW;            // [CI_T][M*u][WPAD]
    const int KR = M * u;

    const int tid = threadIdx.x;
    const int t = tid % TQ, cg = tid / TQ;
    const int r = t % u, qg = t / u;
    const int q0 = blockIdx.x * QT;
    const int co0 = blockIdx.y * CO_T;
    const int b = blockIdx.z;
    const int p = (k - u) / 2;
    const int nthreads = TQ * CGN;

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

    const float* xb = d.x + (int64_t)b * d.x_bs;
    for (int ci0 = 0; ci0 < d.Cin; ci0 += CI_T) {
        for (int idx = tid; idx < CI_T * XW; idx += nthreads) {
            const int ci = idx / XW, j = idx - ci * XW;
            const int q = q0 - (M - 1) + j;
            float v = 0.f;
            if (ci0 + ci < d.Cin && q >= 0 && q < d.Lin) {
                const int64_t off = (int64_t)(ci0 + ci) * d.x_cs + q;
                v = xb[off];
                if (d.x2) v = ((v + d.x2[(int64_t)b * d.x_bs + off]) + d.x3[(int64_t)b * d.x_bs + off]) / d.in_div;
                v = rvc::lrelu(v, d.pre_slope);
            }
            x_s[idx] = v;
        }
        for (int idx = tid; idx < CI_T * KR * CO_T; idx += nthreads) {
            const int c = idx % CO_T;
            const int rr = idx / CO_T;       // ci*KR + kk
            const int ci = rr / KR, kk = rr - ci * KR;
            float v = 0.f;
            if (ci0 + ci < d.Cin && kk < k && co0 + c < d.Cout) v = d.w[((int64_t)(ci0 + ci) * k + kk) * d.Cout + co0 + c];
            w_s[rr * WPAD + c] = v;
        }
        __syncthreads();
        if (cg < CGN) {
#pragma unroll 1
            for (int ci = 0; ci < CI_T; ++ci) {
#pragma unroll 1
                for (int m = 0; m < M; ++m) {
                    const float* wr = w_s + (ci * KR + r + m * u) * WPAD + cg * 8;
                    const float4 wa = *reinterpret_cast<const float4*>(wr);
                    const float4 wb = *reinterpret_cast<const float4*>(wr + 4);
                    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                    const float* xr = x_s + ci * XW + (M - 1) - m + qg;
                    float xv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) xv[i] = xr[QG * i];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(wv[j], xv[i], acc[j][i]);
                }
            }
        }
        __syncthreads();
    }
    if (cg >= CGN) return;

    const int Lout = d.Lin * u;
    const int64_t har_len = (int64_t)Lout * d.sn;
    const float* hb = d.har ? d.har + (int64_t)b * d.har_bs : nullptr;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + qg + QG * i;
        const int l = q * u + r - p;
        if (l < 0 || l >= Lout) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = co0 + cg * 8 + j;
            if (co >= d.Cout) continue;
            float v = acc[j][i] + (d.bias ? d.bias[co] : 0.f);
            if (hb) {
                float nv = d.nbias ? d.nbias[co] : 0.f;
                const int64_t h0 = (int64_t)l * d.sn - d.pn;
                const float* nw = d.nw + (int64_t)co * d.kn;
                for (int jj = 0; jj < d.kn; ++jj) {
                    const int64_t h = h0 + jj;
                    if (h >= 0 && h < har_len) nv = fmaf(nw[jj], hb[h], nv);
                }
                v += nv;
            }
            d.y[(int64_t)b * d.y_bs + (int64_t)co * d.y_cs + l] = v;
        }
    }
}

// conv_post: Cout = 1, k = 7, pad 3, no bias, pre-leaky_relu(slope), tanh.  One thread = 4 consecutive outputs.
__global__ void __launch_bounds__(256) conv_post_f32_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                            const float* __restrict__ x3, float in_div, int64_t x_bs, int64_t x_cs,
                                                            const float* __restrict__ w, float* __restrict__ y, int64_t y_bs,
                                                            int C, int L, float slope) {
    rvc::pdl_prologue();
    extern __shared__ float w_s[];  // [C][7]
    for (int i = threadIdx.x; i < C * 7; i += blockDim.x) w_s[i] = w[i];
    __syncthreads();
    const int b = blockIdx.y;
    const int l0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (l0 >= L) return;
    const float* xb = x + (int64_t)b * x_bs;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const bool interior = !x2 && (l0 >= 4) && (l0 + 8 <= L) && ((x_cs & 3) == 0) && ((reinterpret_cast<uintptr_t>(xb) & 15) == 0);
    int c_first = 0;
    if (interior) {
        // four channels at a time, all twelve 128-bit loads in flight before the first use: channel by channel the thread made C dependent
        // round trips to memory (27 us at C = 32, L = 480 000 for 61 MB).  Same accumulation order (c ascending, tap ascending).
        for (; c_first + 4 <= C; c_first += 4) {
            float4 q[4][3];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float* xr = xb + (int64_t)(c_first + u) * x_cs;
                q[u][0] = *reinterpret_cast<const float4*>(xr + l0 - 4);
                q[u][1] = *reinterpret_cast<const float4*>(xr + l0);
                q[u][2] = *reinterpret_cast<const float4*>(xr + l0 + 4);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float win[12] = {q[u][0].x, q[u][0].y, q[u][0].z, q[u][0].w, q[u][1].x, q[u][1].y, q[u][1].z, q[u][1].w,
                                 q[u][2].x, q[u][2].y, q[u][2].z, q[u][2].w};
#pragma unroll
                for (int j = 0; j < 12; ++j) win[j] = rvc::lrelu(win[j], slope);
#pragma unroll
                for (int kk = 0; kk < 7; ++kk) {
                    const float wv = w_s[(c_first + u) * 7 + kk];
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv, win[i + kk + 1], acc[i]);
                }
            }
        }
    }
    for (int c = c_first; c < C; ++c) {
        const float* xr = xb + (int64_t)c * x_cs;
        float win[12];  // x[l0-4 .. l0+7]
        if (x2) {
            const float* xr2 = x2 + (int64_t)b * x_bs + (int64_t)c * x_cs;
            const float* xr3 = x3 + (int64_t)b * x_bs + (int64_t)c * x_cs;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const int l = l0 - 4 + j;
                win[j] = (l >= 0 && l < L) ? ((xr[l] + xr2[l]) + xr3[l]) / in_div : 0.f;
            }
        } else if (interior) {
            const float4 a = *reinterpret_cast<const float4*>(xr + l0 - 4);
            const float4 bq = *reinterpret_cast<const float4*>(xr + l0);
            const float4 cq = *reinterpret_cast<const float4*>(xr + l0 + 4);
            win[0] = a.x; win[1] = a.y; win[2] = a.z; win[3] = a.w;
            win[4] = bq.x; win[5] = bq.y; win[6] = bq.z; win[7] = bq.w;
            win[8] = cq.x; win[9] = cq.y; win[10] = cq.z; win[11] = cq.w;
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const int l = l0 - 4 + j;
                win[j] = (l >= 0 && l < L) ? xr[l] : 0.f;
            }
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) win[j] = rvc::lrelu(win[j], slope);
#pragma unroll
        for (int kk = 0; kk < 7; ++kk) {
            const float wv = w_s[c * 7 + kk];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv, win[i + kk + 1], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (l0 + i < L) y[(int64_t)b * y_bs + l0 + i] = tanhf(acc[i]);
}

}  // namespace

RVC_API int rvc_convtr1d_f32(const rvc_convtr1d_desc* dp, void* stream) {
    if (!dp) return rvc::fail(RVC_EINVAL, "rvc_convtr1d_f32: null descriptor");
    const rvc_convtr1d_desc& d = *dp;
    if (!d.x || !d.w || !d.y) return rvc::fail(RVC_EINVAL, "rvc_convtr1d_f32: null x/w/y");
    if (d.B <= 0 || d.B > 65535 || d.Cin <= 0 || d.Cout <= 0 || d.Lin < 0 || d.k <= 0 || d.u <= 0)
        return rvc::fail(RVC_ESHAPE, "rvc_convtr1d_f32: bad shape");
    if (d.k < d.u || ((d.k - d.u) & 1) || d.u > 64)
        return rvc::fail(RVC_ESHAPE, "rvc_convtr1d_f32: needs k >= u, (k-u) even, u <= 64 (k=%d u=%d)", d.k, d.u);
    if (d.har && (!d.nw || d.kn <= 0 || d.sn <= 0)) return rvc::fail(RVC_EINVAL, "rvc_convtr1d_f32: noise conv arguments");
    if ((d.x2 == nullptr) != (d.x3 == nullptr)) return rvc::fail(RVC_EINVAL, "rvc_convtr1d_f32: x2 and x3 go together");
    if (d.Lin == 0) return 0;
    const int QG = 128 / d.u > 0 ? 128 / d.u : 1;
    const int M = rvc::cdiv(d.k, d.u);
    const int QT = 4 * QG, XW = QT + M - 1;
    const size_t smem = (size_t)(CI_T * XW + CI_T * M * d.u * WPAD) * sizeof(float);
    if (smem > 200 * 1024) return rvc::fail(RVC_ESHAPE, "rvc_convtr1d_f32: kernel too large for shared memory");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(convtr1d_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_convtr1d_f32: smem opt-in: %s", cudaGetErrorString(e));
    }
    const int p = (d.k - d.u) / 2;
    const int nq = d.Lin + rvc::cdiv(p, d.u);  // q values that can reach a valid output position
    dim3 grid(rvc::cdiv(nq, QT), rvc::cdiv(d.Cout, CO_T), d.B);
    convtr1d_f32_kernel<<<grid, d.u * QG * CGN, smem, (cudaStream_t)stream>>>(d, QG, M);
    return rvc::check_launch("rvc_convtr1d_f32");
}

RVC_API int rvc_conv_post_f32(const float* x, const float* x2, const float* x3, float in_div, int64_t x_bs, int64_t x_cs, const float* w,
                                 float* y, int64_t y_bs, int32_t B, int32_t C, int32_t L, float slope, void* stream) {
    if (!x || !w || !y) return rvc::fail(RVC_EINVAL, "rvc_conv_post_f32: null pointer");
    if ((x2 == nullptr) != (x3 == nullptr)) return rvc::fail(RVC_EINVAL, "rvc_conv_post_f32: x2 and x3 go together");
    if (B <= 0 || B > 65535 || C <= 0 || C > 1024 || L < 0) return rvc::fail(RVC_ESHAPE, "rvc_conv_post_f32: bad shape");
    if (L == 0) return 0;
    dim3 grid(rvc::cdiv(rvc::cdiv(L, 4), 256), B);
    rvc::launch_pdl(conv_post_f32_kernel, dim3(grid), dim3(256), (size_t)(C * 7 * sizeof(float)), (cudaStream_t)stream, x, x2, x3, in_div, x_bs, x_cs, w, y, y_bs, C, L, slope);
    return rvc::check_launch("rvc_conv_post_f32");
}

// ------------------------------------------------------------------------------------------------
// NSF noise convs (synthesizers.py:134-136): y[b,co,l] = bias[co] + sum_j w[co][j] * har[b, l*sn - pn + j], Cin = 1.
// One thread = one output position; its har window sits in registers (16 taps at a time) and is reused by 32 channels.
// ------------------------------------------------------------------------------------------------
namespace {
// One thread = NP output positions (t, t + 128, ... of the CTA's 128 * NP) x 32 channels: a tap costs NP source reads + 8 broadcast
// LDS.128 of weights for 32 * NP FMAs.  NP = 1 (64 registers, 8 CTAs per SM): with NP = 2 the 160 registers of a thread limited an SM to three CTAs —
// 12 warps, one wave of 376 CTAs at stage 0, 4.2 waves at stage 1 — and both launches were latency-bound (ncu: warps active 16 %, issue 42 %).  The source window is stored with one pad word per 32 (index i + i/32): threads read it at
// stride sn, and e.g. sn = 40 would otherwise put every 4th thread on the same bank (8-way conflict).
__device__ __forceinline__ int hpad(int i) { return i + (i >> 5); }

template <int NP>
__global__ void __launch_bounds__(128, NP == 1 ? 8 : 3) noise_conv_f32_kernel(const float* __restrict__ har, int64_t har_bs, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y, int64_t y_bs, int C,
                                                             int Lout, int kn, int sn, int pn, int blocked) {
    rvc::pdl_prologue();
    extern __shared__ __align__(16) float w_s[];   // [kn][32] weights, then the padded source window
    float* h_s = w_s + 32 * kn;
    const int co0 = blockIdx.y * 32, b = blockIdx.z;
    const int64_t hlen = (int64_t)Lout * sn;
    const float* hb = har + (int64_t)b * har_bs;
    const int64_t g0 = (int64_t)blockIdx.x * (128 * NP) * sn - pn;       // source index of window element 0
    const int span = (128 * NP - 1) * sn + kn;
    for (int i = threadIdx.x; i < 32 * kn; i += 128) {
        const int c = i / kn, j = i - c * kn;
        w_s[j * 32 + c] = (co0 + c < C) ? w[(int64_t)(co0 + c) * kn + j] : 0.f;
    }
    // the source window: 128-bit loads where the window lies inside the signal and is 16-byte aligned (a long window — 10 280 samples at sn = 40 —
    // filled one sample per thread and round trip was a third of the launch)
    const bool vec = (((g0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(hb) & 15) == 0) && ((har_bs & 3) == 0));
    const int nvec = vec ? span / 4 : 0;
#pragma unroll 4
    for (int i4 = threadIdx.x; i4 < nvec; i4 += 128) {
        const int64_t idx = g0 + 4 * (int64_t)i4;
        float4 v;
        if (idx >= 0 && idx + 3 < hlen) v = *reinterpret_cast<const float4*>(hb + idx);
        else {
            v.x = (idx >= 0 && idx < hlen) ? hb[idx] : 0.f;
            v.y = (idx + 1 >= 0 && idx + 1 < hlen) ? hb[idx + 1] : 0.f;
            v.z = (idx + 2 >= 0 && idx + 2 < hlen) ? hb[idx + 2] : 0.f;
            v.w = (idx + 3 >= 0 && idx + 3 < hlen) ? hb[idx + 3] : 0.f;
        }
        const int i = 4 * i4;
        h_s[hpad(i)] = v.x; h_s[hpad(i + 1)] = v.y; h_s[hpad(i + 2)] = v.z; h_s[hpad(i + 3)] = v.w;
    }
    for (int i = 4 * nvec + threadIdx.x; i < span; i += 128) {
        const int64_t idx = g0 + i;
        h_s[hpad(i)] = (idx >= 0 && idx < hlen) ? hb[idx] : 0.f;
    }
    __syncthreads();
    const int l0 = blockIdx.x * (128 * NP) + threadIdx.x;
    if (l0 >= Lout) return;
    float acc[NP][32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const float bv = (co0 + c < C && bias) ? bias[co0 + c] : 0.f;
#pragma unroll
        for (int h = 0; h < NP; ++h) acc[h][c] = bv;
    }
    for (int j = 0; j < kn; ++j) {
        float hv[NP];
#pragma unroll
        for (int h = 0; h < NP; ++h) hv[h] = h_s[hpad((threadIdx.x + 128 * h) * sn + j)];
        const float4* wj = reinterpret_cast<const float4*>(w_s + j * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            const float4 wv = wj[c4];
#pragma unroll
            for (int h = 0; h < NP; ++h) {
                acc[h][4 * c4 + 0] = fmaf(wv.x, hv[h], acc[h][4 * c4 + 0]);
                acc[h][4 * c4 + 1] = fmaf(wv.y, hv[h], acc[h][4 * c4 + 1]);
                acc[h][4 * c4 + 2] = fmaf(wv.z, hv[h], acc[h][4 * c4 + 2]);
                acc[h][4 * c4 + 3] = fmaf(wv.w, hv[h], acc[h][4 * c4 + 3]);
            }
        }
    }
#pragma unroll
    for (int half = 0; half < NP; ++half) {
        const int l = l0 + 128 * half;
        if (l >= Lout) break;
        if (blocked) {   // [b][c/8][l][8]: one 32-byte sector per 8 channels
#pragma unroll
            for (int g = 0; g < 4; ++g)
                if (co0 + 8 * g < C) {
                    // ONE 256-bit store per lane: a warp writes 1 KB of whole 32-byte sectors (two 128-bit stores per lane touch every sector twice, half each)
                    float* yp = y + (int64_t)b * y_bs + ((int64_t)(co0 / 8 + g) * Lout + l) * 8;
                    const float* v = &acc[half][8 * g];
                    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                                 ::"l"(yp), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
                }
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (co0 + c < C) y[(int64_t)b * y_bs + (int64_t)(co0 + c) * Lout + l] = acc[half][c];
        }
    }
}
}  // namespace

RVC_API int rvc_noise_conv_f32(const float* har, int64_t har_bs, const float* w, const float* bias, float* y, int64_t y_bs, int32_t B,
                               int32_t C, int32_t Lout, int32_t kn, int32_t sn, int32_t pn, int32_t blocked, void* stream) {
    if (!har || !w || !y) return rvc::fail(RVC_EINVAL, "rvc_noise_conv_f32: null pointer");
    if (blocked && ((C & 7) || ((uintptr_t)y & 31) || (y_bs & 7))) return rvc::fail(RVC_ESHAPE, "rvc_noise_conv_f32: blocked output needs C %% 8 == 0 and 32-byte alignment");
    if (B <= 0 || B > 65535 || C <= 0 || Lout < 0 || kn <= 0 || kn > 1024 || sn <= 0 || pn < 0) return rvc::fail(RVC_ESHAPE, "rvc_noise_conv_f32: bad shape");
    if (Lout == 0) return 0;
    // short kernels (stage >= 1: 8 taps and fewer) are latency-bound: one position per thread, 8 CTAs per SM (stage 1: 31 -> 23 us); the long
    // stage-0 kernel (80 taps) is bound by its shared-memory reads, which two positions per thread halve (33 us against 35)
    const int NP = kn > 16 ? 2 : 1;
    const int span = (128 * NP - 1) * sn + kn;
    const size_t smem = (size_t)(32 * kn + span + span / 32 + 1) * sizeof(float);
    if (smem > 200 * 1024) return rvc::fail(RVC_ESHAPE, "rvc_noise_conv_f32: kn = %d, sn = %d need %zu bytes of shared memory", kn, sn, smem);
    if (smem > 48 * 1024) {
        cudaError_t e = NP == 2 ? cudaFuncSetAttribute(noise_conv_f32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                : cudaFuncSetAttribute(noise_conv_f32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_noise_conv_f32: smem opt-in: %s", cudaGetErrorString(e));
    }
    dim3 grid(rvc::cdiv(Lout, 128 * NP), rvc::cdiv(C, 32), B);
    if (NP == 2)
        rvc::launch_pdl(noise_conv_f32_kernel<2>, dim3(grid), dim3(128), smem, (cudaStream_t)stream, har, har_bs, w, bias, y, y_bs, C, Lout, kn, sn, pn, (int)blocked);
    else
        rvc::launch_pdl(noise_conv_f32_kernel<1>, dim3(grid), dim3(128), smem, (cudaStream_t)stream, har, har_bs, w, bias, y, y_bs, C, Lout, kn, sn, pn, (int)blocked);
    return rvc::check_launch("rvc_noise_conv_f32");
}
