// Generic stride-1 Conv1d on CUDA cores, fp32, with the path's fusions (see include/rvc_b200.h: rvc_conv1d_f32).
//
// Tiling: one CTA = 128 threads computes a [64 out-channel x 64 position] tile (32 gate channels when gated).
// Thread (tl = tid%16, tc = tid/16) owns positions l0+tl+16*i (i<4; interleaved so shared-memory reads of the
// input window are bank-conflict free and global stores coalesce) and 8 weight columns (two float4 broadcast loads).
// Input channels stream through shared memory 8 at a time: x_s[8][64+(K-1)*dil] with the pre-activation and mask
// applied at load, w_s[8][K][64].  HBM traffic per CTA = the input tile once per 64 output channels + the output once.
#include "common.cuh"

namespace {

constexpr int L_T = 64;
constexpr int CI_T = 8;
constexpr int NT = 128;
constexpr int WCOLS = 64;

template <bool GATED>
__global__ void __launch_bounds__(NT) conv1d_f32_kernel(const rvc_conv1d_desc d) {
    extern __shared__ float smem[];
    const int K = d.K, dil = d.dil;
    const int XW = L_T + (K - 1) * dil;
    float* x_s = smem;                 // [CI_T][XW]
    float* w_s = smem + CI_T * XW;     // [CI_T][K][WCOLS]

    const int tid = threadIdx.x;
    const int tl = tid & 15, tc = tid >> 4;
    const int l0 = blockIdx.x * L_T;
    const int b = blockIdx.z;
    constexpr int CO_TILE = GATED ? 32 : 64;
    const int co0 = blockIdx.y * CO_TILE;
    const int Cw = GATED ? 2 * d.Cout : d.Cout;
    const int offA = GATED ? tc * 4 : tc * 8;
    const int offB = GATED ? 32 + tc * 4 : tc * 8 + 4;

    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

    const float* xb = d.x + (int64_t)b * d.x_bs;
    const float* mb = d.mask ? d.mask + (int64_t)b * d.L : nullptr;

    for (int ci0 = 0; ci0 < d.Cin; ci0 += CI_T) {
        // ---- stage input window (pre-activation + mask fused)
        for (int idx = tid; idx < CI_T * XW; idx += NT) {
            const int ci = idx / XW, j = idx - ci * XW;
            const int l = l0 - d.pad + j;
            float v = 0.f;
            if (ci0 + ci < d.Cin && l >= 0 && l < d.L) {
                v = xb[(int64_t)(ci0 + ci) * d.x_cs + l];
                if (d.pre_lrelu) v = rvc::lrelu(v, d.pre_slope);
                if (d.in_mask) v *= mb[l];
            }
            x_s[idx] = v;
        }
        // ---- stage weights [ci][k][col]
        for (int idx = tid; idx < CI_T * K * WCOLS; idx += NT) {
            const int c = idx & (WCOLS - 1);
            const int r = idx >> 6;  // ci*K + kk
            const int ci = r / K;
            int col;
            bool ok;
            if (GATED) {
                const int cc = c & 31;
                col = (c < 32 ? 0 : d.Cout) + co0 + cc;
                ok = (co0 + cc) < d.Cout;
            } else {
                col = co0 + c;
                ok = col < d.Cout;
            }
            float v = 0.f;
            if (ok && ci0 + ci < d.Cin) v = d.w[((int64_t)ci0 * K + r) * Cw + col];
            w_s[idx] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < CI_T; ++ci) {
            const float* xr = x_s + ci * XW + tl;
            const float* wr = w_s + ci * K * WCOLS;
#pragma unroll 1
            for (int kk = 0; kk < K; ++kk) {
                float xv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) xv[i] = xr[kk * dil + 16 * i];
                const float4 wa = *reinterpret_cast<const float4*>(wr + kk * WCOLS + offA);
                const float4 wb = *reinterpret_cast<const float4*>(wr + kk * WCOLS + offB);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(wv[j], xv[i], acc[j][i]);
            }
        }
        __syncthreads();
    }

    // ---- epilogue
    const float* bb = d.bias_b ? d.bias_b + (int64_t)b * d.bias_b_bs : nullptr;
    constexpr int NJ = GATED ? 4 : 8;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int co = co0 + (GATED ? tc * 4 : tc * 8) + j;
        if (co >= d.Cout) continue;
        float b0 = d.bias ? d.bias[co] : 0.f;
        if (bb) b0 += bb[co];
        float b1 = 0.f;
        if (GATED) {
            b1 = d.bias ? d.bias[co + d.Cout] : 0.f;
            if (bb) b1 += bb[co + d.Cout];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int l = l0 + tl + 16 * i;
            if (l >= d.L) continue;
            float v;
            if (GATED) {
                v = tanhf(acc[j][i] + b0) * rvc::sigmoidf_acc(acc[j + 4][i] + b1);
            } else {
                v = acc[j][i] + b0;
                if (d.act == RVC_ACT_RELU) v = fmaxf(v, 0.f);
                else if (d.act == RVC_ACT_TANH) v = tanhf(v);
            }
            v *= d.alpha;
            if (d.res) v += d.res[(int64_t)b * d.res_bs + (int64_t)co * d.res_cs + l];
            v *= d.out_scale;
            if (d.out_mask) v *= mb[l];
            float* yp = d.y + (int64_t)b * d.y_bs + (int64_t)co * d.y_cs + l;
            if (d.beta != 0.f) v += d.beta * (*yp);
            *yp = v;
        }
    }
}

}  // namespace

RVC_API int rvc_conv1d_f32(const rvc_conv1d_desc* dp, void* stream) {
    if (!dp) return rvc::fail(RVC_EINVAL, "rvc_conv1d_f32: null descriptor");
    const rvc_conv1d_desc& d = *dp;
    if (!d.x || !d.w || !d.y) return rvc::fail(RVC_EINVAL, "rvc_conv1d_f32: null x/w/y");
    if (d.B <= 0 || d.Cin <= 0 || d.Cout <= 0 || d.L < 0 || d.K <= 0 || d.dil <= 0)
        return rvc::fail(RVC_ESHAPE, "rvc_conv1d_f32: bad shape B=%d Cin=%d Cout=%d L=%d K=%d dil=%d", d.B, d.Cin, d.Cout, d.L, d.K, d.dil);
    if ((d.in_mask || d.out_mask) && !d.mask) return rvc::fail(RVC_EINVAL, "rvc_conv1d_f32: mask flag set without mask pointer");
    if (d.B > 65535) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_f32: B > 65535");
    if (d.L == 0) return 0;
    const int XW = L_T + (d.K - 1) * d.dil;
    const size_t smem = (size_t)(CI_T * XW + CI_T * d.K * WCOLS) * sizeof(float);
    if (smem > 200 * 1024) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_f32: K*dil too large for the shared-memory tile");
    const int co_tile = d.gated ? 32 : 64;
    dim3 grid(rvc::cdiv(d.L, L_T), rvc::cdiv(d.Cout, co_tile), d.B);
    auto kern = d.gated ? conv1d_f32_kernel<true> : conv1d_f32_kernel<false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_conv1d_f32: smem opt-in: %s", cudaGetErrorString(e));
    }
    kern<<<grid, NT, smem, (cudaStream_t)stream>>>(d);
    return rvc::check_launch("rvc_conv1d_f32");
}
