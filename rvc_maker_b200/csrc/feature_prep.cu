// Feature preparation of VC.voice_conversion (main/inference/convert.py:361-378), the step between the embedder and
// Synthesizer.infer: nearest x2 up-sampling of the 50 fps features to the 100 fps pitch grid, the "protect" blend with the
// un-retrieved features on unvoiced frames, and the abs-max of a synthesised piece for the pipeline's peak normalisation (:451).
#include "common.cuh"

namespace {

// out[b,t,:] = feats[b,t/2,:] * pf + feats0[b,t/2,:] * (1 - pf);  pf = protect where pitchf < 1, else 1 (:372-378).
// Products and the sum are rounded separately (no fma) so the result equals the reference's three elementwise ops bit for bit.
__global__ void __launch_bounds__(256) feature_prep_kernel(const float4* __restrict__ feats, const float4* __restrict__ feats0,
                                                           const float* __restrict__ pitchf, int64_t pitchf_bs, float protect,
                                                           float4* __restrict__ out, int Th, int T, int D4) {
    rvc::pdl_prologue();
    const int b = blockIdx.z;
    const int t = blockIdx.y;
    const float4* src = feats + ((int64_t)b * Th + (t >> 1)) * D4;
    float4* dst = out + ((int64_t)b * T + t) * D4;
    if (!feats0) {
        for (int d = threadIdx.x; d < D4; d += 256) dst[d] = src[d];
        return;
    }
    const float pv = pitchf[(int64_t)b * pitchf_bs + t];
    // pitchff = pitchf.clone(); pitchff[pitchf > 0] = 1; pitchff[pitchf < 1] = protect   (NaN stays NaN)
    const float pf = pv < 1.f ? protect : (pv > 0.f ? 1.f : pv);
    const float om = __fsub_rn(1.f, pf);
    const float4* src0 = feats0 + ((int64_t)b * Th + (t >> 1)) * D4;
    for (int d = threadIdx.x; d < D4; d += 256) {
        const float4 a = src[d], c = src0[d];
        float4 r;
        r.x = __fadd_rn(__fmul_rn(a.x, pf), __fmul_rn(c.x, om));
        r.y = __fadd_rn(__fmul_rn(a.y, pf), __fmul_rn(c.y, om));
        r.z = __fadd_rn(__fmul_rn(a.z, pf), __fmul_rn(c.z, om));
        r.w = __fadd_rn(__fmul_rn(a.w, pf), __fmul_rn(c.w, om));
        dst[d] = r;
    }
}

// result[b] = max(result[b], max_i |x[b, i]|) — non-negative floats order like their bit patterns, so atomicMax on the bits
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, int64_t x_bs, int64_t n, unsigned int* __restrict__ result) {
    const int b = blockIdx.y;
    const float* xb = x + (int64_t)b * x_bs;
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) m = fmaxf(m, fabsf(xb[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, ws[w]);
        atomicMax(result + b, __float_as_uint(m));
    }
}

// The retrieval blend of VC.voice_conversion after `score, ix = index.search(npy, k=8)` (main/inference/convert.py:353-359):
//   weight = np.square(1 / score);  weight /= weight.sum(axis=1, keepdims=True)
//   npy = np.sum(big_npy[ix] * np.expand_dims(weight, axis=2), axis=1)
//   feats = torch.from_numpy(npy).to(device) * index_rate + (1 - index_rate) * feats
// with numpy's roundings and orders: the 8 weights summed pairwise (numpy's contiguous-axis reduction), the 8 weighted neighbours added one after
// the other (a reduction over a middle axis), every product / quotient / sum rounded separately, then torch's mul, mul, add.
__global__ void __launch_bounds__(256) retrieval_blend_kernel(const float* __restrict__ feats, const float* __restrict__ big, const float* __restrict__ score,
                                                              const int64_t* __restrict__ ix, float rate, float one_minus_rate, float* __restrict__ out,
                                                              int D) {
    rvc::pdl_prologue();
    const int t = blockIdx.x;
    float w[8];
    const float* rows[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float r = __fdiv_rn(1.f, score[(int64_t)t * 8 + k]);
        w[k] = __fmul_rn(r, r);
        rows[k] = big + ix[(int64_t)t * 8 + k] * (int64_t)D;
    }
    const float ws = __fadd_rn(__fadd_rn(__fadd_rn(w[0], w[1]), __fadd_rn(w[2], w[3])), __fadd_rn(__fadd_rn(w[4], w[5]), __fadd_rn(w[6], w[7])));
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = __fdiv_rn(w[k], ws);
    for (int d = threadIdx.x; d < D; d += 256) {
        float acc = __fmul_rn(rows[0][d], w[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) acc = __fadd_rn(acc, __fmul_rn(rows[k][d], w[k]));
        out[(int64_t)t * D + d] = __fadd_rn(__fmul_rn(acc, rate), __fmul_rn(one_minus_rate, feats[(int64_t)t * D + d]));
    }
}

}  // namespace

RVC_API int rvc_retrieval_blend_f32(const float* feats, const float* big_npy, const float* score, const int64_t* ix, float index_rate,
                                    float one_minus_rate, float* out, int32_t T, int32_t D, int32_t K, void* stream) {
    if (!feats || !big_npy || !score || !ix || !out) return rvc::fail(RVC_EINVAL, "rvc_retrieval_blend_f32: null pointer");
    if (K != 8) return rvc::fail(RVC_ESHAPE, "rvc_retrieval_blend_f32: k must be 8 (index.search(npy, k=8), convert.py:353; numpy's summation order depends on it)");
    if (T < 0 || D <= 0) return rvc::fail(RVC_ESHAPE, "rvc_retrieval_blend_f32: bad shape");
    if (T == 0) return 0;
    retrieval_blend_kernel<<<T, 256, 0, (cudaStream_t)stream>>>(feats, big_npy, score, ix, index_rate, one_minus_rate, out, D);
    return rvc::check_launch("rvc_retrieval_blend_f32");
}

RVC_API int rvc_feature_prep_f32(const float* feats, const float* feats0, const float* pitchf, int64_t pitchf_bs, float protect, float* out,
                                 int32_t B, int32_t Th, int32_t T, int32_t D, void* stream) {
    if (!feats || !out) return rvc::fail(RVC_EINVAL, "rvc_feature_prep_f32: null pointer");
    if (feats0 && !pitchf) return rvc::fail(RVC_EINVAL, "rvc_feature_prep_f32: the protect blend needs pitchf");
    if (B <= 0 || B > 65535 || Th <= 0 || T < 0 || T > 2 * Th || T > 65535 || D <= 0 || (D & 3))
        return rvc::fail(RVC_ESHAPE, "rvc_feature_prep_f32: bad shape (need 0 <= T <= 2*Th, T <= 65535, D %% 4 == 0)");
    if (((uintptr_t)feats | (uintptr_t)feats0 | (uintptr_t)out) & 15) return rvc::fail(RVC_EINVAL, "rvc_feature_prep_f32: pointers must be 16-byte aligned");
    if (T == 0) return 0;
    dim3 grid(1, T, B);
    feature_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)feats, (const float4*)feats0, pitchf, pitchf_bs, protect, (float4*)out, Th, T,
                                                               D / 4);
    return rvc::check_launch("rvc_feature_prep_f32");
}

RVC_API int rvc_absmax_f32(const float* x, int64_t x_bs, int64_t n, float* result, int32_t B, void* stream) {
    if (!x || !result) return rvc::fail(RVC_EINVAL, "rvc_absmax_f32: null pointer");
    if (B <= 0 || B > 65535 || n < 0) return rvc::fail(RVC_ESHAPE, "rvc_absmax_f32: bad shape");
    if (n == 0) return 0;
    const int64_t blocks = (n + 255) / 256;
    dim3 grid((unsigned)(blocks < 592 ? blocks : 592), B);
    absmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_bs, n, (unsigned int*)result);
    return rvc::check_launch("rvc_absmax_f32");
}
