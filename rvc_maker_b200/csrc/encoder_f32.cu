// TextEncoder-side fp32 kernels: phone/pitch embedding, channel LayerNorm, relative-position attention, z_p sampling.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

namespace {

// ------------------------------------------------------------------------------------------------
// embedding: x[b,c,t] = lrelu((phone[b,t,:].W[c,:] + bias[c] + E[pitch[b,t],c]) * sqrt(H), 0.1) * mask
// CTA = 128 threads -> tile 32 t x 64 c; thread (tt = tid%16, tc = tid/16): t = t0+tt+16*i (i<2), c = c0+tc*8+j.
// ------------------------------------------------------------------------------------------------
constexpr int E_KT = 16;
__global__ void __launch_bounds__(128) embed_f32_kernel(const float* __restrict__ phone, const int64_t* __restrict__ pitch,
                                                        const int64_t* __restrict__ lengths, const float* __restrict__ wT,
                                                        const float* __restrict__ bias, const float* __restrict__ emb_pitch,
                                                        float* __restrict__ x, float* __restrict__ mask, int T, int D, int H,
                                                        float scale) {
    __shared__ float a_s[E_KT][33];
    __shared__ float w_s[E_KT][64];
    const int tid = threadIdx.x, tt = tid & 15, tc = tid >> 4;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 64, b = blockIdx.z;
    const float* pb = phone + (int64_t)b * T * D;
    float acc[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.f;
    for (int d0 = 0; d0 < D; d0 += E_KT) {
        for (int idx = tid; idx < 32 * E_KT; idx += 128) {
            const int dd = idx & (E_KT - 1), t = idx >> 4;
            a_s[dd][t] = (t0 + t < T && d0 + dd < D) ? pb[(int64_t)(t0 + t) * D + d0 + dd] : 0.f;
        }
        for (int idx = tid; idx < E_KT * 64; idx += 128) {
            const int c = idx & 63, dd = idx >> 6;
            w_s[dd][c] = (d0 + dd < D && c0 + c < H) ? wT[(int64_t)(d0 + dd) * H + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int dd = 0; dd < E_KT; ++dd) {
            const float a0 = a_s[dd][tt], a1 = a_s[dd][tt + 16];
            const float4 wa = *reinterpret_cast<const float4*>(&w_s[dd][tc * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&w_s[dd][tc * 8 + 4]);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[j][0] = fmaf(a0, wv[j], acc[j][0]);
                acc[j][1] = fmaf(a1, wv[j], acc[j][1]);
            }
        }
        __syncthreads();
    }
    const int64_t len = lengths[b];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int t = t0 + tt + 16 * i;
        if (t >= T) continue;
        const float m = (t < len) ? 1.f : 0.f;
        if (blockIdx.y == 0 && tc == 0) mask[(int64_t)b * T + t] = m;
        const int64_t pi = pitch ? pitch[(int64_t)b * T + t] : 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + tc * 8 + j;
            if (c >= H) continue;
            float v = acc[j][i] + bias[c];
            if (pitch) v += emb_pitch[pi * H + c];
            v = rvc::lrelu(v * scale, 0.1f) * m;
            x[((int64_t)b * H + c) * T + t] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over channels of [B,C,T]; blockDim (32 t, 8 channel groups).  Group ty owns the 8-channel chunks
// ty, ty+8, ... so that it can also emit the packed (16 B = 8 channels) fp16/bf16 copy the tensor-core convs read.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ln_pack2(float a, float b, int fp16) {
    if (fp16) {
        __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// low plane of the split-fp16 form: fp16(value - float(hi)) for the two values packed in `hi`
__device__ __forceinline__ uint32_t ln_lo2(float a, float b, uint32_t hi) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    __half2 h = __floats2half2_rn(a - hf.x, b - hf.y);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                            const float* __restrict__ r2, const float* __restrict__ r3,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mask, float* __restrict__ y, int C, int T,
                                                            float eps, int out_mask, uint8_t* __restrict__ y_act, int Lp, int fp16) {
    rvc::pdl_prologue();
    __shared__ float red[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int t = blockIdx.x * 32 + tx, b = blockIdx.y;
    const bool ok = t < T;
    const int64_t base = (int64_t)b * C * T + t;
    const int nchunk = (C + 7) / 8;
    float s = 0.f;
    if (ok)
        for (int ch = ty; ch < nchunk; ch += 8)
            for (int c = ch * 8; c < min(C, ch * 8 + 8); ++c) {
                float v = x[base + (int64_t)c * T];
                if (r) v += r[base + (int64_t)c * T];
                if (r2) v += r2[base + (int64_t)c * T] + r3[base + (int64_t)c * T];
                s += v;
            }
    red[ty][tx] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) mean += red[i][tx];
    mean /= (float)C;
    __syncthreads();
    float q = 0.f;
    if (ok)
        for (int ch = ty; ch < nchunk; ch += 8)
            for (int c = ch * 8; c < min(C, ch * 8 + 8); ++c) {
                float v = x[base + (int64_t)c * T];
                if (r) v += r[base + (int64_t)c * T];
                if (r2) v += r2[base + (int64_t)c * T] + r3[base + (int64_t)c * T];
                v -= mean;
                q = fmaf(v, v, q);
            }
    red[ty][tx] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) var += red[i][tx];
    var /= (float)C;
    const float rstd = rsqrtf(var + eps);
    if (!ok) return;
    const float mk = mask ? mask[(int64_t)b * T + t] : 1.f;
    const float m = out_mask ? mk : 1.f;
    for (int ch = ty; ch < nchunk; ch += 8) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = ch * 8 + i;
            o[i] = 0.f;
            if (c < C) {
                float v = x[base + (int64_t)c * T];
                if (r) v += r[base + (int64_t)c * T];
                if (r2) v += r2[base + (int64_t)c * T] + r3[base + (int64_t)c * T];
                o[i] = (v - mean) * rstd * gamma[c] + beta[c];
                y[base + (int64_t)c * T] = o[i] * m;
            }
        }
        if (y_act) {  // packed copy is always masked: it only ever feeds convs whose input the reference masks (or 1x1 convs)
            uint4 pk;
            pk.x = ln_pack2(o[0] * mk, o[1] * mk, fp16 & 1); pk.y = ln_pack2(o[2] * mk, o[3] * mk, fp16 & 1);
            pk.z = ln_pack2(o[4] * mk, o[5] * mk, fp16 & 1); pk.w = ln_pack2(o[6] * mk, o[7] * mk, fp16 & 1);
            if (fp16 & 2) {   // RVC_DT_FP16_SPLIT: planes [hi | lo] of 2 * nchunk chunks per batch item, lo = fp16(value - hi)
                *reinterpret_cast<uint4*>(y_act + (((size_t)b * 2 * nchunk + ch) * Lp + 64 + t) * 16) = pk;
                *reinterpret_cast<uint4*>(y_act + (((size_t)b * 2 * nchunk + nchunk + ch) * Lp + 64 + t) * 16) =
                    make_uint4(ln_lo2(o[0] * mk, o[1] * mk, pk.x), ln_lo2(o[2] * mk, o[3] * mk, pk.y), ln_lo2(o[4] * mk, o[5] * mk, pk.z), ln_lo2(o[6] * mk, o[7] * mk, pk.w));
            } else
            *reinterpret_cast<uint4*>(y_act + (((size_t)b * nchunk + ch) * Lp + 64 + t) * 16) = pk;
        }
    }
}

// Register-resident variant for C <= 64 * NCH (the encoder's C = 192 -> NCH = 3): every thread loads its NCH x 8 channels (and
// the residual) once, all loads in flight together, and both reductions and the output come from registers — one global round
// trip instead of three (11.9 -> ~4 us at [192, 1000]: the launch is latency-bound, 32 CTAs).
template <int NCH>
__global__ void __launch_bounds__(256) layernorm_reg_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                            const float* __restrict__ r2, const float* __restrict__ r3,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mask, float* __restrict__ y, int C, int T,
                                                            float eps, int out_mask, uint8_t* __restrict__ y_act, int Lp, int fp16) {
    rvc::pdl_prologue();
    __shared__ float red[2][8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int t = blockIdx.x * 32 + tx, b = blockIdx.y;
    const bool ok = t < T;
    const int64_t base = (int64_t)b * C * T + t;
    const int nchunk = (C + 7) / 8;
    float v[NCH][8];
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = (ty + 8 * k) * 8 + i;
            v[k][i] = (ok && c < C) ? x[base + (int64_t)c * T] : 0.f;
        }
    if (r) {
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = (ty + 8 * k) * 8 + i;
                if (ok && c < C) v[k][i] += r[base + (int64_t)c * T];
            }
    }
    if (r2) {   // split-K partial sums of the producing conv (the FFN's second conv runs as three input-channel slices)
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = (ty + 8 * k) * 8 + i;
                if (ok && c < C) v[k][i] += r2[base + (int64_t)c * T] + r3[base + (int64_t)c * T];
            }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[k][i];
    red[0][ty][tx] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) mean += red[0][i][tx];
    mean /= (float)C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = (ty + 8 * k) * 8 + i;
            const float d = v[k][i] - mean;
            if (c < C) q = fmaf(d, d, q);
        }
    red[1][ty][tx] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) var += red[1][i][tx];
    var /= (float)C;
    const float rstd = rsqrtf(var + eps);
    if (!ok) return;
    const float mk = mask ? mask[(int64_t)b * T + t] : 1.f;
    const float m = out_mask ? mk : 1.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int ch = ty + 8 * k;
        if (ch >= nchunk) continue;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = ch * 8 + i;
            o[i] = 0.f;
            if (c < C) {
                o[i] = (v[k][i] - mean) * rstd * gamma[c] + beta[c];
                y[base + (int64_t)c * T] = o[i] * m;
            }
        }
        if (y_act) {
            uint4 pk;
            pk.x = ln_pack2(o[0] * mk, o[1] * mk, fp16 & 1); pk.y = ln_pack2(o[2] * mk, o[3] * mk, fp16 & 1);
            pk.z = ln_pack2(o[4] * mk, o[5] * mk, fp16 & 1); pk.w = ln_pack2(o[6] * mk, o[7] * mk, fp16 & 1);
            if (fp16 & 2) {   // RVC_DT_FP16_SPLIT: planes [hi | lo] of 2 * nchunk chunks per batch item, lo = fp16(value - hi)
                *reinterpret_cast<uint4*>(y_act + (((size_t)b * 2 * nchunk + ch) * Lp + 64 + t) * 16) = pk;
                *reinterpret_cast<uint4*>(y_act + (((size_t)b * 2 * nchunk + nchunk + ch) * Lp + 64 + t) * 16) =
                    make_uint4(ln_lo2(o[0] * mk, o[1] * mk, pk.x), ln_lo2(o[2] * mk, o[3] * mk, pk.y), ln_lo2(o[4] * mk, o[5] * mk, pk.z), ln_lo2(o[6] * mk, o[7] * mk, pk.w));
            } else
            *reinterpret_cast<uint4*>(y_act + (((size_t)b * nchunk + ch) * Lp + 64 + t) * 16) = pk;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// relative-position attention, flash-style fp32.  CTA = 128 threads, 32 queries of one (b, head); key tiles of 32.
// thread (row = tid/4, sub = tid%4): scores for keys sub+4*jj (jj<8); output dims sub+4*dd.
// ------------------------------------------------------------------------------------------------
constexpr int ATT_BQ = 32, ATT_BK = 32, ATT_MAXDD = 32;  // dk <= 128

__global__ void __launch_bounds__(128) rel_attention_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                const float* __restrict__ emb_k, const float* __restrict__ emb_v,
                                                                float* __restrict__ out, int H, int n_heads, int T, int W) {
    extern __shared__ float smem[];
    const int dk = H / n_heads;
    const int DKP = dk + 1;
    const int NR = 2 * W + 1;
    float* q_s = smem;                        // [32][DKP]
    float* k_s = q_s + ATT_BQ * DKP;          // [32][DKP]
    float* v_s = k_s + ATT_BK * DKP;          // [32][DKP]
    float* p_s = v_s + ATT_BK * DKP;          // [32][33]
    float* qe_s = p_s + ATT_BQ * 33;          // [32][NR]
    float* ev_s = qe_s + ATT_BQ * NR;         // [NR][DKP]
    float* mk_s = ev_s + NR * DKP;            // [32] key mask

    const int tid = threadIdx.x, row = tid >> 2, sub = tid & 3;
    const int i0 = blockIdx.x * ATT_BQ;
    const int h = blockIdx.y % n_heads, b = blockIdx.y / n_heads;
    const float* qb = qkv + ((int64_t)b * 3 * H + h * dk) * T;
    const float* kb = qb + (int64_t)H * T;
    const float* vb = kb + (int64_t)H * T;
    const float* mb = mask + (int64_t)b * T;
    const float rs = sqrtf((float)dk);

    for (int idx = tid; idx < ATT_BQ * dk; idx += 128) {
        const int i = idx & 31, dd = idx >> 5;
        q_s[i * DKP + dd] = (i0 + i < T) ? qb[(int64_t)dd * T + i0 + i] / rs : 0.f;
    }
    for (int idx = tid; idx < NR * dk; idx += 128) {
        const int dd = idx % dk, rr = idx / dk;
        ev_s[rr * DKP + dd] = emb_v[rr * dk + dd];
    }
    __syncthreads();
    for (int idx = tid; idx < ATT_BQ * NR; idx += 128) {
        const int i = idx / NR, rr = idx - i * NR;
        float s = 0.f;
        for (int dd = 0; dd < dk; ++dd) s = fmaf(q_s[i * DKP + dd], emb_k[rr * dk + dd], s);
        qe_s[idx] = s;
    }
    const int qi = i0 + row;
    const float mq = (qi < T) ? mb[qi] : 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    float o[ATT_MAXDD];
#pragma unroll
    for (int dd = 0; dd < ATT_MAXDD; ++dd) o[dd] = 0.f;
    const int ndd = dk >> 2;

    for (int j0 = 0; j0 < T; j0 += ATT_BK) {
        __syncthreads();
        for (int idx = tid; idx < ATT_BK * dk; idx += 128) {
            const int j = idx & 31, dd = idx >> 5;
            const bool ok = j0 + j < T;
            k_s[j * DKP + dd] = ok ? kb[(int64_t)dd * T + j0 + j] : 0.f;
            v_s[j * DKP + dd] = ok ? vb[(int64_t)dd * T + j0 + j] : 0.f;
        }
        if (tid < ATT_BK) mk_s[tid] = (j0 + tid < T) ? mb[j0 + tid] : 0.f;
        __syncthreads();
        float s[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) s[jj] = 0.f;
        const float* qr = q_s + row * DKP;
        for (int dd = 0; dd < dk; ++dd) {
            const float qv = qr[dd];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) s[jj] = fmaf(qv, k_s[(sub + 4 * jj) * DKP + dd], s[jj]);
        }
        float tmax = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int jl = sub + 4 * jj, j = j0 + jl;
            const int rel = j - qi + W;
            if (rel >= 0 && rel < NR) s[jj] += qe_s[row * NR + rel];
            if (mq * mk_s[jl] == 0.f) s[jj] = -1e4f;
            if (j >= T) s[jj] = -INFINITY;
            tmax = fmaxf(tmax, s[jj]);
        }
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
        const float m_new = fmaxf(m_run, tmax);
        const float corr = expf(m_run - m_new);  // exp(-inf) = 0 on the first tile
        float psum = 0.f;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const float pv = expf(s[jj] - m_new);
            p_s[row * 33 + sub + 4 * jj] = pv;
            psum += pv;
        }
        psum += __shfl_xor_sync(0xffffffffu, psum, 1);
        psum += __shfl_xor_sync(0xffffffffu, psum, 2);
        l_run = l_run * corr + psum;
        m_run = m_new;
        __syncwarp();
#pragma unroll
        for (int dd = 0; dd < ATT_MAXDD; ++dd) o[dd] *= corr;
        const bool band = (j0 + ATT_BK - 1 >= i0 - W) && (j0 <= i0 + ATT_BQ - 1 + W);
        for (int jl = 0; jl < ATT_BK; ++jl) {
            const float pv = p_s[row * 33 + jl];
            const float* vr = v_s + jl * DKP + sub;
            const int rel = j0 + jl - qi + W;
            if (band && rel >= 0 && rel < NR) {
                const float* er = ev_s + rel * DKP + sub;
#pragma unroll
                for (int dd = 0; dd < ATT_MAXDD; ++dd)
                    if (dd < ndd) o[dd] = fmaf(pv, vr[4 * dd] + er[4 * dd], o[dd]);
            } else {
#pragma unroll
                for (int dd = 0; dd < ATT_MAXDD; ++dd)
                    if (dd < ndd) o[dd] = fmaf(pv, vr[4 * dd], o[dd]);
            }
        }
    }
    if (qi < T) {
        const float inv = 1.f / l_run;
        float* ob = out + ((int64_t)b * H + h * dk) * T + qi;
#pragma unroll
        for (int dd = 0; dd < ATT_MAXDD; ++dd)
            if (dd < ndd) ob[(int64_t)(sub + 4 * dd) * T] = o[dd] * inv;
    }
}

__global__ void sample_zp_f32_kernel(const float* __restrict__ stats, const float* __restrict__ eps, const float* __restrict__ mask,
                                     float* __restrict__ zp, int C, int T, int64_t n) {
    rvc::pdl_prologue();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = (int)(i % T);
    const int64_t bc = i / T;
    const int c = (int)(bc % C);
    const int64_t b = bc / C;
    const float m = stats[(b * 2 * C + c) * T + t];
    const float lg = stats[(b * 2 * C + C + c) * T + t];
    zp[i] = (m + expf(lg) * eps[i] * 0.66666f) * mask[b * T + t];
}



// y[b][co] = bias[co] + sum_ci w[ci][co] * x[b][ci]   (w packed [Cin][Cout]); block = 32 output columns x 8 ci slices
__global__ void __launch_bounds__(256) cond_gemv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ y, int Cin, int Cout) {
    rvc::pdl_prologue();
    __shared__ float red[8][33];
    const int b = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int co = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (co < Cout)
        for (int ci = ty; ci < Cin; ci += 8) a = fmaf(w[(int64_t)ci * Cout + co], x[(int64_t)b * Cin + ci], a);
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && co < Cout) {
        float s = bias ? bias[co] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][tx];
        y[(int64_t)b * Cout + co] = s;
    }
}

}  // namespace

RVC_API int rvc_cond_gemv_f32(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Cout, void* stream) {
    if (!x || !w || !y) return rvc::fail(RVC_EINVAL, "rvc_cond_gemv_f32: null pointer");
    if (B <= 0 || B > 65535 || Cin <= 0 || Cin > 8192 || Cout <= 0) return rvc::fail(RVC_ESHAPE, "rvc_cond_gemv_f32: bad shape");
    dim3 grid(rvc::cdiv(Cout, 32), B);
    rvc::launch_pdl(cond_gemv_f32_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, x, w, bias, y, Cin, Cout);
    return rvc::check_launch("rvc_cond_gemv_f32");
}

namespace {
}  // namespace

namespace {
// TextEncoder front end for the tensor-core path: mask[b,t] = t < len[b]; pb[b,c,t] = E[pitch[b,t]][c] * scale (exact gather)
__global__ void __launch_bounds__(256) embed_prep_kernel(const int64_t* __restrict__ pitch, const int64_t* __restrict__ lengths,
                                                         const float* __restrict__ emb_pitch, float scale, float* __restrict__ pb,
                                                         float* __restrict__ mask, int T, int H) {
    rvc::pdl_prologue();
    const int t = blockIdx.x * 256 + threadIdx.x, b = blockIdx.z;
    if (t >= T) return;
    if (blockIdx.y == 0) mask[(int64_t)b * T + t] = t < lengths[b] ? 1.f : 0.f;
    if (!pitch) return;
    const int64_t pi = pitch[(int64_t)b * T + t];
    const int c0 = blockIdx.y * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (c0 + i < H) pb[((int64_t)b * H + c0 + i) * T + t] = emb_pitch[pi * H + c0 + i] * scale;
}

// channels-last fp32 [B, T, D] -> packed fp16/bf16 [B][D/8][Lp][8]
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int T, int D, int Lp, int fp16) {
    rvc::pdl_prologue();
    const int t = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    const float4* xp = reinterpret_cast<const float4*>(x + ((int64_t)b * T + t) * D + c * 8);
    const float4 a = xp[0], d = xp[1];
    uint4 pk;
    pk.x = ln_pack2(a.x, a.y, fp16 & 1); pk.y = ln_pack2(a.z, a.w, fp16 & 1);
    pk.z = ln_pack2(d.x, d.y, fp16 & 1); pk.w = ln_pack2(d.z, d.w, fp16 & 1);
    if (fp16 & 2) {   // RVC_DT_FP16_SPLIT: [hi | lo] planes
        *reinterpret_cast<uint4*>(y + (((int64_t)b * 2 * (D / 8) + c) * Lp + 64 + t) * 16) = pk;
        *reinterpret_cast<uint4*>(y + (((int64_t)b * 2 * (D / 8) + D / 8 + c) * Lp + 64 + t) * 16) =
            make_uint4(ln_lo2(a.x, a.y, pk.x), ln_lo2(a.z, a.w, pk.y), ln_lo2(d.x, d.y, pk.z), ln_lo2(d.z, d.w, pk.w));
        return;
    }
    *reinterpret_cast<uint4*>(y + (((int64_t)b * (D / 8) + c) * Lp + 64 + t) * 16) = pk;
}
}  // namespace

RVC_API int rvc_embed_prep(const int64_t* pitch, const int64_t* lengths, const float* emb_pitch, float scale, float* pb, float* mask,
                           int32_t B, int32_t T, int32_t H, void* stream) {
    if (!lengths || !mask || (pitch && (!emb_pitch || !pb))) return rvc::fail(RVC_EINVAL, "rvc_embed_prep: null pointer");
    if (B <= 0 || B > 65535 || T < 0 || H <= 0) return rvc::fail(RVC_ESHAPE, "rvc_embed_prep: bad shape");
    if (T == 0) return 0;
    dim3 grid(rvc::cdiv(T, 256), pitch ? rvc::cdiv(H, 16) : 1, B);
    rvc::launch_pdl(embed_prep_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, pitch, lengths, emb_pitch, scale, pb, mask, T, H);
    return rvc::check_launch("rvc_embed_prep");
}

RVC_API int rvc_pack_rows(const float* x, void* y, int32_t B, int32_t T, int32_t D, int32_t dtype, void* stream) {
    if (!x || !y) return rvc::fail(RVC_EINVAL, "rvc_pack_rows: null pointer");
    if (dtype != RVC_DT_BF16 && dtype != RVC_DT_FP16 && dtype != RVC_DT_FP16_SPLIT)
        return rvc::fail(RVC_EINVAL, "rvc_pack_rows: dtype must be RVC_DT_BF16, RVC_DT_FP16 or RVC_DT_FP16_SPLIT");
    if (B <= 0 || B > 65535 || T < 0 || D <= 0 || D % 8 || (reinterpret_cast<uintptr_t>(x) & 15)) return rvc::fail(RVC_ESHAPE, "rvc_pack_rows: bad shape / alignment");
    if (T == 0) return 0;
    dim3 grid(rvc::cdiv(T, 256), D / 8, B);
    rvc::launch_pdl(pack_rows_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, x, (uint8_t*)y, T, D, (int)rvc_packed_rows(T),
                    dtype == RVC_DT_FP16 ? 1 : dtype == RVC_DT_FP16_SPLIT ? 3 : 0);
    return rvc::check_launch("rvc_pack_rows");
}

RVC_API int rvc_embed_f32(const float* phone, const int64_t* pitch, const int64_t* lengths, const float* wT, const float* bias,
                             const float* emb_pitch, float* x, float* mask, int32_t B, int32_t T, int32_t D, int32_t H, void* stream) {
    if (!phone || !lengths || !wT || !bias || !x || !mask) return rvc::fail(RVC_EINVAL, "rvc_embed_f32: null pointer");
    if (pitch && !emb_pitch) return rvc::fail(RVC_EINVAL, "rvc_embed_f32: pitch without embedding table");
    if (B <= 0 || B > 65535 || T < 0 || D <= 0 || H <= 0) return rvc::fail(RVC_ESHAPE, "rvc_embed_f32: bad shape");
    if (T == 0) return 0;
    dim3 grid(rvc::cdiv(T, 32), rvc::cdiv(H, 64), B);
    embed_f32_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(phone, pitch, lengths, wT, bias, emb_pitch, x, mask, T, D, H, sqrtf((float)H));
    return rvc::check_launch("rvc_embed_f32");
}

RVC_API int rvc_layernorm_f32(const float* x, const float* r, const float* r2, const float* r3, const float* gamma, const float* beta,
                                 const float* mask, float* y, int32_t B, int32_t C, int32_t T, float eps, int32_t out_mask, void* y_act,
                                 int32_t act_dtype, void* stream) {
    if (!x || !gamma || !beta || !y || (out_mask && !mask)) return rvc::fail(RVC_EINVAL, "rvc_layernorm_f32: null pointer");
    if ((r2 == nullptr) != (r3 == nullptr) || (r2 && !r)) return rvc::fail(RVC_EINVAL, "rvc_layernorm_f32: r2 and r3 go together (and need r)");
    if (B <= 0 || B > 65535 || C <= 0 || T < 0) return rvc::fail(RVC_ESHAPE, "rvc_layernorm_f32: bad shape");
    if (y_act && (C % 8 || !mask || (act_dtype != RVC_DT_BF16 && act_dtype != RVC_DT_FP16 && act_dtype != RVC_DT_FP16_SPLIT)))
        return rvc::fail(RVC_EINVAL, "rvc_layernorm_f32: packed output needs C %% 8 == 0, a mask and a bf16/fp16 dtype");
    if (T == 0) return 0;
    dim3 grid(rvc::cdiv(T, 32), B);
    const int Lp = (int)rvc_packed_rows(T), f16 = act_dtype == RVC_DT_FP16 ? 1 : act_dtype == RVC_DT_FP16_SPLIT ? 3 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int nper = rvc::cdiv(C, 64);   // 8-channel chunks per thread
    if (nper == 1) rvc::launch_pdl(layernorm_reg_kernel<1>, dim3(grid), dim3(32, 8), (size_t)0, st, x, r, r2, r3, gamma, beta, mask, y, C, T, eps, out_mask, (uint8_t*)y_act, Lp, f16);
    else if (nper == 2) rvc::launch_pdl(layernorm_reg_kernel<2>, dim3(grid), dim3(32, 8), (size_t)0, st, x, r, r2, r3, gamma, beta, mask, y, C, T, eps, out_mask, (uint8_t*)y_act, Lp, f16);
    else if (nper == 3) rvc::launch_pdl(layernorm_reg_kernel<3>, dim3(grid), dim3(32, 8), (size_t)0, st, x, r, r2, r3, gamma, beta, mask, y, C, T, eps, out_mask, (uint8_t*)y_act, Lp, f16);
    else if (nper == 4) rvc::launch_pdl(layernorm_reg_kernel<4>, dim3(grid), dim3(32, 8), (size_t)0, st, x, r, r2, r3, gamma, beta, mask, y, C, T, eps, out_mask, (uint8_t*)y_act, Lp, f16);
    else rvc::launch_pdl(layernorm_f32_kernel, dim3(grid), dim3(32, 8), (size_t)0, st, x, r, r2, r3, gamma, beta, mask, y, C, T, eps, out_mask, (uint8_t*)y_act, Lp, f16);
    return rvc::check_launch("rvc_layernorm_f32");
}

RVC_API int rvc_rel_attention_f32(const float* qkv, const float* mask, const float* emb_k, const float* emb_v, float* out,
                                     int32_t B, int32_t H, int32_t n_heads, int32_t T, int32_t window, void* stream) {
    if (!qkv || !mask || !emb_k || !emb_v || !out) return rvc::fail(RVC_EINVAL, "rvc_rel_attention_f32: null pointer");
    if (B <= 0 || H <= 0 || n_heads <= 0 || H % n_heads || T < 0 || window < 0) return rvc::fail(RVC_ESHAPE, "rvc_rel_attention_f32: bad shape");
    const int dk = H / n_heads;
    if (dk % 4 || dk > 4 * ATT_MAXDD) return rvc::fail(RVC_ESHAPE, "rvc_rel_attention_f32: head dim %d must be a multiple of 4 and <= 128", dk);
    if ((int64_t)B * n_heads > 65535) return rvc::fail(RVC_ESHAPE, "rvc_rel_attention_f32: B*heads > 65535");
    if (T == 0) return 0;
    const int DKP = dk + 1, NR = 2 * window + 1;
    const size_t smem = (size_t)(3 * 32 * DKP + 32 * 33 + 32 * NR + NR * DKP + 32) * sizeof(float);
    if (smem > 200 * 1024) return rvc::fail(RVC_ESHAPE, "rvc_rel_attention_f32: window too large");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rel_attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_rel_attention_f32: smem opt-in: %s", cudaGetErrorString(e));
    }
    dim3 grid(rvc::cdiv(T, ATT_BQ), B * n_heads);
    rel_attention_f32_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(qkv, mask, emb_k, emb_v, out, H, n_heads, T, window);
    return rvc::check_launch("rvc_rel_attention_f32");
}

RVC_API int rvc_sample_zp_f32(const float* stats, const float* eps, const float* mask, float* zp, int32_t B, int32_t C, int32_t T,
                                 void* stream) {
    if (!stats || !eps || !mask || !zp) return rvc::fail(RVC_EINVAL, "rvc_sample_zp_f32: null pointer");
    if (B <= 0 || C <= 0 || T < 0) return rvc::fail(RVC_ESHAPE, "rvc_sample_zp_f32: bad shape");
    const int64_t n = (int64_t)B * C * T;
    if (n == 0) return 0;
    rvc::launch_pdl(sample_zp_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, stats, eps, mask, zp, C, T, n);
    return rvc::check_launch("rvc_sample_zp_f32");
}
