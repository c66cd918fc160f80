// Softmax stage of the tensor-core attention path (MultiHeadAttention.attention, synthesizers.py:221-250).
//
// The two contractions run on the tcgen05 GEMM engine (conv1d_tc.cu): S = q.k^T as a 1x1 "conv" whose weights are the packed
// k activations (written channel-blocked, S[b][h][key/8][query][8], so a query's 8 keys of a chunk are one 32-byte load and a
// warp of queries reads 1 KB contiguous), and O = P.v with the v^T tile as weights.  This kernel sits between them:
//   s[i,j] = S[j][i] / sqrt(dk) + [|j-i| <= w] (q_i / sqrt(dk)).Ek[j-i+w];   s = -1e4 where mask_i * mask_j == 0
//   P[i,j] = softmax_j(s)            -> packed fp16 [b][h][j/8][64 + i][8]   (the A operand of the P.v GEMM)
//   R[i,:] = sum_{|j-i|<=w} P[i,j] Ev[j-i+w]                                  (fp32, the residual input of the P.v GEMM)
//   vw     = v^T packed over keys   -> fp16 [b][h][j/8][d][8]                (the weights of the P.v GEMM)
// Two variants: short sequences (T <= 1024) keep a row's scores in registers (16 queries x 32 key groups per CTA); long ones
// read S twice with 32-query tiles (fully coalesced 128-byte key rows).
#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace {

constexpr int AG = 32;

// one key chunk of S for one query: 8 consecutive keys = 32 contiguous bytes in the channel-blocked layout [key/8][query][8]
__device__ __forceinline__ void ld_chunk(const float* p, float (&v)[8]) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

__device__ __forceinline__ uint32_t h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Short-sequence variant (T <= 1024: a lane's share of a key row is CPT <= 4 chunks of 8 keys): ONE WARP PER QUERY.  The launch is
// latency-bound (2 M scores per layer at T = 1000), so the kernel is organised to have as few dependent steps as possible and no block-wide
// barrier after the table staging: all of a query's scores are loaded once (lane = key chunk modulo 32), max / sum are warp shuffles, the
// relative-key dot products q.Ek[r] run on lanes r < 2W+1 from a per-warp copy of q in shared memory, the 2W+1 band probabilities go
// through a per-warp shared array into the relative-value sum.  The eight warps of a block own eight consecutive queries, so the
// 32-byte score chunks / probability stores of a block are 256 contiguous bytes.  (The previous 16-query x 32-group block needed four
// block barriers and took 20 us per layer; this one takes ~5.)
constexpr int WQ = 8;            // queries (= warps) per block
template <int CPT>
__global__ void __launch_bounds__(WQ * 32, 3) attn_softmax_fast_kernel(const float* __restrict__ S, const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                    const float* __restrict__ emb_k, const float* __restrict__ emb_v, uint8_t* __restrict__ P,
                                                                    float* __restrict__ R, uint8_t* __restrict__ vw, int H, int nh, int T, int Tk, int Lp,
                                                                    int W, int nqb, int nvb, int BH) {
    rvc::pdl_prologue();
    extern __shared__ float sm[];
    constexpr int NTH = WQ * 32;
    const int dk = H / nh, NR = 2 * W + 1, dkp = dk + 1;
    const int nsm = nqb * BH;
    const bool is_pack = (int)blockIdx.x >= nsm;
    const int bh = is_pack ? ((int)blockIdx.x - nsm) / nvb : (int)blockIdx.x / nqb;
    const int bx = is_pack ? ((int)blockIdx.x - nsm) % nvb : (int)blockIdx.x % nqb;
    const int h = bh % nh, b = bh / nh;
    const int nchunk = Tk / 8;
    const int tid = threadIdx.x;
    const float* vb = qkv + ((int64_t)b * 3 * H + 2 * H + h * dk) * T;
    if (is_pack) {
        const int idx = bx * NTH + tid;
        if (idx >= dk * nchunk) return;
        const int d = idx / nchunk, c = idx - d * nchunk;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (8 * c + i < T) ? vb[(int64_t)d * T + 8 * c + i] : 0.f;
        *reinterpret_cast<uint4*>(vw + ((((int64_t)b * nh + h) * nchunk + c) * dk + d) * 16) = make_uint4(h2(v[0], v[1]), h2(v[2], v[3]), h2(v[4], v[5]), h2(v[6], v[7]));
        return;
    }
    float* ek_s = sm;                        // [NR][dk + 1]  (padded: lanes r read row r at the same column)
    float* ev_s = ek_s + NR * dkp;           // [NR][dk]
    float* wq_s = ev_s + NR * dk;            // [WQ][dk]      q_i / sqrt(dk)
    float* qe_s = wq_s + WQ * dk;            // [WQ][32]      q_i . Ek[r]
    float* pb_s = qe_s + WQ * 32;            // [WQ][32]      band probabilities p[i][i + r - W]
    const int w = tid >> 5, lane = tid & 31;
    const int i = bx * WQ + w;
    const bool qok = i < T;
    const float* qb = qkv + ((int64_t)b * 3 * H + h * dk) * T;
    const float* mb = mask + (int64_t)b * T;
    const float* Sb = S + ((int64_t)b * nh + h) * Tk * T;
    const float rs = sqrtf((float)dk), inv = 1.f / rs;

    // this query's raw scores and key masks first: their latency overlaps the table staging
    float sc[CPT][8];
    uint32_t kmbits = 0;        // bit 8k + u: key mask of key 8 * (lane + 32k) + u is non-zero
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = lane + 32 * k;
        const bool cok = qok && c < nchunk;
        if (cok) ld_chunk(Sb + ((int64_t)c * T + i) * 8, sc[k]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = 8 * c + u;
            const bool in = cok && j < T;
            if (!in) sc[k][u] = -INFINITY;
            if (in && mb[j] != 0.f) kmbits |= 1u << (8 * k + u);
        }
    }
    for (int idx = tid; idx < NR * dk; idx += NTH) {
        const int r = idx / dk, d = idx - r * dk;
        ek_s[r * dkp + d] = emb_k[idx];
        ev_s[idx] = emb_v[idx];
    }
    for (int d = lane; d < dk; d += 32) wq_s[w * dk + d] = qok ? qb[(int64_t)d * T + i] / rs : 0.f;
    if (lane < 32) pb_s[w * 32 + lane] = 0.f;
    __syncthreads();
    if (!qok) return;          // (no block barrier below)
    if (lane < NR) {
        float a0 = 0.f, a1 = 0.f;
        const float* e = ek_s + lane * dkp;
        const float* q = wq_s + w * dk;
        for (int d = 0; d + 1 < dk; d += 2) {
            a0 = fmaf(q[d], e[d], a0);
            a1 = fmaf(q[d + 1], e[d + 1], a1);
        }
        if (dk & 1) a0 = fmaf(q[dk - 1], e[dk - 1], a0);
        qe_s[w * 32 + lane] = a0 + a1;
    }
    __syncwarp();

    const float mq = mb[i];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = lane + 32 * k;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = 8 * c + u;
            if (sc[k][u] != -INFINITY) {           // same arithmetic as the long-sequence kernel's chunk_scores()
                float s = sc[k][u] * inv;
                const int rel = j - i + W;
                if (rel >= 0 && rel < NR) s += qe_s[w * 32 + rel];
                if (mq == 0.f || !((kmbits >> (8 * k + u)) & 1u)) s = -1e4f;
                sc[k][u] = s;
                m = fmaxf(m, s);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
#pragma unroll
    for (int k = 0; k < CPT; ++k)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            sc[k][u] = __expf(sc[k][u] - m);         // exp(-inf) = 0 for keys >= T
            l += sc[k][u];
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float rl = 1.f / l;

    uint8_t* Pb = P + (((int64_t)b * nh + h) * nchunk * Lp + 64 + i) * 16;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = lane + 32 * k;
        if (c < nchunk) {
            float p8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                p8[u] = sc[k][u] * rl;
                const int rel = 8 * c + u - i + W;
                if (rel >= 0 && rel < NR) pb_s[w * 32 + rel] = p8[u];
            }
            *reinterpret_cast<uint4*>(Pb + (int64_t)c * Lp * 16) = make_uint4(h2(p8[0], p8[1]), h2(p8[2], p8[3]), h2(p8[4], p8[5]), h2(p8[6], p8[7]));
        }
    }
    __syncwarp();
    // relative-value term: lane owns output dims lane, lane + 32, ...
    float* Rb = R + ((int64_t)b * H + h * dk) * T + i;
    for (int d = lane; d < dk; d += 32) {
        float acc = 0.f;
        for (int r = 0; r < NR; ++r) acc = fmaf(pb_s[w * 32 + r], ev_s[r * dk + d], acc);     // 0 where i + r - W falls outside [0, T)
        Rb[(int64_t)d * T] = acc;
    }
}

// Long-sequence variant (T > 1024, e.g. the 4000-frame chunks of a stream): S no longer fits in registers, so it is read twice
// (max / sum, then probabilities) — but with a warp = 32 consecutive queries, every key row is one fully coalesced 128-byte line,
// and the band probabilities for the relative-value term go through shared memory.  The generic kernel's 16-query tiles read
// 64-byte segments three times (0.81 ms per layer at B = 2, T = 4000, a quarter of that step).
constexpr int LQ = 32, LG = 16;   // queries per CTA (= lanes), key groups (= warps)
__global__ void __launch_bounds__(LQ * LG) attn_softmax_long_kernel(const float* __restrict__ S, const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                    const float* __restrict__ emb_k, const float* __restrict__ emb_v, uint8_t* __restrict__ P,
                                                                    float* __restrict__ R, uint8_t* __restrict__ vw, int H, int nh, int T, int Tk, int Lp,
                                                                    int W, int nqb, int nvb, int BH) {
    rvc::pdl_prologue();
    extern __shared__ float sm[];
    constexpr int NTH = LQ * LG;
    const int dk = H / nh, NR = 2 * W + 1;
    const int nsm = nqb * BH;
    const bool is_pack = (int)blockIdx.x >= nsm;
    const int bh = is_pack ? ((int)blockIdx.x - nsm) / nvb : (int)blockIdx.x / nqb;
    const int bx = is_pack ? ((int)blockIdx.x - nsm) % nvb : (int)blockIdx.x % nqb;
    const int h = bh % nh, b = bh / nh;
    const int nchunk = Tk / 8;
    const int tid = threadIdx.x;
    const float* vb = qkv + ((int64_t)b * 3 * H + 2 * H + h * dk) * T;
    if (is_pack) {
        const int idx = bx * NTH + tid;
        if (idx >= dk * nchunk) return;
        const int d = idx / nchunk, c = idx - d * nchunk;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (8 * c + i < T) ? vb[(int64_t)d * T + 8 * c + i] : 0.f;
        *reinterpret_cast<uint4*>(vw + ((((int64_t)b * nh + h) * nchunk + c) * dk + d) * 16) = make_uint4(h2(v[0], v[1]), h2(v[2], v[3]), h2(v[4], v[5]), h2(v[6], v[7]));
        return;
    }
    float* qe_s = sm;                    // [LQ][NR]
    float* ev_s = qe_s + LQ * NR;        // [NR][dk]
    float* m_s = ev_s + NR * dk;         // [LG][LQ]
    float* l_s = m_s + LG * LQ;          // [LG][LQ]
    float* q_s = l_s + LG * LQ;          // [dk][LQ]   q / sqrt(dk)
    float* ek_s = q_s + dk * LQ;         // [NR][dk]
    float* pb_s = ek_s + NR * dk;        // [LQ][NR]   band probabilities
    const int ti = tid % LQ, g = tid / LQ;
    const int i0 = bx * LQ, i = i0 + ti;
    const float* qb = qkv + ((int64_t)b * 3 * H + h * dk) * T;
    const float* mb = mask + (int64_t)b * T;
    const float* Sb = S + ((int64_t)b * nh + h) * Tk * T;
    const float rs = sqrtf((float)dk);
    const bool qok = i < T;

    for (int idx = tid; idx < NR * dk; idx += NTH) { ev_s[idx] = emb_v[idx]; ek_s[idx] = emb_k[idx]; }
    for (int idx = tid; idx < dk * LQ; idx += NTH) {
        const int q = idx % LQ, d = idx / LQ;
        q_s[idx] = (i0 + q < T) ? qb[(int64_t)d * T + i0 + q] / rs : 0.f;
    }
    for (int idx = tid; idx < LQ * NR; idx += NTH) pb_s[idx] = 0.f;
    __syncthreads();
    for (int idx = tid; idx < LQ * NR; idx += NTH) {
        const int q = idx % LQ, r = idx / LQ;
        float a0 = 0.f, a1 = 0.f;
        for (int d = 0; d + 1 < dk; d += 2) {
            a0 = fmaf(q_s[d * LQ + q], ek_s[r * dk + d], a0);
            a1 = fmaf(q_s[(d + 1) * LQ + q], ek_s[r * dk + d + 1], a1);
        }
        if (dk & 1) a0 = fmaf(q_s[(dk - 1) * LQ + q], ek_s[r * dk + dk - 1], a0);
        qe_s[q * NR + r] = a0 + a1;
    }
    __syncthreads();

    const float mq = qok ? mb[i] : 0.f;
    const float inv = 1.f / rs;
    // scores of one key chunk (8 keys): eight independent coalesced loads, then bias / mask (same arithmetic as score())
    auto chunk_scores = [&](int c, float (&s8)[8]) {
        float raw[8], km[8];
        if (qok) ld_chunk(Sb + ((int64_t)c * T + i) * 8, raw);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = 8 * c + u;
            if (!(qok && j < T)) raw[u] = 0.f;
            km[u] = j < T ? mb[j] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = 8 * c + u;
            float s = raw[u] * inv;
            const int rel = j - i + W;
            if (rel >= 0 && rel < NR) s += qe_s[ti * NR + rel];
            if (mq * km[u] == 0.f) s = -1e4f;
            s8[u] = (qok && j < T) ? s : -INFINITY;
        }
    };
    // ---- pass 1: online max / sum over this warp's key chunks
    float m = -INFINITY, l = 0.f;
    for (int c = g; c < nchunk; c += LG) {
        float s8[8];
        chunk_scores(c, s8);
        float cm = s8[0];
#pragma unroll
        for (int u = 1; u < 8; ++u) cm = fmaxf(cm, s8[u]);
        if (cm == -INFINITY) continue;
        const float mn = fmaxf(m, cm);
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += __expf(s8[u] - mn);
        l = l * __expf(m - mn) + acc;
        m = mn;
    }
    m_s[g * LQ + ti] = m;
    l_s[g * LQ + ti] = l;
    __syncthreads();
    float M = -INFINITY, Lsum = 0.f;
#pragma unroll
    for (int gg = 0; gg < LG; ++gg) M = fmaxf(M, m_s[gg * LQ + ti]);
#pragma unroll
    for (int gg = 0; gg < LG; ++gg) {
        const float mg = m_s[gg * LQ + ti];
        if (mg != -INFINITY) Lsum += l_s[gg * LQ + ti] * __expf(mg - M);
    }
    const float rl = qok ? 1.f / Lsum : 0.f;
    // ---- pass 2: probabilities, packed fp16 (+ band members to shared memory)
    uint8_t* Pb = P + (((int64_t)b * nh + h) * nchunk * Lp + 64 + i) * 16;
    for (int c = g; c < nchunk; c += LG) {
        float s8[8], p8[8];
        chunk_scores(c, s8);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            p8[u] = __expf(s8[u] - M) * rl;
            const int rel = 8 * c + u - i + W;
            if (qok && rel >= 0 && rel < NR) pb_s[ti * NR + rel] = p8[u];
        }
        if (qok) *reinterpret_cast<uint4*>(Pb + (int64_t)c * Lp * 16) = make_uint4(h2(p8[0], p8[1]), h2(p8[2], p8[3]), h2(p8[4], p8[5]), h2(p8[6], p8[7]));
    }
    __syncthreads();
    if (!qok) return;
    // ---- relative-value term: warp g owns output dims d = g, g + LG, ...
    float* Rb = R + ((int64_t)b * H + h * dk) * T + i;
    for (int d = g; d < dk; d += LG) {
        float acc = 0.f;
        for (int r = 0; r < NR; ++r) acc = fmaf(pb_s[ti * NR + r], ev_s[r * dk + d], acc);
        Rb[(int64_t)d * T] = acc;
    }
}

}  // namespace

RVC_API int rvc_attn_softmax(const float* S, const float* qkv, const float* mask, const float* emb_k, const float* emb_v, void* P, float* R,
                             void* vw, int32_t B, int32_t H, int32_t n_heads, int32_t T, int32_t Tk, int32_t window, void* stream) {
    if (!S || !qkv || !mask || !emb_k || !emb_v || !P || !R || !vw) return rvc::fail(RVC_EINVAL, "rvc_attn_softmax: null pointer");
    if (B <= 0 || H <= 0 || n_heads <= 0 || H % n_heads || T <= 0 || Tk < T || Tk % 8 || window < 0)
        return rvc::fail(RVC_ESHAPE, "rvc_attn_softmax: bad shape");
    const int dk = H / n_heads;
    if (dk % AG || dk / AG > 16) return rvc::fail(RVC_ESHAPE, "rvc_attn_softmax: head dim %d must be a multiple of 32 and <= 512", dk);
    if ((int64_t)B * n_heads > 65535) return rvc::fail(RVC_ESHAPE, "rvc_attn_softmax: B*heads > 65535");
    const int NR = 2 * window + 1;
    if (NR > 32) return rvc::fail(RVC_ESHAPE, "rvc_attn_softmax: window %d > 15", window);
    const size_t smem = (size_t)(NR * (dk + 1) + NR * dk + WQ * dk + 2 * WQ * 32) * sizeof(float);
    if (smem > 200 * 1024) return rvc::fail(RVC_ESHAPE, "rvc_attn_softmax: window too large");
    const int nqb = rvc::cdiv(T, WQ);
    const int nvb = rvc::cdiv(dk * (Tk / 8), WQ * 32);
    dim3 grid((unsigned)((nqb + nvb) * B * n_heads));
    const int cpt = rvc::cdiv(Tk / 8, 32);     // key chunks per lane
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return rvc::fail((int)e, "rvc_attn_softmax: smem opt-in: %s", cudaGetErrorString(e));
        }
        rvc::launch_pdl(kern, dim3(grid), dim3(WQ * 32), (size_t)(smem), (cudaStream_t)stream, S, qkv, mask, emb_k, emb_v, (uint8_t*)P, R, (uint8_t*)vw, H, n_heads, T, Tk,
                        (int)rvc_packed_rows(T), window, nqb, nvb, B * n_heads);
        return rvc::check_launch("rvc_attn_softmax");
    };
    if (cpt <= 2) return launch(attn_softmax_fast_kernel<2>);
    if (cpt <= 4) return launch(attn_softmax_fast_kernel<4>);
    {   // long sequences: 32-query tiles (its own grid and shared-memory size)
        const int nqb_l = rvc::cdiv(T, LQ);
        const int nvb_l = rvc::cdiv(dk * (Tk / 8), LQ * LG);
        const size_t smem_l = (size_t)(2 * LQ * NR + 2 * NR * dk + 2 * LG * LQ + dk * LQ) * sizeof(float);
        if (smem_l > 200 * 1024) return rvc::fail(RVC_ESHAPE, "rvc_attn_softmax: window too large");
        if (smem_l > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(attn_softmax_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
            if (e != cudaSuccess) return rvc::fail((int)e, "rvc_attn_softmax: smem opt-in: %s", cudaGetErrorString(e));
        }
        rvc::launch_pdl(attn_softmax_long_kernel, dim3((unsigned)((nqb_l + nvb_l) * B * n_heads)), dim3(LQ * LG), smem_l, (cudaStream_t)stream, S, qkv, mask, emb_k,
                        emb_v, (uint8_t*)P, R, (uint8_t*)vw, H, n_heads, T, Tk, (int)rvc_packed_rows(T), window, nqb_l, nvb_l, B * n_heads);
        return rvc::check_launch("rvc_attn_softmax");
    }
}
