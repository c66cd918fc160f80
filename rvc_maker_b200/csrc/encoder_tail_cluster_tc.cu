// The position-local tail of an encoder layer (Encoder.forward, synthesizers.py:340-348) as ONE cluster kernel:
//   x1 = LayerNorm1(x + conv_o(attn_out));   x2 = LayerNorm2(x1 + FFN(x1));   FFN(x) = conv_2(relu(conv_1(x * mask)) * mask) * mask  (k = 3, :302-315)
// Per-launch path: conv_o, LayerNorm, conv_1, conv_2 (three input-channel slices), LayerNorm = 5 launches, ~34 us per layer at T = 1000 for ~3 us of
// tensor work.  Everything here is local in time except the +-1 frame of the two k = 3 convs, so a CLUSTER of six CTAs owns a tile of 128 frames
// (124 kept) and nothing but the layer's input and output touches global memory:
//   * conv_o (split precision, three terms) N-split: CTA c computes output channels [32c, 32c + 32); + bias + residual x -> LayerNorm statistics:
//     every thread (one frame x 16 channels) sends its partial (sum, sum of squares) to all six CTAs through distributed shared memory;
//   * the normalised slice (fp16, masked) is broadcast into the six x1 operand tiles by the copy engine (as in flow_cluster_tc.cu);
//   * conv_1 N-split: CTA c computes hidden channels [128c, 128c + 128), relu, mask -> ITS OWN operand tile (no exchange: conv_2 is K-split);
//   * conv_2 K-split: CTA c contracts over its 128 hidden channels for all 192 outputs; the partial sums are reduce-scattered: the 32-channel slice k
//     of every CTA's partial goes to CTA k (st.shared::cluster, 16 KB per peer), which adds the six slices, bias, mask and the residual x1;
//   * LayerNorm2 the same way; x2 leaves as fp32 [B, 192, T] and as the masked packed [hi | lo] planes the next layer's qkv projection reads.
// Frames outside [0, T) and masked frames enter the convs as zeros (the reference masks the FFN input and zero-pads); rows whose halo is incomplete
// (the first / last two of a tile) are computed and discarded.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int CS = 6;
constexpr int H = 192;             // hidden channels
constexpr int FC = 768;            // FFN channels
constexpr int KW = 3;              // FFN kernel size
constexpr int BM = 128;
constexpr int HALO = 2;            // frames lost per side (one per k = 3 conv)
constexpr int KEEP = BM - 2 * HALO;
constexpr int TR = BM + 2;         // rows of the x1 / hidden operand tiles (one zero row each side)
constexpr int NSTAGE = 4;
constexpr int STAGE_BYTES = 16384;
constexpr int NG = 2;              // epilogue column groups
constexpr int NEPI = NG * 128;
constexpr int NTHREADS = 128 + NEPI;   // warp 0: loader + weight producer, warp 1: MMA issuer, warps 4-11: epilogue (lane quarter = warp % 4, group = (warp - 4) / 4)
constexpr int PADL = 64;
constexpr int HS = FC / CS;        // hidden channels per CTA (128)

constexpr int OT_BYTES = 2 * (H / 8) * BM * 16;        // attention output tile, [hi | lo] planes: 98304   (later: the reduce-scatter receive buffer)
constexpr int RECV_BYTES = CS * BM * 32 * 4;           // [src][row][32 floats]: 98304
constexpr int XT_BYTES = (H / 8) * TR * 16;            // x1 operand tile: 49920                            (later: the hidden slice, HS / 8 chunks)
constexpr int ST_BYTES = CS * NG * BM * 8;             // LayerNorm partial statistics [src][group][row] float2 (reused by LayerNorm2: a peer gets there
                                                       // only after this CTA's conv_2 slices, i.e. after this CTA has read the LayerNorm1 statistics)
static_assert(OT_BYTES >= RECV_BYTES && XT_BYTES >= (HS / 8) * TR * 16, "overlays");
constexpr int SMEM_BYTES = OT_BYTES + XT_BYTES + NSTAGE * STAGE_BYTES + ST_BYTES + 256;
constexpr int COL_O = 0, COL_F1 = 32, COL_F2 = 160, TMEM_COLS = 512;

struct TParams {
    const uint8_t* att; int64_t att_bs;          // packed [hi | lo] attention output: [b][2 * H/8][Lp][16 B]
    const float* x;                               // residual stream in, fp32 [B, H, T]
    float* y;                                     // ... out
    uint8_t* ya; int64_t ya_bs;                   // packed [hi | lo] planes of y * mask: [b][2 * H/8][Lp][16 B]
    const float* mask;
    const void* wo_hi; const void* wo_lo; const float* bo;      // conv_o: packed [1][H/8][H][8] planes
    const void* w1; const float* b1;                            // conv_1: packed [3][H/8][FC][8]
    const void* w2; const float* b2;                            // conv_2: packed [3][FC/8][H][8]
    const float* g1; const float* be1; const float* g2; const float* be2;
    int32_t B, T, Lp, ntiles, out_mask;
    float eps;
};

struct Chain {
    uint32_t a_addr, lbo_rows;
    int K, nch;
    const uint8_t* w;
    int wnch, wch0, wrows, row0, N, dcol, acc;   // weights: [K][wnch chunks][wrows][8]; this chain reads chunks [wch0, wch0 + nch), rows [row0, row0 + N)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
}
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack2h(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float2 h2f2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

__device__ __forceinline__ int n_chains(int ph) { return ph == 0 ? 3 : 1; }
__device__ __forceinline__ Chain get_chain(const TParams& p, int ph, int which, int c, uint32_t ot, uint32_t xt) {
    if (ph == 0) {   // conv_o, three split-precision terms: (hi, W_hi), (lo, W_hi), (hi, W_lo)
        const uint32_t a = ot + (which == 1 ? (uint32_t)(OT_BYTES / 2) : 0u);
        return Chain{a, BM, 1, H / 8, (const uint8_t*)(which == 2 ? p.wo_lo : p.wo_hi), H / 8, 0, H, 32 * c, 32, COL_O, which > 0};
    }
    if (ph == 1) return Chain{xt, TR, KW, H / 8, (const uint8_t*)p.w1, H / 8, 0, FC, HS * c, HS, COL_F1, 0};
    return Chain{xt, TR, KW, HS / 8, (const uint8_t*)p.w2, FC / 8, (HS / 8) * c, H, 0, H, COL_F2, 0};
}

__global__ void __launch_bounds__(NTHREADS, 1) encoder_tail_kernel(const TParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* ot_s = smem;                                   // attention output [hi | lo]; later recv[src][row][32]
    uint8_t* xt_s = ot_s + OT_BYTES;                        // x1 tile [H/8][TR][16]; later the hidden slice [HS/8][TR][16]
    uint8_t* w_smem = xt_s + XT_BYTES;
    float2* st_s = reinterpret_cast<float2*>(w_smem + NSTAGE * STAGE_BYTES);     // [CS][NG][BM]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(st_s) + ST_BYTES);
    const uint32_t bar0 = tc::smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    const uint32_t O_FULL = bar0 + 8u * (2 * NSTAGE);       // attention output tile loaded
    const uint32_t OPR = bar0 + 8u * (2 * NSTAGE + 1);      // operand of the next GEMM complete (x1: own slice + five peers; hidden: own)
    const uint32_t ACC_FULL = bar0 + 8u * (2 * NSTAGE + 2);
    const uint32_t STATS = bar0 + 8u * (2 * NSTAGE + 3);    // LayerNorm partial statistics of all CTAs / groups arrived (CS * 8 warps)
    const uint32_t RECV = bar0 + 8u * (2 * NSTAGE + 4);     // conv_2 partial slices of the five peers arrived ((CS - 1) * 8 warps)
    const uint32_t ACK = bar0 + 8u * (2 * NSTAGE + 5);      // the five peers have received this CTA's x1 slice (its source chunks may be overwritten)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 6);
    const uint32_t ot = tc::smem_u32(ot_s), xt = tc::smem_u32(xt_s);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int c = (int)cluster_ctarank();
    const int cl = (int)blockIdx.x / CS;
    const int b = cl / p.ntiles, tile = cl - b * p.ntiles;
    const int pos0 = tile * KEEP - HALO;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { tc::mbar_init(W_FULL(s), 1); tc::mbar_init(W_EMPTY(s), 1); }
        tc::mbar_init(O_FULL, 1);
        tc::mbar_init(OPR, 1);
        tc::mbar_init(ACC_FULL, 1);
        tc::mbar_init(STATS, CS * 2 * 4);
        tc::mbar_init(RECV, (CS - 1) * 2 * 4);
        tc::mbar_init(ACK, CS - 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
    // zero rows around the x1 / hidden tile (never written afterwards)
    for (int i = threadIdx.x; i < (H / 8) * 2; i += NTHREADS) {
        const int ch = i >> 1, row = (i & 1) ? TR - 1 : 0;
        *reinterpret_cast<uint4*>(xt_s + ((size_t)ch * TR + row) * 16) = make_uint4(0, 0, 0, 0);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    cluster_sync_all();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ======================= loader (attention output tile) + weight producer =======================
        {
            const int row0 = PADL + pos0;                                   // >= 62
            const int rows = min(BM, p.Lp - row0);                          // (the last tile may reach past the packed rows: those frames are >= T)
            if (lane == 0) tc::mbar_arrive_expect_tx(O_FULL, (uint32_t)(2 * (H / 8) * rows * 16));
            __syncwarp();
            const uint8_t* base = p.att + (size_t)b * p.att_bs;
            for (int u = lane; u < 2 * (H / 8); u += 32)
                tc::bulk_g2s(tc::smem_u32(ot_s + (size_t)u * BM * 16), base + ((size_t)u * p.Lp + row0) * 16, (uint32_t)rows * 16, O_FULL);
        }
        uint32_t wst = 0, wph = 0;
        for (int ph = 0; ph < 3; ++ph)
            for (int wc = 0; wc < n_chains(ph); ++wc) {
                const Chain ch = get_chain(p, ph, wc, c, ot, xt);
                const int ub = ch.N * 16, ups = (STAGE_BYTES / ub) & ~1;
                const int nflat = ch.K * ch.nch;
                for (int f0 = 0; f0 < nflat; f0 += ups) {
                    const int nu = min(ups, nflat - f0);
                    if (lane == 0) {
                        tc::mbar_wait(W_EMPTY(wst), wph ^ 1);
                        tc::mbar_arrive_expect_tx(W_FULL(wst), (uint32_t)nu * ub);
                    }
                    __syncwarp();
                    const uint32_t dst = tc::smem_u32(w_smem + wst * STAGE_BYTES);
                    for (int u = lane; u < nu; u += 32) {
                        const int f = f0 + u, tap = f / ch.nch, cc = f - tap * ch.nch;
                        tc::bulk_g2s(dst + u * ub, ch.w + (((size_t)tap * ch.wnch + ch.wch0 + cc) * ch.wrows + ch.row0) * 16, (uint32_t)ub, W_FULL(wst));
                    }
                    if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
                }
            }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        uint32_t wst = 0, wph = 0;
        for (int ph = 0; ph < 3; ++ph) {
            if (ph == 0) tc::mbar_wait(O_FULL, 0);
            else tc::mbar_wait(OPR, (uint32_t)((ph - 1) & 1));
            tc::fence_after_sync();
            if (ph == 1 && lane == 0) {      // the x1 tile is complete here, so every peer's slice has been read out of its source chunks: tell them
#pragma unroll
                for (int k = 1; k < CS; ++k) mbar_arrive_remote(mapa(ACK, (uint32_t)((c + k) % CS)));
            }
            __syncwarp();
            for (int wc = 0; wc < n_chains(ph); ++wc) {
                const Chain ch = get_chain(p, ph, wc, c, ot, xt);
                const uint32_t idesc = tc::make_idesc(0u, (uint32_t)ch.N);
                const uint32_t a_lo_hi = (ch.lbo_rows & 0x3FFFu) << 16, b_lo_hi = ((uint32_t)ch.N & 0x3FFFu) << 16;
                const uint32_t da_base = a_lo_hi | ((ch.a_addr >> 4) & 0x3FFFu);
                const uint32_t d_tmem = tmem_base + (uint32_t)ch.dcol;
                const int ups = (STAGE_BYTES / (ch.N * 16)) & ~1;
                const int nflat = ch.K * ch.nch;
                uint32_t accumulate = (uint32_t)ch.acc;
                int cch = 0, tap = 0;
                for (int f0 = 0; f0 < nflat; f0 += ups) {
                    const int nu = min(ups, nflat - f0);
                    tc::mbar_wait(W_FULL(wst), wph);
                    tc::fence_after_sync();
                    uint32_t db_lo = b_lo_hi | ((tc::smem_u32(w_smem + wst * STAGE_BYTES) >> 4) & 0x3FFFu);
                    for (int u = 0; u < nu; u += 2) {
                        const uint32_t da_lo = da_base + (uint32_t)tap + (uint32_t)cch * ch.lbo_rows;
                        if (tc::elect_one())
                            tc::umma<false>(d_tmem, ((uint64_t)desc_hi << 32) | da_lo, ((uint64_t)desc_hi << 32) | db_lo, idesc, accumulate);
                        accumulate = 1;
                        db_lo += 2u * (uint32_t)ch.N;
                        cch += 2;
                        if (cch == ch.nch) { cch = 0; ++tap; }
                    }
                    if (tc::elect_one()) tc::umma_commit(W_EMPTY(wst));
                    if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
                }
            }
            if (tc::elect_one()) tc::umma_commit(ACC_FULL);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ======================= epilogue: thread = frame pos0 + r, column group g =======================
        const int q = warp & 3, g = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const int et = threadIdx.x - 128;
        const int pos = pos0 + r;
        const bool inside = pos >= 0 && pos < p.T;
        const float mk = inside ? p.mask[(size_t)b * p.T + pos] : 0.f;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const int ch0 = 32 * c + 16 * g;                   // this thread's 16 channels of the hidden stream
        const float* xb = p.x + (size_t)b * H * p.T;
        float xs[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) xs[i] = inside ? xb[(size_t)(ch0 + i) * p.T + pos] : 0.f;

        // LayerNorm over the 192 channels of a frame: partial (sum, sum of squares) of 16 channels -> all six CTAs -> mean, rstd
        auto layer_norm = [&](float (&v)[16], int use, const float* gamma, const float* beta) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); }
            const uint32_t slot = tc::smem_u32(st_s + (c * NG + g) * BM + r);
#pragma unroll
            for (int k = 0; k < CS; ++k) st_cluster_v2(mapa(slot, (uint32_t)k), s1, s2);
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < CS; ++k) mbar_arrive_remote(mapa(STATS, (uint32_t)k));
            }
            mbar_wait_cluster(STATS, (uint32_t)use);
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int k = 0; k < CS * NG; ++k) {
                const float2 s = st_s[k * BM + r];
                t1 += s.x; t2 += s.y;
            }
            const float mean = t1 * (1.f / H);
            const float var = fmaxf(t2 * (1.f / H) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.eps);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (v[i] - mean) * rstd * __ldg(gamma + ch0 + i) + __ldg(beta + ch0 + i);
        };

        // ---- conv_o + bias + residual -> LayerNorm1 -> x1 (kept in registers), fp16(x1 * mask) broadcast into the six x1 tiles
        tc::mbar_wait(ACC_FULL, 0);
        tc::fence_after_sync();
        {
            uint32_t v[16];
            tc::tmem_ld<16>(t_row + COL_O + 16 * g, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) xs[i] += __uint_as_float(v[i]) + __ldg(p.bo + ch0 + i);
        }
        layer_norm(xs, 0, p.g1, p.be1);
        {
            uint4 pk[2];
#pragma unroll
            for (int u = 0; u < 2; ++u)
                pk[u] = mk != 0.f ? make_uint4(pack2h(xs[8 * u + 0], xs[8 * u + 1]), pack2h(xs[8 * u + 2], xs[8 * u + 3]),
                                               pack2h(xs[8 * u + 4], xs[8 * u + 5]), pack2h(xs[8 * u + 6], xs[8 * u + 7]))
                                  : make_uint4(0, 0, 0, 0);      // (a select, not a product: rows past the loaded part of the tile may hold NaN patterns)
            const int chunk0 = 4 * c + 2 * g;
#pragma unroll
            for (int u = 0; u < 2; ++u) *reinterpret_cast<uint4*>(xt_s + ((size_t)(chunk0 + u) * TR + r + 1) * 16) = pk[u];
            tc::fence_before_sync();
            tc::fence_proxy_async();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) {   // this CTA's four chunks (zero rows included) -> the five peers, by the copy engine
                const uint32_t bytes = 4u * TR * 16u, off = (uint32_t)(4 * c * TR * 16);
                tc::mbar_arrive_expect_tx(OPR, (CS - 1) * bytes);
#pragma unroll
                for (int k = 1; k < CS; ++k) {
                    const uint32_t peer = (uint32_t)((c + k) % CS);
                    bulk_s2peer(mapa(xt + off, peer), xt + off, bytes, mapa(OPR, peer));
                }
            }
        }
        // ---- conv_1: relu(acc + b1) * mask -> this CTA's hidden slice (its own conv_2 operand; overlays the x1 tile, which the GEMM has finished reading
        //      and whose own chunks the peers have received)
        tc::mbar_wait(ACC_FULL, 1);
        tc::fence_after_sync();
        mbar_wait_cluster(ACK, 0);
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
            uint32_t v[32];
            tc::tmem_ld<32>(t_row + COL_F1 + 64 * g + 32 * cb, v);
            tc::tmem_ld_wait();
            const float* bb = p.b1 + HS * c + 64 * g + 32 * cb;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    w[i] = mk != 0.f ? pack2h(fmaxf(__uint_as_float(v[8 * u + 2 * i]) + __ldg(bb + 8 * u + 2 * i), 0.f),
                                              fmaxf(__uint_as_float(v[8 * u + 2 * i + 1]) + __ldg(bb + 8 * u + 2 * i + 1), 0.f)) : 0u;
                *reinterpret_cast<uint4*>(xt_s + ((size_t)(8 * g + 4 * cb + u) * TR + r + 1) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        tc::fence_before_sync();
        tc::fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) tc::mbar_arrive(OPR);
        // ---- conv_2 partial sums: slice k of this CTA's [128 x 192] partial goes to CTA k; this group handles the slices 3g .. 3g+2
        tc::mbar_wait(ACC_FULL, 0);
        tc::fence_after_sync();
        float own[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) own[i] = 0.f;
#pragma unroll 1
        for (int kk = 0; kk < 3; ++kk) {
            const int k = 3 * g + kk;
            uint32_t v[32];
            tc::tmem_ld<32>(t_row + COL_F2 + 32 * k, v);
            tc::tmem_ld_wait();
            if (k == c) continue;              // own slice: read again below by both groups (16 channels each)
            const uint32_t dst = mapa(tc::smem_u32(ot_s + ((size_t)c * BM + r) * 128), (uint32_t)k);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                st_cluster_v4(dst + 16 * u, make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]), __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3])));
        }
        {
            uint32_t v[16];
            tc::tmem_ld<16>(t_row + COL_F2 + 32 * c + 16 * g, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) own[i] = __uint_as_float(v[i]);
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = 1; k < CS; ++k) mbar_arrive_remote(mapa(RECV, (uint32_t)((c + k) % CS)));
        }
        mbar_wait_cluster(RECV, 0);
        // ---- y = conv_2 (sum of the six slices) + bias, masked, + residual x1 -> LayerNorm2 -> out
#pragma unroll
        for (int k = 0; k < CS; ++k) {
            if (k == c) continue;
            const float4* rp = reinterpret_cast<const float4*>(ot_s + ((size_t)k * BM + r) * 128 + 64 * g);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 t = rp[u];
                own[4 * u] += t.x; own[4 * u + 1] += t.y; own[4 * u + 2] += t.z; own[4 * u + 3] += t.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) xs[i] += mk != 0.f ? own[i] + __ldg(p.b2 + ch0 + i) : 0.f;
        layer_norm(xs, 1, p.g2, p.be2);
        if (r >= HALO && r < HALO + KEEP && pos < p.T) {
            const float m = p.out_mask ? mk : 1.f;
            float* yb = p.y + (size_t)b * H * p.T;
#pragma unroll
            for (int i = 0; i < 16; ++i) yb[(size_t)(ch0 + i) * p.T + pos] = xs[i] * m;
            uint8_t* ya = p.ya + (size_t)b * p.ya_bs + ((size_t)(4 * c + 2 * g) * p.Lp + PADL + pos) * 16;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float f0 = xs[8 * u + 2 * i] * mk, f1 = xs[8 * u + 2 * i + 1] * mk;
                    hw[i] = pack2h(f0, f1);
                    const float2 hf = h2f2(hw[i]);
                    lw[i] = pack2h(f0 - hf.x, f1 - hf.y);
                }
                *reinterpret_cast<uint4*>(ya + (size_t)u * p.Lp * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                *reinterpret_cast<uint4*>(ya + (size_t)(H / 8 + u) * p.Lp * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
    }

    tc::fence_before_sync();
    cluster_sync_all();
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

RVC_API int rvc_encoder_tail_tc(const rvc_encoder_tail_desc* dp, void* stream) {
    if (!dp) return rvc::fail(RVC_EINVAL, "rvc_encoder_tail_tc: null descriptor");
    const rvc_encoder_tail_desc& d = *dp;
    if (d.hidden != H || d.filter != FC || d.kernel_size != KW)
        return rvc::fail(RVC_ESHAPE, "rvc_encoder_tail_tc: built for hidden 192, filter 768, kernel 3 (got %d, %d, %d)", d.hidden, d.filter, d.kernel_size);
    if (d.B <= 0 || d.T <= 0 || (int64_t)d.B * ((d.T + KEEP - 1) / KEEP) * CS > 0x7fffffffLL) return rvc::fail(RVC_ESHAPE, "rvc_encoder_tail_tc: bad shape");
    if (!d.att || !d.x || !d.y || !d.ya || !d.mask || !d.wo_hi || !d.wo_lo || !d.bo || !d.w1 || !d.b1 || !d.w2 || !d.b2 || !d.gamma1 || !d.beta1 || !d.gamma2 || !d.beta2)
        return rvc::fail(RVC_EINVAL, "rvc_encoder_tail_tc: null pointer");
    if (((uintptr_t)d.att | (uintptr_t)d.ya | (uintptr_t)d.wo_hi | (uintptr_t)d.wo_lo | (uintptr_t)d.w1 | (uintptr_t)d.w2) & 15)
        return rvc::fail(RVC_EINVAL, "rvc_encoder_tail_tc: packed buffers need 16-byte alignment");
    TParams p;
    p.att = (const uint8_t*)d.att; p.att_bs = d.att_bs; p.x = d.x; p.y = d.y; p.ya = (uint8_t*)d.ya; p.ya_bs = d.ya_bs; p.mask = d.mask;
    p.wo_hi = d.wo_hi; p.wo_lo = d.wo_lo; p.bo = d.bo; p.w1 = d.w1; p.b1 = d.b1; p.w2 = d.w2; p.b2 = d.b2;
    p.g1 = d.gamma1; p.be1 = d.beta1; p.g2 = d.gamma2; p.be2 = d.beta2;
    p.B = d.B; p.T = d.T; p.Lp = (int)rvc_packed_rows(d.T); p.ntiles = (d.T + KEEP - 1) / KEEP; p.out_mask = d.out_mask ? 1 : 0; p.eps = d.eps;
    auto kern = encoder_tail_kernel;
    static bool opted[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_encoder_tail_tc: smem opt-in: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) opted[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(d.B * p.ntiles * CS));
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
    if (le != cudaSuccess) return rvc::fail((int)le, "rvc_encoder_tail_tc: launch: %s", cudaGetErrorString(le));
    return rvc::check_launch("rvc_encoder_tail_tc");
}
