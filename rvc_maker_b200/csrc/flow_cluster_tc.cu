// The whole reverse ResidualCouplingBlock (residuals.py:87-95: Flip, C3, Flip, C2, Flip, C1, Flip, C0 with the Flips folded into the
// weights; ResidualCouplingLayer :127-137; WaveNet modules.py:35-51) as ONE kernel that never leaves the chip between layers.
//
// At T = 1000 the flow is 32 dependent convs of 24-48 tiles each: as separate launches (or as a grid-barrier chain, conv1d_tc_chain.cu)
// every layer costs 8-9 us of dependent L2 round trips (publish, barrier, operand load, residual load) for 0.3-1.5 us of tensor work.
// Every layer is local in time except for the +-2 frames of the k = 5 convs, so here a CLUSTER of six CTAs owns a tile of 128 frames
// through all four couplings (12 k = 5 layers -> 24 frames of halo per side are recomputed, 80 frames per tile are kept):
//   * CTA c of the cluster owns a channel slice — 32 of the 192 hidden / gate / skip channels, 16 of the 96 coupling channels;
//   * every layer is one tcgen05 GEMM tile per CTA, D[128 frames, slice] = A[128, K] . W[slice, K]^T, whose A operand (all K channels) sits in
//     this CTA's shared memory in the packed core-matrix layout of conv1d_tc.cu (a conv tap = the same tile 16 B further down);
//   * the fp32 state lives in TMEM: the WaveNet's hidden stream h and the skip sum are ACCUMULATED by the res / skip GEMMs themselves
//     (biases are added when the epilogue reads them), the coupling halves live in the epilogue threads' registers;
//   * after a layer, the epilogue thread of frame r turns its slice into the next layer's fp16 operand and writes it into the operand
//     tiles of ALL six CTAs through distributed shared memory (st.shared::cluster), then arrives on every CTA's "operands ready"
//     mbarrier; the MMA warp of each CTA waits for the 24 arrivals (6 CTAs x 4 warps), fences the async proxy and issues the next layer.
//     No global memory between layers: weights stream through a ring (never blocked by the layer barrier), z_p is read once, z written once.
// Frames outside [0, T) and masked frames are written as ZERO operands at every layer (the reference multiplies by x_mask after every
// layer and zero-pads each conv), so garbage in TMEM rows of such frames never reaches a kept output.
// Not bit-identical to the per-layer launches (h + res is summed by the tensor core's fp32 accumulator instead of the epilogue), same
// operand rounding; parity against the oracle is what the tests check.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int CS = 6;              // CTAs per cluster = channel slices
constexpr int H = 192;             // hidden channels of the WaveNet
constexpr int HALF = 96;           // channels per coupling half
constexpr int NL = 3;              // WaveNet layers per coupling
constexpr int KW = 5;              // WaveNet kernel size
constexpr int NF = 4;              // couplings
constexpr int BM = 128;            // frames per tile (TMEM lanes)
constexpr int HALO = (KW - 1) / 2 * NL * NF;   // 24 frames per side
constexpr int KEEP = BM - 2 * HALO;            // 80 frames kept per tile
constexpr int HPAD = (KW - 1) / 2;             // zero rows around the h tile
constexpr int HROWS = BM + 2 * HPAD;
constexpr int NSTAGE = 6;             // weight ring: 96 KB (an in_layer slice is 120 KB: the stream is latency-bound, depth is what hides it)
constexpr int STAGE_BYTES = 16384;
constexpr int NTHREADS = 256;      // warp 0: weight producer, warp 1: MMA issuer, warps 4-7: epilogue (TMEM lane quarter = warp % 4)
constexpr int PADL = 64;

constexpr int XT_BYTES = (HALF / 8) * BM * 16;      // x0 operand tile
constexpr int HT_BYTES = (H / 8) * HROWS * 16;      // h operand tile (with the conv's zero rows)
constexpr int AT_BYTES = (H / 8) * BM * 16;         // gated activations, later the masked skip sum
constexpr int SMEM_BYTES = XT_BYTES + HT_BYTES + AT_BYTES + NSTAGE * STAGE_BYTES + 256;

constexpr int COL_G = 0, COL_H = 64, COL_S = 96, COL_P = 128, TMEM_COLS = 256;

struct FlowParams {
    const void* pre_w[NF];  const void* in_w[NF][NL];  const void* res_w[NF][NL - 1];  const void* skip_w[NF][NL];  const void* post_w[NF];
    const float* hb[NF][NL];     // cumulative bias of the hidden stream when layer l reads it: b_pre (+ b_res0 (+ b_res1))
    const float* in_b[NF][NL];   // gate-ordered bias of in_layers[l] (per 64 columns: 32 tanh, 32 sigmoid)
    const float* sb[NF];         // sum of the three skip biases
    const float* post_b[NF];
    const float* gc; int64_t gc_bs; int32_t gc_off[NF];    // speaker conditioning [B][...]: gate-ordered, 3 x 384 per coupling from gc_off[f]
    const float* zp; float* z; void* za; const float* mask;
    int32_t B, T, Lp, ntiles, post_rows, fp16_;
};

struct Chain {          // one accumulation chain of a layer: D[:, dcol : dcol + N] (+)= A . W[row0 : row0 + N]^T
    uint32_t a_addr;    // shared-memory byte address of the operand tile
    uint32_t lbo_rows;  // rows per 8-channel chunk of that tile
    int K, nch;         // taps, 8-channel chunks
    const uint8_t* w;   // packed weights [K][nch][wrows][8]
    int wrows, row0, N, dcol, acc;
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
// bulk copy own shared memory -> a peer CTA's shared memory (async proxy), completing `bytes` on the PEER's mbarrier
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
__device__ __forceinline__ float gate_fast(float t, float s) {   // conv1d_tc.cu
    const float e2 = __expf(2.f * fminf(fmaxf(t, -15.f), 15.f));
    const float es = __expf(-fmaxf(s, -30.f));
    return __fdividef(e2 - 1.f, (e2 + 1.f) * (1.f + es));
}

// the layer program: 8 phases per coupling — PRE, IN0, RS0, IN1, RS1, IN2, RS2, POST — in the reverse pass's order f = 3, 2, 1, 0
__device__ __forceinline__ int n_chains(int t) { return (t == 2 || t == 4) ? 2 : 1; }
__device__ __forceinline__ Chain get_chain(const FlowParams& p, int f, int t, int which, int c, uint32_t xt, uint32_t ht, uint32_t at) {
    Chain ch;
    if (t == 0) {
        ch = Chain{xt, BM, 1, HALF / 8, (const uint8_t*)p.pre_w[f], H, 32 * c, 32, COL_H, 0};
    } else if (t == 7) {
        ch = Chain{at, BM, 1, H / 8, (const uint8_t*)p.post_w[f], p.post_rows, 16 * c, 16, COL_P, 0};
    } else if (t & 1) {
        ch = Chain{ht, HROWS, KW, H / 8, (const uint8_t*)p.in_w[f][(t - 1) >> 1], 2 * H, 64 * c, 64, COL_G, 0};
    } else {
        const int l = (t >> 1) - 1;
        if (l < NL - 1 && which == 0) ch = Chain{at, BM, 1, H / 8, (const uint8_t*)p.res_w[f][l], H, 32 * c, 32, COL_H, 1};
        else ch = Chain{at, BM, 1, H / 8, (const uint8_t*)p.skip_w[f][l], H, 32 * c, 32, COL_S, l > 0};
    }
    return ch;
}

__global__ void __launch_bounds__(NTHREADS, 1) flow_cluster_kernel(const FlowParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* xt_s = smem;
    uint8_t* ht_s = xt_s + XT_BYTES;
    uint8_t* at_s = ht_s + HT_BYTES;
    uint8_t* w_smem = at_s + AT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + NSTAGE * STAGE_BYTES);
    const uint32_t bar0 = tc::smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    const uint32_t OPR = bar0 + 8u * (2 * NSTAGE);          // operands of the next layer are complete: own slice written + the five peers' slices landed
    const uint32_t ACC_FULL = bar0 + 8u * (2 * NSTAGE + 1);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2);
    const uint32_t xt = tc::smem_u32(xt_s), ht = tc::smem_u32(ht_s), at = tc::smem_u32(at_s);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int c = (int)cluster_ctarank();
    const int cl = (int)blockIdx.x / CS;              // cluster id -> (batch item, tile)
    const int b = cl / p.ntiles, tile = cl - b * p.ntiles;
    const int pos0 = tile * KEEP - HALO;              // frame of tile row 0

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { tc::mbar_init(W_FULL(s), 1); tc::mbar_init(W_EMPTY(s), 1); }
        tc::mbar_init(OPR, 1);
        tc::mbar_init(ACC_FULL, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
    // the conv's zero rows around the h tile (never written afterwards)
    for (int i = threadIdx.x; i < (H / 8) * 2 * HPAD; i += NTHREADS) {
        const int ch = i / (2 * HPAD), rr = i - ch * 2 * HPAD;
        const int row = rr < HPAD ? rr : BM + rr;
        *reinterpret_cast<uint4*>(ht_s + ((size_t)ch * HROWS + row) * 16) = make_uint4(0, 0, 0, 0);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    cluster_sync_all();                                  // every CTA's barriers exist before anyone arrives on them remotely
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ======================= weight producer: all 32 layers' slices through one ring, never blocked by a layer barrier =======================
        uint32_t wst = 0, wph = 0;
        for (int ph = 0; ph < NF * 8; ++ph) {
            const int f = NF - 1 - (ph >> 3), t = ph & 7;
            for (int wc = 0; wc < n_chains(t); ++wc) {
                const Chain ch = get_chain(p, f, t, wc, c, xt, ht, at);
                const int ub = ch.N * 16, ups = STAGE_BYTES / ub;     // bytes per unit (one tap x one 8-channel chunk), units per stage
                const int nflat = ch.K * ch.nch;
                for (int f0 = 0; f0 < nflat; f0 += ups) {
                    const int nu = min(ups, nflat - f0);
                    if (lane == 0) {
                        tc::mbar_wait(W_EMPTY(wst), wph ^ 1);
                        tc::mbar_arrive_expect_tx(W_FULL(wst), (uint32_t)nu * ub);
                    }
                    __syncwarp();
                    const uint32_t dst = tc::smem_u32(w_smem + wst * STAGE_BYTES);
                    for (int u = lane; u < nu; u += 32)
                        tc::bulk_g2s(dst + u * ub, ch.w + ((size_t)(f0 + u) * ch.wrows + ch.row0) * 16, (uint32_t)ub, W_FULL(wst));
                    if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);            // SBO = 128 B, no swizzle
        uint32_t wst = 0, wph = 0;
        for (int ph = 0; ph < NF * 8; ++ph) {
            const int f = NF - 1 - (ph >> 3), t = ph & 7;
            tc::mbar_wait(OPR, (uint32_t)(ph & 1));                     // this layer's operands: own slice + five slices copied in by the peers
            tc::fence_after_sync();
            for (int wc = 0; wc < n_chains(t); ++wc) {
                const Chain ch = get_chain(p, f, t, wc, c, xt, ht, at);
                const uint32_t idesc = tc::make_idesc(0u, (uint32_t)ch.N);
                const uint32_t a_lo_hi = (ch.lbo_rows & 0x3FFFu) << 16, b_lo_hi = ((uint32_t)ch.N & 0x3FFFu) << 16;
                const uint32_t da_base = a_lo_hi | ((ch.a_addr >> 4) & 0x3FFFu);
                const uint32_t d_tmem = tmem_base + (uint32_t)ch.dcol;
                const int ub16 = ch.N, ups = STAGE_BYTES / (ch.N * 16);
                const int nflat = ch.K * ch.nch;
                uint32_t accumulate = (uint32_t)ch.acc;
                int cch = 0, tap = 0;
                for (int f0 = 0; f0 < nflat; f0 += ups) {
                    const int nu = min(ups, nflat - f0);
                    tc::mbar_wait(W_FULL(wst), wph);
                    tc::fence_after_sync();
                    uint32_t db_lo = b_lo_hi | ((tc::smem_u32(w_smem + wst * STAGE_BYTES) >> 4) & 0x3FFFu);
                    for (int u = 0; u < nu; u += 2) {
                        const uint32_t da_lo = da_base + (uint32_t)tap + (uint32_t)cch * ch.lbo_rows;   // 16-byte units: tap = row shift, chunk stride = lbo
                        if (tc::elect_one())
                            tc::umma<false>(d_tmem, ((uint64_t)desc_hi << 32) | da_lo, ((uint64_t)desc_hi << 32) | db_lo, idesc, accumulate);
                        accumulate = 1;
                        db_lo += 2u * ub16;
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
        // ======================= epilogue: thread r owns frame pos0 + r =======================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int pos = pos0 + r;
        const bool inside = pos >= 0 && pos < p.T;
        const float mk = inside ? p.mask[(size_t)b * p.T + pos] : 0.f;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const int er = threadIdx.x - 128;      // 0 .. 127
        // Publish this CTA's channel slice of an operand tile: every thread stores its frame's 16-byte chunks into the LOCAL tile, the slice
        // (whole 8-channel chunks, contiguous) is then copied into the same place of the five peers' tiles by the copy engine
        // (cp.async.bulk shared::cta -> shared::cluster), which completes the bytes on the peer's OPR barrier; this CTA's own OPR phase
        // expects the five slices coming the other way.  (Per-thread st.shared::cluster stores to six CTAs cost the epilogue 24 stores
        // per thread and the SM's shared-memory port ~1.4 us per layer.)
        auto publish = [&](uint8_t* tile, uint32_t tile_addr, int rows, int row, int chunk0, const uint4* pk, int nchunk) {
            for (int u = 0; u < nchunk; ++u) *reinterpret_cast<uint4*>(tile + ((size_t)(chunk0 + u) * rows + row) * 16) = pk[u];
            tc::fence_proxy_async();                       // generic stores to OWN shared memory -> visible to the async proxy (bulk-copy source,
                                                           // tensor-core operand); the all-spaces form of the fence cost 0.9 us per layer here
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (er == 0) {
                const uint32_t bytes = (uint32_t)(nchunk * rows * 16), off = (uint32_t)(chunk0 * rows * 16);
                tc::mbar_arrive_expect_tx(OPR, (CS - 1) * bytes);
#pragma unroll
                for (int k = 1; k < CS; ++k) {
                    const uint32_t peer = (uint32_t)((c + k) % CS);
                    bulk_s2peer(mapa(tile_addr + off, peer), tile_addr + off, bytes, mapa(OPR, peer));
                }
            }
        };
        // the two coupling halves of this CTA's 16-channel slice: xa = the half read as x0 by the next coupling, xb = the half it updates
        float xa[16], xb[16];
        {
            const float* zb = p.zp + (size_t)b * 2 * HALF * p.T;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                xa[i] = inside ? zb[(size_t)(HALF + 16 * c + i) * p.T + pos] : 0.f;     // coupling 3 (flipped): x0 = second half
                xb[i] = inside ? zb[(size_t)(16 * c + i) * p.T + pos] : 0.f;
            }
            uint4 pk[2];
#pragma unroll
            for (int u = 0; u < 2; ++u)
                pk[u] = make_uint4(pack2h(xa[8 * u + 0], xa[8 * u + 1]), pack2h(xa[8 * u + 2], xa[8 * u + 3]), pack2h(xa[8 * u + 4], xa[8 * u + 5]),
                                   pack2h(xa[8 * u + 6], xa[8 * u + 7]));
            publish(xt_s, xt, BM, r, 2 * c, pk, 2);
        }
        const float* gcb = p.gc + (size_t)b * p.gc_bs;
        for (int ph = 0; ph < NF * 8; ++ph) {
            const int f = NF - 1 - (ph >> 3), t = ph & 7;
            tc::mbar_wait(ACC_FULL, (uint32_t)(ph & 1));
            tc::fence_after_sync();
            if (t == 7) {
                // ---- post: m = (acc + b) * mask;  x1 <- (x1 - m) * mask;  the updated half is the next coupling's x0
                uint32_t v[16];
                tc::tmem_ld<16>(t_row + COL_P, v);
                tc::tmem_ld_wait();
                const float* pb = p.post_b[f] + 16 * c;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float nx = (xb[i] - (__uint_as_float(v[i]) + __ldg(pb + i))) * mk;
                    xb[i] = xa[i];
                    xa[i] = nx;
                }
                if (ph + 1 < NF * 8) {
                    uint4 pk[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        pk[u] = make_uint4(pack2h(xa[8 * u + 0], xa[8 * u + 1]), pack2h(xa[8 * u + 2], xa[8 * u + 3]), pack2h(xa[8 * u + 4], xa[8 * u + 5]),
                                           pack2h(xa[8 * u + 6], xa[8 * u + 7]));
                    tc::fence_before_sync();
                    publish(xt_s, xt, BM, r, 2 * c, pk, 2);
                }
            } else if (t & 1) {
                // ---- in_layer: acts = tanh(a_t) * sigmoid(a_s), a = acc + b_in + cond  (commons.py:35-41); not masked (modules.py:45)
                const int l = (t - 1) >> 1;
                uint32_t vt[32], vs[32];
                tc::tmem_ld<32>(t_row + COL_G, vt);
                tc::tmem_ld<32>(t_row + COL_G + 32, vs);
                tc::tmem_ld_wait();
                const float4* ib = reinterpret_cast<const float4*>(p.in_b[f][l] + 64 * c);                       // 16 x (tanh) then 16 x (sigmoid)
                const float4* cb = reinterpret_cast<const float4*>(gcb + p.gc_off[f] + l * 2 * H + 64 * c);
                uint4 pk[4];
                uint32_t w[16];
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 bt = __ldg(ib + i4), bs_ = __ldg(ib + 8 + i4), ct = __ldg(cb + i4), cs_ = __ldg(cb + 8 + i4);
                    const float a0 = gate_fast(__uint_as_float(vt[4 * i4 + 0]) + bt.x + ct.x, __uint_as_float(vs[4 * i4 + 0]) + bs_.x + cs_.x);
                    const float a1 = gate_fast(__uint_as_float(vt[4 * i4 + 1]) + bt.y + ct.y, __uint_as_float(vs[4 * i4 + 1]) + bs_.y + cs_.y);
                    const float a2 = gate_fast(__uint_as_float(vt[4 * i4 + 2]) + bt.z + ct.z, __uint_as_float(vs[4 * i4 + 2]) + bs_.z + cs_.z);
                    const float a3 = gate_fast(__uint_as_float(vt[4 * i4 + 3]) + bt.w + ct.w, __uint_as_float(vs[4 * i4 + 3]) + bs_.w + cs_.w);
                    w[2 * i4] = pack2h(a0, a1);
                    w[2 * i4 + 1] = pack2h(a2, a3);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) pk[u] = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
                tc::fence_before_sync();
                publish(at_s, at, BM, r, 4 * c, pk, 4);
            } else {
                // ---- pre / res+skip: read the hidden stream (t = 0, 2, 4: the next in_layer's operand) or the finished skip sum (t = 6: post's operand)
                const bool skip = t == 6;
                const float* bias = (skip ? p.sb[f] : p.hb[f][t >> 1]) + 32 * c;
                uint32_t v[32];
                tc::tmem_ld<32>(t_row + (skip ? COL_S : COL_H), v);
                tc::tmem_ld_wait();
                uint4 pk[4];
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2)
                    w[i >> 1] = pack2h((__uint_as_float(v[i]) + __ldg(bias + i)) * mk, (__uint_as_float(v[i + 1]) + __ldg(bias + i + 1)) * mk);
#pragma unroll
                for (int u = 0; u < 4; ++u) pk[u] = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
                tc::fence_before_sync();
                if (skip) publish(at_s, at, BM, r, 4 * c, pk, 4);
                else publish(ht_s, ht, HROWS, r + HPAD, 4 * c, pk, 4);     // (whole chunks travel, zero rows included)
            }
        }
        // ---- result: z = [first half | second half], fp32 [B, 192, T] and (optionally) the masked packed copy conv_pre reads
        if (r >= HALO && r < HALO + KEEP && pos < p.T) {
            float* zb = p.z + (size_t)b * 2 * HALF * p.T;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                zb[(size_t)(16 * c + i) * p.T + pos] = xb[i];
                zb[(size_t)(HALF + 16 * c + i) * p.T + pos] = xa[i];
            }
            if (p.za) {
                uint8_t* za = reinterpret_cast<uint8_t*>(p.za) + ((size_t)b * (2 * HALF / 8) * p.Lp + PADL + pos) * 16;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    *reinterpret_cast<uint4*>(za + (size_t)(2 * c + u) * p.Lp * 16) =
                        make_uint4(pack2h(xb[8 * u + 0] * mk, xb[8 * u + 1] * mk), pack2h(xb[8 * u + 2] * mk, xb[8 * u + 3] * mk),
                                   pack2h(xb[8 * u + 4] * mk, xb[8 * u + 5] * mk), pack2h(xb[8 * u + 6] * mk, xb[8 * u + 7] * mk));
                    *reinterpret_cast<uint4*>(za + (size_t)(HALF / 8 + 2 * c + u) * p.Lp * 16) =
                        make_uint4(pack2h(xa[8 * u + 0] * mk, xa[8 * u + 1] * mk), pack2h(xa[8 * u + 2] * mk, xa[8 * u + 3] * mk),
                                   pack2h(xa[8 * u + 4] * mk, xa[8 * u + 5] * mk), pack2h(xa[8 * u + 6] * mk, xa[8 * u + 7] * mk));
                }
            }
        }
    }

    tc::fence_before_sync();
    cluster_sync_all();          // no CTA leaves while a peer may still write into its shared memory or arrive on its barriers
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

RVC_API int rvc_flow_reverse_tc(const rvc_flow_desc* dp, void* stream) {
    if (!dp) return rvc::fail(RVC_EINVAL, "rvc_flow_reverse_tc: null descriptor");
    const rvc_flow_desc& d = *dp;
    if (d.hidden != H || d.half != HALF || d.n_layers != NL || d.kernel_size != KW || d.n_flows != NF)
        return rvc::fail(RVC_ESHAPE, "rvc_flow_reverse_tc: built for hidden 192, 2 x 96 channels, 4 couplings of 3 k = 5 layers (got %d, %d, %d, %d, %d)", d.hidden,
                         d.half, d.n_flows, d.n_layers, d.kernel_size);
    if (d.B <= 0 || d.T <= 0 || (int64_t)d.B * ((d.T + KEEP - 1) / KEEP) * CS > 0x7fffffffLL) return rvc::fail(RVC_ESHAPE, "rvc_flow_reverse_tc: bad shape");
    if (!d.zp || !d.z || !d.mask || !d.gc) return rvc::fail(RVC_EINVAL, "rvc_flow_reverse_tc: null pointer");
    if (d.post_rows < HALF || d.post_rows % 16) return rvc::fail(RVC_ESHAPE, "rvc_flow_reverse_tc: post_rows");
    FlowParams p;
    for (int f = 0; f < NF; ++f) {
        const rvc_flow_coupling& cp = d.flows[f];
        if (!cp.pre_w || !cp.post_w || !cp.post_b || !cp.sb) return rvc::fail(RVC_EINVAL, "rvc_flow_reverse_tc: coupling %d has a null pointer", f);
        p.pre_w[f] = cp.pre_w; p.post_w[f] = cp.post_w; p.post_b[f] = cp.post_b; p.sb[f] = cp.sb; p.gc_off[f] = cp.gc_off;
        for (int l = 0; l < NL; ++l) {
            if (!cp.in_w[l] || !cp.in_b[l] || !cp.skip_w[l] || !cp.hb[l] || (l < NL - 1 && !cp.res_w[l])) return rvc::fail(RVC_EINVAL, "rvc_flow_reverse_tc: coupling %d layer %d has a null pointer", f, l);
            p.in_w[f][l] = cp.in_w[l]; p.in_b[f][l] = cp.in_b[l]; p.skip_w[f][l] = cp.skip_w[l]; p.hb[f][l] = cp.hb[l];
            if (l < NL - 1) p.res_w[f][l] = cp.res_w[l];
        }
    }
    p.gc = d.gc; p.gc_bs = d.gc_bs; p.zp = d.zp; p.z = d.z; p.za = d.za; p.mask = d.mask;
    p.B = d.B; p.T = d.T; p.Lp = (int)rvc_packed_rows(d.T); p.ntiles = (d.T + KEEP - 1) / KEEP; p.post_rows = d.post_rows; p.fp16_ = 1;
    auto kern = flow_cluster_kernel;
    static bool opted[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_flow_reverse_tc: smem opt-in: %s", cudaGetErrorString(e));
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
    if (le != cudaSuccess) return rvc::fail((int)le, "rvc_flow_reverse_tc: launch: %s", cudaGetErrorString(le));
    return rvc::check_launch("rvc_flow_reverse_tc");
}
