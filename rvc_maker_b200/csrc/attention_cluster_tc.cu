// Scores + softmax of the relative-position attention (MultiHeadAttention.attention, synthesizers.py:221-250) for T <= 1024 as ONE kernel:
//   s[i,j] = (q_i . k_j + [|j-i| <= w] q_i . Ek[j-i+w]) / sqrt(dk);   s = -1e4 where mask_i * mask_j == 0;   P = softmax_j(s)
// A CLUSTER of ceil(T / 256) <= 4 CTAs owns 128 queries of one (batch item, head); CTA kt of the cluster owns the keys [256 kt, 256 kt + 256)
// (clusters of eight 128-key CTAs did not all fit the GPU at once: 15 of the 16 clusters of a T = 1000 layer, i.e. two rounds):
//   * one tcgen05 tile S[128 queries, 256 keys] = Q . K^T per CTA, operands = the packed q / k activation planes the qkv projection wrote
//     (three split-precision terms hi.hi + lo.hi + hi.lo when the planes come in pairs), plus a 32-column tile Q . Ek^T for the banded relative-key term;
//   * the softmax runs on the accumulator with sixteen warps (thread = one query x 64 keys): scale, band term, mask, local max -> combined over
//     the four column groups in shared memory, then over the CTAs of the cluster through distributed shared memory; exp, sums the same way,
//     normalise; logits and exponentials are parked in TMEM between the passes (tcgen05.st), so a thread holds 32 values at a time;
//   * P leaves as the packed fp16 A operand of the P.v GEMM, [b][h][key / 8][row][8]; the 2w+1 band probabilities of a row are ALSO written as 32 extra
//     "key" columns (chunks Tk/8 .. Tk/8+3), whose matching weight rows are Ev^T: the relative-value term sum_r p[i, i+r-w] Ev[r] becomes four more
//     k-steps of the P.v GEMM instead of a separate fp32 pass and a residual input;
//   * the clusters of query tile 0 also write v^T (the P.v GEMM's per-batch weights) for their key tile.
// Replaces the S = q.k^T launch (scores through HBM: 8 MB per layer at T = 1000), the softmax launch and the R buffer of the per-launch path.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace {

constexpr int DK = 96;               // head dim
constexpr int NCHK = DK / 8;
constexpr int BM = 128;              // queries per CTA
constexpr int BK = 256;              // keys per CTA
constexpr int NG = 4;                // softmax column groups (64 keys each)
constexpr int NRP = 32;              // relative positions, padded (2w+1 <= 32)
constexpr int NEPI = NG * 128;       // softmax threads
constexpr int NTHREADS = 128 + NEPI; // warp 0: loader, warp 1: MMA issuer, warps 2-3: idle, warps 4-19: softmax (lane quarter = warp % 4, group = (warp - 4) / 4)
constexpr int PADL = 64;
constexpr int MAXC = 4;              // CTAs per cluster
constexpr int QPLANE_BYTES = NCHK * BM * 16;         // one fp16 plane of the 128-query tile
constexpr int KPLANE_BYTES = NCHK * BK * 16;         // ... of the 256-key tile
constexpr int EKP_BYTES = NCHK * NRP * 16;
constexpr int QE_LD = 24;
constexpr int SMEM_BASE = 2 * QPLANE_BYTES + 2 * KPLANE_BYTES + 2 * EKP_BYTES + BM * QE_LD * 4 + BK * 4 + 2 * NG * BM * 4 + 2 * MAXC * BM * 4 + 64;
constexpr int RS_BYTES = NCHK * BM * 32;             // fused P.v: fp32 partial sums received from the CTAs of the cluster, [src][chunk][row][8] = 12 x 4 KB
constexpr int SMEM_BYTES = SMEM_BASE + RS_BYTES;
constexpr int COL_S = 0, COL_R = 256, COL_O = 288, TMEM_COLS = 512;
constexpr int PCH = BK / 8 + NRP / 8;                // chunks of the P tile / of the v^T tile: 32 key chunks + 4 relative-position chunks
// fused P.v: the P tile [PCH][128 rows][16 B] and the v^T tile [PCH][96 dims][16 B] take the place of the q / k tiles once S is complete
static_assert(PCH * BM * 16 + PCH * DK * 16 <= 2 * QPLANE_BYTES + 2 * KPLANE_BYTES, "P and v^T tiles must fit the q / k region");

struct AParams {
    const uint8_t* qk;        // packed q / k planes of the qkv projection: [b][planes * cpp chunks][Lp][16 B]
    int64_t qk_bs;            // bytes per batch item
    int32_t q_chunk0, k_chunk0, head_chunks, lo_chunks;    // first chunk of q / k of head 0, chunks per head, chunk offset of the lo plane (0: single plane)
    const uint8_t* ek;        // packed Ek planes of this layer: [planes][NCHK][32 rows = relative positions][16 B]
    const float* qkv;         // fp32 [B, 3H, T] (v rows for the v^T weights)
    const float* mask;        // [B, T]
    uint8_t* P;               // [b][h][nchunk_tot][Lp][16 B]
    uint8_t* vw;              // [b][h][nchunk_tot][DK][16 B]
    int32_t B, H, nh, T, Lp, W, nkt, nqt, nchunk_tot, planes;
    float inv;
    // fused P.v (att != NULL): O = P.v + sum_r p[i, i+r-w] Ev[r] computed here, written as the packed fp16 operand of conv_o
    uint8_t* att; int64_t att_bs; int32_t att_lo_chunks; const uint8_t* evT;
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
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t h2u_sat(float a, float b) {      // as the conv epilogue packs: saturating
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t h2u(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// PV: the P.v GEMM (and the relative-value term) runs here too — P never leaves the chip — and the cluster writes att, the packed operand of conv_o
template <bool PV>
__global__ void __launch_bounds__(NTHREADS, 1) attn_scores_cluster_kernel(const AParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* q_s = smem;                                  // [planes][NCHK][128][16]
    uint8_t* k_s = q_s + 2 * QPLANE_BYTES;                // [planes][NCHK][256][16]
    uint8_t* e_s = k_s + 2 * KPLANE_BYTES;                // [planes][NCHK][32][16]
    float* qe_s = reinterpret_cast<float*>(e_s + 2 * EKP_BYTES);     // [128][QE_LD]  (q_i . Ek[r]) / sqrt(dk)
    float* km_s = qe_s + BM * QE_LD;                      // [256] key mask of this CTA's keys (-1 beyond T)
    float* lm_s = km_s + BK;                              // [NG][128] row maxima of the column groups
    float* ls_s = lm_s + NG * BM;                         // [NG][128] row sums of the column groups
    float* mx_s = ls_s + NG * BM;                         // [MAXC][128] row maxima of every CTA of the cluster
    float* sm_s = mx_s + MAXC * BM;                       // [MAXC][128] row sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_s + MAXC * BM);
    const uint32_t LOAD_FULL = tc::smem_u32(bars), ACC_FULL = LOAD_FULL + 8, P_READY = LOAD_FULL + 16, O_FULL = LOAD_FULL + 24;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    uint8_t* p_s = smem;                                  // PV: [PCH][128][16] probabilities (K-major A operand), over the dead q / k tiles
    uint8_t* v_s = smem + PCH * BM * 16;                  // PV: [PCH][96][16] v^T of this CTA's keys + Ev^T (K-major B operand)
    float* rs_s = reinterpret_cast<float*>(smem + SMEM_BASE);     // PV: [src CTA][chunk][128][8] partial sums of this CTA's output chunks

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int kt = (int)cluster_ctarank();
    const int cl = (int)blockIdx.x / p.nkt;                // cluster -> (batch item, head, query tile)
    const int qt = cl % p.nqt, bh = cl / p.nqt;
    const int h = bh % p.nh, b = bh / p.nh;
    const int nterm = p.planes == 2 ? 3 : 1;

    if (threadIdx.x == 0) {
        tc::mbar_init(LOAD_FULL, 1);
        tc::mbar_init(ACC_FULL, 1);
        tc::mbar_init(P_READY, NEPI / 32);
        tc::mbar_init(O_FULL, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ======================= loader: q tile, k tile (both planes), Ek planes =======================
        const uint32_t nbytes = (uint32_t)p.planes * (QPLANE_BYTES + KPLANE_BYTES + EKP_BYTES);
        if (lane == 0) tc::mbar_arrive_expect_tx(LOAD_FULL, nbytes);
        __syncwarp();
        const uint8_t* base = p.qk + (size_t)b * p.qk_bs;
        for (int u = lane; u < p.planes * NCHK * 3; u += 32) {
            const int what = u / (p.planes * NCHK), rr = u - what * p.planes * NCHK;     // 0: q, 1: k, 2: Ek
            const int pl = rr / NCHK, c = rr - pl * NCHK;
            if (what == 2) {
                tc::bulk_g2s(tc::smem_u32(e_s + (size_t)pl * EKP_BYTES + (size_t)c * NRP * 16), p.ek + ((size_t)pl * NCHK + c) * NRP * 16, NRP * 16, LOAD_FULL);
            } else {
                const int chunk = (what == 0 ? p.q_chunk0 : p.k_chunk0) + h * p.head_chunks + c + pl * p.lo_chunks;
                if (what == 0)
                    tc::bulk_g2s(tc::smem_u32(q_s + (size_t)pl * QPLANE_BYTES + (size_t)c * BM * 16), base + ((size_t)chunk * p.Lp + PADL + qt * BM) * 16, BM * 16, LOAD_FULL);
                else   // (rows past the sequence are the zero rows of the packed buffer; the last key tile may end before 256 rows: Lp covers roundup(T, 256) + 64)
                    tc::bulk_g2s(tc::smem_u32(k_s + (size_t)pl * KPLANE_BYTES + (size_t)c * BK * 16), base + ((size_t)chunk * p.Lp + PADL + kt * BK) * 16, BK * 16, LOAD_FULL);
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer: S = sum_terms Q . K^T (N = 128), rel = sum_terms Q . Ek^T (N = 32) =======================
        tc::mbar_wait(LOAD_FULL, 0);
        tc::fence_after_sync();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t idS = tc::make_idesc(0u, BK), idR = tc::make_idesc(0u, NRP);
        const uint32_t lboA = ((uint32_t)BM & 0x3FFFu) << 16, lboK = ((uint32_t)BK & 0x3FFFu) << 16, lboE = ((uint32_t)NRP & 0x3FFFu) << 16;
        const uint32_t qa = tc::smem_u32(q_s) >> 4, ka = tc::smem_u32(k_s) >> 4, ea = tc::smem_u32(e_s) >> 4;
        uint32_t accS = 0, accR = 0;
        for (int t = 0; t < nterm; ++t) {
            const uint32_t qp = t == 1 ? QPLANE_BYTES / 16 : 0, kp = t == 2 ? KPLANE_BYTES / 16 : 0, ep = t == 2 ? EKP_BYTES / 16 : 0;    // (hi,hi) (lo,hi) (hi,lo)
            for (int ks = 0; ks < NCHK / 2; ++ks) {
                const uint32_t da = lboA | ((qa + qp + (uint32_t)(2 * ks * BM)) & 0x3FFFu);
                const uint32_t db = lboK | ((ka + kp + (uint32_t)(2 * ks * BK)) & 0x3FFFu);
                const uint32_t de = lboE | ((ea + ep + (uint32_t)(2 * ks * NRP)) & 0x3FFFu);
                if (tc::elect_one()) {
                    tc::umma<false>(tmem_base + COL_S, ((uint64_t)desc_hi << 32) | da, ((uint64_t)desc_hi << 32) | db, idS, accS);
                    tc::umma<false>(tmem_base + COL_R, ((uint64_t)desc_hi << 32) | da, ((uint64_t)desc_hi << 32) | de, idR, accR);
                }
                accS = accR = 1;
            }
        }
        if (tc::elect_one()) tc::umma_commit(ACC_FULL);
        __syncwarp();
    } else if (warp >= 4) {
        // ======================= softmax warps: thread = query 128 qt + r, keys [64 g, 64 g + 64) of this CTA's 256 =======================
        const int q = warp & 3, g = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const int et = threadIdx.x - 128;
        const int i = qt * BM + r;
        const bool qok = i < p.T;
        const int key0 = kt * BK;
        const float* mb = p.mask + (size_t)b * p.T;
        // PV: this thread's share of v^T (six 8-key units of one dim each), requested now, converted and stored once the q / k tiles are dead
        constexpr int VU = (BK / 8) * DK / NEPI;           // 6
        float vst[PV ? VU : 1][8];
        if constexpr (PV) {
            const float* vb = p.qkv + ((size_t)b * 3 * p.H + 2 * p.H + h * DK) * p.T;
#pragma unroll
            for (int n = 0; n < VU; ++n) {
                const int u = et + n * NEPI;
                const int d = u / (BK / 8), cc = u - d * (BK / 8);      // lanes run over the key chunks of one dim: contiguous reads
                const int k0 = key0 + 8 * cc;
#pragma unroll
                for (int t = 0; t < 8; ++t) vst[n][t] = k0 + t < p.T ? vb[(size_t)d * p.T + k0 + t] : 0.f;
            }
        }
        if (!PV && qt == 0) {
            // v^T for this CTA's keys (the P.v GEMM's per-batch weights), written by the clusters of query tile 0: [key / 8][d][8 keys] fp16
            const float* vb = p.qkv + ((size_t)b * 3 * p.H + 2 * p.H + h * DK) * p.T;
            uint8_t* vw = p.vw + ((size_t)(b * p.nh + h) * p.nchunk_tot) * DK * 16;
            for (int u = et; u < (BK / 8) * DK; u += NEPI) {
                const int d = u / (BK / 8), cc = u - d * (BK / 8);      // lanes run over the key chunks of one dim: contiguous reads
                const int k0 = key0 + 8 * cc;
                float v[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) v[t] = k0 + t < p.T ? vb[(size_t)d * p.T + k0 + t] : 0.f;
                *reinterpret_cast<uint4*>(vw + ((size_t)(k0 / 8) * DK + d) * 16) = make_uint4(h2u(v[0], v[1]), h2u(v[2], v[3]), h2u(v[4], v[5]), h2u(v[6], v[7]));
            }
        }
        const float mq = qok ? mb[i] : 0.f;
        if (et < BK) km_s[et] = key0 + et < p.T ? mb[key0 + et] : -1.f;      // -1: beyond the sequence (excluded), 0: masked (-1e4), else valid
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const int NR = 2 * p.W + 1;
        tc::mbar_wait(ACC_FULL, 0);
        tc::fence_after_sync();
        if constexpr (PV) {
            // S is complete: the q / k tiles are dead.  v^T of this CTA's keys, Ev^T behind it, zeros in the band chunks of the P tile.
#pragma unroll
            for (int n = 0; n < VU; ++n) {
                const int u = et + n * NEPI;
                const int d = u / (BK / 8), cc = u - d * (BK / 8);
                *reinterpret_cast<uint4*>(v_s + ((size_t)cc * DK + d) * 16) =
                    make_uint4(h2u(vst[n][0], vst[n][1]), h2u(vst[n][2], vst[n][3]), h2u(vst[n][4], vst[n][5]), h2u(vst[n][6], vst[n][7]));
            }
            if (et < (NRP / 8) * DK) *reinterpret_cast<uint4*>(v_s + ((size_t)(BK / 8) * DK + et) * 16) = reinterpret_cast<const uint4*>(p.evT)[et];
            *reinterpret_cast<uint4*>(p_s + ((size_t)(BK / 8) * BM + et) * 16) = make_uint4(0, 0, 0, 0);      // 4 chunks x 128 rows = 512 units
        }
        if (g == 0) {   // (q_i . Ek[r]) / sqrt(dk) for the 2w+1 relative positions -> row r of qe_s
            uint32_t v[32];
            tc::tmem_ld<32>(t_row + COL_R, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < QE_LD; ++t) qe_s[r * QE_LD + t] = __uint_as_float(v[t]) * p.inv;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");           // km_s, qe_s
        // ---- pass A: logits (scale, relative-key term, mask) parked in TMEM; maximum over this thread's 64 keys
        float m = -INFINITY;
        const int rel0 = key0 + 64 * g - i + p.W;                 // rel of this thread's column 0
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
            uint32_t v[32];
            tc::tmem_ld<32>(t_row + COL_S + 64 * g + cb * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const int col = cb * 32 + t;
                const float km = km_s[64 * g + col];
                float sv = __uint_as_float(v[t]) * p.inv;
                const int rel = rel0 + col;
                if ((unsigned)rel < (unsigned)NR) sv += qe_s[r * QE_LD + rel];
                if (mq * km == 0.f) sv = -1e4f;
                if (km < 0.f || !qok) sv = -INFINITY;
                m = fmaxf(m, sv);
                v[t] = __float_as_uint(sv);
            }
            tmem_st32(t_row + COL_S + 64 * g + cb * 32, v);
        }
        tmem_st_wait();
        lm_s[g * BM + r] = m;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (g == 0) {
            const float mm = fmaxf(fmaxf(lm_s[r], lm_s[BM + r]), fmaxf(lm_s[2 * BM + r], lm_s[3 * BM + r]));
            const uint32_t a = tc::smem_u32(mx_s + kt * BM + r);
            for (int k = 0; k < p.nkt; ++k) st_cluster_f32(mapa(a, (uint32_t)k), mm);
        }
    }
    cluster_sync_all();                                    // (1) row maxima of all key tiles are in place (every thread of the cluster takes part)
    if (warp >= 4) {
        const int q = warp & 3, g = (warp - 4) >> 2, r = q * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        float M = -INFINITY, l = 0.f;
        for (int k = 0; k < p.nkt; ++k) M = fmaxf(M, mx_s[k * BM + r]);
        // ---- pass B: exponentials parked in TMEM; sum over this thread's 64 keys
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
            uint32_t v[32];
            tc::tmem_ld<32>(t_row + COL_S + 64 * g + cb * 32, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                const float e = M == -INFINITY ? 0.f : __expf(__uint_as_float(v[t]) - M);      // exp(-inf) = 0 for keys beyond the sequence
                l += e;
                v[t] = __float_as_uint(e);
            }
            tmem_st32(t_row + COL_S + 64 * g + cb * 32, v);
        }
        tmem_st_wait();
        ls_s[g * BM + r] = l;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (g == 0) {
            const float ll = (ls_s[r] + ls_s[BM + r]) + (ls_s[2 * BM + r] + ls_s[3 * BM + r]);
            const uint32_t a = tc::smem_u32(sm_s + kt * BM + r);
            for (int k = 0; k < p.nkt; ++k) st_cluster_f32(mapa(a, (uint32_t)k), ll);
        }
    }
    cluster_sync_all();                                    // (2) row sums
    if constexpr (PV) {
        if (warp == 1) {
            // O[128, 96] = P[128, 256 + 32] . v^T[96, 256 + 32]^T once the softmax warps have written the P tile (pass C) and v^T is staged
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);
            tc::mbar_wait(P_READY, 0);
            tc::fence_after_sync();
            const uint32_t idO = tc::make_idesc(0u, DK);
            const uint32_t lboP = ((uint32_t)BM & 0x3FFFu) << 16, lboV = ((uint32_t)DK & 0x3FFFu) << 16;
            const uint32_t pa = tc::smem_u32(p_s) >> 4, va = tc::smem_u32(v_s) >> 4;
            for (int ks = 0; ks < PCH / 2; ++ks) {
                const uint32_t da = lboP | ((pa + (uint32_t)(2 * ks * BM)) & 0x3FFFu);
                const uint32_t db = lboV | ((va + (uint32_t)(2 * ks * DK)) & 0x3FFFu);
                if (tc::elect_one()) tc::umma<false>(tmem_base + COL_O, ((uint64_t)desc_hi << 32) | da, ((uint64_t)desc_hi << 32) | db, idO, ks ? 1u : 0u);
            }
            if (tc::elect_one()) tc::umma_commit(O_FULL);
            __syncwarp();
        }
    }
    if (warp >= 4) {
        const int q = warp & 3, g = (warp - 4) >> 2, r = q * 32 + lane;
        const int i = qt * BM + r;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        float L = 0.f;
        for (int k = 0; k < p.nkt; ++k) L += sm_s[k * BM + r];
        const float rl = (i < p.T && L > 0.f) ? 1.f / L : 0.f;
        const int NR = 2 * p.W + 1;
        const int rel0 = kt * BK + 64 * g - i + p.W;
        uint8_t* Pb = p.P + ((size_t)(b * p.nh + h) * p.nchunk_tot * p.Lp + PADL + i) * 16;
        // ---- pass C: probabilities -> packed fp16 (+ the band members as the extra "relative position" columns)
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
            uint32_t v[32];
            tc::tmem_ld<32>(t_row + COL_S + 64 * g + cb * 32, v);
            tc::tmem_ld_wait();
            if (PV || i < p.T) {      // (PV: the whole tile is an MMA operand — rows past the sequence are written as zeros, rl = 0)
                const bool band = rel0 + cb * 32 < NR && rel0 + cb * 32 + 31 >= 0;      // (any of these 32 keys within +-w of the query?)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float p8[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) p8[t] = __uint_as_float(v[8 * u + t]) * rl;
                    const uint4 pk = make_uint4(h2u(p8[0], p8[1]), h2u(p8[2], p8[3]), h2u(p8[4], p8[5]), h2u(p8[6], p8[7]));
                    if constexpr (PV) *reinterpret_cast<uint4*>(p_s + ((size_t)(g * 8 + cb * 4 + u) * BM + r) * 16) = pk;
                    else *reinterpret_cast<uint4*>(Pb + (size_t)(kt * (BK / 8) + g * 8 + cb * 4 + u) * p.Lp * 16) = pk;
                    if (band && i < p.T) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            const int rel = rel0 + cb * 32 + 8 * u + t;
                            if ((unsigned)rel < (unsigned)NR) {
                                if constexpr (PV) *reinterpret_cast<__half*>(p_s + ((size_t)(BK / 8 + rel / 8) * BM + r) * 16 + (rel % 8) * 2) = __float2half_rn(p8[t]);
                                else *reinterpret_cast<__half*>(Pb + (size_t)(p.nchunk_tot - NRP / 8 + rel / 8) * p.Lp * 16 + (rel % 8) * 2) = __float2half_rn(p8[t]);
                            }
                        }
                    }
                }
            }
        }
        if constexpr (PV) {
            // ---- O = P . v: hand the tiles to the tensor core, then reduce-scatter the partial sums of the cluster's key tiles
            tc::fence_before_sync();
            tc::fence_proxy_async();          // the P / v^T tiles are read through the async proxy
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(P_READY);
            tc::mbar_wait(O_FULL, 0);
            tc::fence_after_sync();
            // thread (r, g) holds output chunks 3g .. 3g+2 (8 dims each) of query r; chunk c belongs to CTA c / cpd of the cluster
            const int cpd = NCHK / p.nkt;
            uint32_t o16[16], o8[8];
            tc::tmem_ld<16>(t_row + COL_O + 24 * g, o16);
            tmem_ld8(t_row + COL_O + 24 * g + 16, o8);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c8 = 0; c8 < 3; ++c8) {
                const int ch = 3 * g + c8, dst = ch / cpd, pos = ch - dst * cpd;
                const uint32_t a = mapa(tc::smem_u32(rs_s + ((size_t)(kt * cpd + pos) * BM + r) * 8), (uint32_t)dst);
                const uint32_t* o = c8 < 2 ? &o16[8 * c8] : &o8[0];
                st_cluster_v4(a, __uint_as_float(o[0]), __uint_as_float(o[1]), __uint_as_float(o[2]), __uint_as_float(o[3]));
                st_cluster_v4(a + 16, __uint_as_float(o[4]), __uint_as_float(o[5]), __uint_as_float(o[6]), __uint_as_float(o[7]));
            }
        }
    }
    if constexpr (PV) {
        tc::fence_before_sync();
        cluster_sync_all();                                // (3) every CTA holds the partial sums of its output chunks
        if (warp >= 4) {
            const int q = warp & 3, g = (warp - 4) >> 2, r = q * 32 + lane;
            const int i = qt * BM + r;
            const int cpd = NCHK / p.nkt;
            if (i < p.T)
                for (int pos = g; pos < cpd; pos += NG) {
                    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int k = 0; k < p.nkt; ++k) {
                        const float4 a = *reinterpret_cast<const float4*>(rs_s + ((size_t)(k * cpd + pos) * BM + r) * 8);
                        const float4 c = *reinterpret_cast<const float4*>(rs_s + ((size_t)(k * cpd + pos) * BM + r) * 8 + 4);
                        o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w; o[4] += c.x; o[5] += c.y; o[6] += c.z; o[7] += c.w;
                    }
                    // the packed operand of conv_o: hi = fp16(O), lo = fp16(O - hi) (what the P.v launch's epilogue writes: y_act / y_lo)
                    uint8_t* dstp = p.att + (size_t)b * p.att_bs + ((size_t)(h * NCHK + kt * cpd + pos) * p.Lp + PADL + i) * 16;
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        hw[t] = h2u_sat(o[2 * t], o[2 * t + 1]);
                        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[t]));
                        lw[t] = h2u_sat(o[2 * t] - hf.x, o[2 * t + 1] - hf.y);
                    }
                    *reinterpret_cast<uint4*>(dstp) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    if (p.att_lo_chunks) *reinterpret_cast<uint4*>(dstp + (size_t)p.att_lo_chunks * p.Lp * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
        }
    }

    tc::fence_before_sync();
    cluster_sync_all();          // no CTA leaves while a peer may still write its row maxima / sums into it
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace

RVC_API int rvc_attn_scores_cluster_tc(const rvc_attn_scores_desc* dp, void* stream) {
    if (!dp) return rvc::fail(RVC_EINVAL, "rvc_attn_scores_cluster_tc: null descriptor");
    const rvc_attn_scores_desc& d = *dp;
    const bool pv = d.att != nullptr;
    if (!d.qk || !d.ek || !d.qkv || !d.mask || (!pv && (!d.P || !d.vw))) return rvc::fail(RVC_EINVAL, "rvc_attn_scores_cluster_tc: null pointer");
    if (pv && (!d.evT || (((uintptr_t)d.att | (uintptr_t)d.evT) & 15) || d.att_bs <= 0 || d.att_lo_chunks < 0))
        return rvc::fail(RVC_EINVAL, "rvc_attn_scores_cluster_tc: the fused P.v needs evT, a 16-byte aligned att and its batch stride");
    if (d.B <= 0 || d.T <= 0 || d.T > BK * MAXC || d.n_heads <= 0 || d.H != d.n_heads * DK)
        return rvc::fail(RVC_ESHAPE, "rvc_attn_scores_cluster_tc: built for head dim 96 and T <= 1024 (got H = %d, heads = %d, T = %d)", d.H, d.n_heads, d.T);
    if (d.window < 0 || 2 * d.window + 1 > QE_LD) return rvc::fail(RVC_ESHAPE, "rvc_attn_scores_cluster_tc: window %d > 11", d.window);
    if (d.planes != 1 && d.planes != 2) return rvc::fail(RVC_EINVAL, "rvc_attn_scores_cluster_tc: planes must be 1 or 2");
    if (((uintptr_t)d.qk | (uintptr_t)d.ek | (uintptr_t)(pv ? nullptr : d.P) | (uintptr_t)(pv ? nullptr : d.vw)) & 15) return rvc::fail(RVC_EINVAL, "rvc_attn_scores_cluster_tc: packed buffers need 16-byte alignment");
    AParams p;
    p.qk = (const uint8_t*)d.qk; p.qk_bs = d.qk_bs; p.q_chunk0 = d.q_chunk0; p.k_chunk0 = d.k_chunk0; p.head_chunks = NCHK; p.lo_chunks = d.planes == 2 ? d.lo_chunks : 0;
    p.ek = (const uint8_t*)d.ek; p.qkv = d.qkv; p.mask = d.mask; p.P = (uint8_t*)d.P; p.vw = (uint8_t*)d.vw;
    p.B = d.B; p.H = d.H; p.nh = d.n_heads; p.T = d.T; p.Lp = (int)rvc_packed_rows(d.T); p.W = d.window;
    p.nkt = (d.T + BK - 1) / BK; p.nqt = (d.T + BM - 1) / BM; p.nchunk_tot = p.nkt * (BK / 8) + NRP / 8; p.planes = d.planes;
    p.inv = 1.f / sqrtf((float)DK);
    p.att = (uint8_t*)d.att; p.att_bs = d.att_bs; p.att_lo_chunks = d.att_lo_chunks; p.evT = (const uint8_t*)d.evT;
    auto kern = pv ? attn_scores_cluster_kernel<true> : attn_scores_cluster_kernel<false>;
    static bool opted[2][64] = {{false}};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[pv][dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_attn_scores_cluster_tc: smem opt-in: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) opted[pv][dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(d.B * d.n_heads * p.nqt * p.nkt));
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.nkt; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
    if (le != cudaSuccess) return rvc::fail((int)le, "rvc_attn_scores_cluster_tc: launch: %s", cudaGetErrorString(le));
    return rvc::check_launch("rvc_attn_scores_cluster_tc");
}
