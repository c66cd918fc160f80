// Conv1d (Cin -> Cout, odd kernel K, dilation d, "same" padding, stride 1) as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[128 positions, NT out-channels] += A_tap[128 positions, 16 in-channels] * W_tap[NT, 16]^T     (tcgen05.mma kind::f16,
//                                                                                                    fp32 accumulators in TMEM)
// Data layout ("C8 packed", DT = bf16 or fp16):  act[b][c / 8][64 + l][8].  One 16-byte unit holds 8 consecutive channels of
// ONE position and consecutive positions are 16 B apart — exactly the K-major no-swizzle UMMA core-matrix layout (8 rows x
// 16 B contiguous).  A conv tap that reads position l + j*dil - pad is therefore the SAME shared-memory tile with the
// descriptor start address advanced by 16*j*dil bytes: the activation tile (TM + halo positions x a stage of input
// channels) is staged once with plain 1-D bulk async copies (one per 16-byte channel chunk; the zero rows around the
// sequence live in the global buffer, so there is no boundary logic) and reused by every tap.  Weights, pre-packed
// [K][Cin/8][Cout][8], stream through a 4-slot ring.
//
// Warp roles (512 threads): warp 0 = activation producer, warp 2 = weight producer, warps 1 and 3 = MMA issuers (sub-tiles
// split between them; warp 1 also owns the TMEM allocation), warps 4-15 = epilogue, three warpgroups (TMEM -> registers ->
// bias / gate / relu / residual / mask -> fp32 NCL and packed-DT stores, all coalesced along positions).  setmaxnreg moves
// registers from warpgroup 0 (56) to the epilogue warpgroups (152): the epilogue is instruction-issue bound (ncu).  Accumulators are double-buffered in TMEM so the epilogue of tile t overlaps the MMAs of tile
// t+1; CTAs are persistent over a static round-robin tile list spanning up to three independent jobs (the three ResBlocks
// of a generator stage, or the res/skip halves of a WaveNet layer).
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128;          // positions per accumulator (TMEM lanes)
constexpr int PADL = 64;         // zero rows in front of position 0 in packed buffers
constexpr int PADR = 64;
constexpr int ROWQ = 256;        // packed row count is rounded up to this
constexpr int NSTAGE = 4;        // weight ring slots
constexpr int STAGE_BYTES = 16384;
constexpr int NTHREADS = 512;      // 16 warps: warpgroup 0 = A producer, MMA issuers, W producer; warpgroups 1-3 = epilogue
constexpr int NEG = 3;             // epilogue warpgroups
constexpr int MAXJOBS = 6;
constexpr int MAXABUF = 8;        // activation ring slots (as many as fit next to the weight ring)
constexpr int MAXNG = 4;          // TMEM accumulator groups

__host__ __device__ inline int packed_rows(int L) { return PADL + ((L + ROWQ - 1) / ROWQ) * ROWQ + PADR; }

// Granlund-Montgomery division by an invariant (n < 2^31): q = (umulhi(m, n) + n) >> l.  The role loops decode a tile id per
// iteration; hardware integer division costs the single issuing warp ~130 cycles each (tools/mma_layout_bench.cu).
struct FastDiv {
    uint32_t m, l, d;
    __host__ void init(uint32_t div) {
        d = div;
        l = 0;
        while ((1ull << l) < div) ++l;
        m = (uint32_t)(((1ull << 32) * ((1ull << l) - div)) / div + 1);
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return (uint32_t)(((uint64_t)__umulhi(m, n) + n) >> l); }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

struct Params {
    rvc_tc_job jobs[MAXJOBS];
    int njobs, B, Cin, Cout, L, Lp;
    int nch;          // Cin / 8: 16-byte input-channel chunks
    int nch_out;      // packed output chunks (Cout/8, or Cout/16 when gated)
    int ach;          // chunks per activation stage
    int ncs;          // activation stages per tile
    int rows_max;     // rows allocated per chunk in an activation stage (TM + max halo)
    int n_lt, n_nt;   // tiles along L and along Cout
    int tiles_per_job, total_tiles;
    FastDiv fd_tpj, fd_up, fd_nnt, fd_nlt;
    int fp16;         // operand format: 0 = bf16, 1 = fp16
    int nabuf;        // activation ring depth (2..MAXABUF)
    int sum_jobs;     // 1: every tile runs all G jobs of its group into one accumulator (stage average of the ResBlocks; split-precision terms)
    int G;            // group size: 1 = every job has its own tiles; otherwise jobs [g*G, (g+1)*G) share the tiles of group g
    int has_res;      // some job carries a residual (enables the epilogue's residual prefetch)
    int res_split;    // ... as the split-fp16 stream (res_hi / res_lo); a launch takes its residuals in one format
    int blocked;      // bit 0: res is channel-blocked fp32 [b][c/8][l][8]; bit 1: y_f32 is
    int dbg;          // timing experiments only: 1 = epilogue skips global I/O, 2 = skip MMA issue, 4 = skip operand loads,
                      // 8 = skip residual loads, 16 = skip fp32 stores, 32 = skip packed stores, 64 = no residual prefetch
    int up, pshift;   // transposed-conv phase mode: output position = l * up + phase - pshift (up = 1: plain conv)
    int Lout, Lp_out; // output length and packed rows of the output
    int ucols, cr_shift;   // phase-column mode (rvc_conv1d_tc_desc.up_cols): the Cout weight columns are ucols phase blocks of Cr = 1 << cr_shift channels
};

__device__ __forceinline__ uint32_t pack2(float a, float b, int fp16) {
    uint32_t r;
    if (fp16) {   // saturating: values beyond +-65504 clamp instead of becoming inf
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    } else {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    }
    return r;
}
// tanh(t) * sigmoid(s) = (e^{2t} - 1) / ((e^{2t} + 1) (1 + e^{-s})) with the fast exponential and one fast division: ~14 instructions
// instead of ~60 for tanhf * 1/(1+expf) — the gated epilogue of a WaveNet in_layer runs on 4 warps per tile and was a third of
// that launch.  Absolute error <= ~2e-7 (the result is rounded to fp16 right after); arguments clamped where the functions
// are saturated to 1e-13 anyway so that no intermediate overflows.
__device__ __forceinline__ float gate_fast(float t, float s) {
    const float e2 = __expf(2.f * fminf(fmaxf(t, -15.f), 15.f));
    const float es = __expf(-fmaxf(s, -30.f));
    return __fdividef(e2 - 1.f, (e2 + 1.f) * (1.f + es));
}
// leaky_relu for 0 <= slope <= 1 as max(v, slope*v): two instructions, no select
__device__ __forceinline__ float lrelu_max(float v, float slope) { return fmaxf(v, v * slope); }
// its inverse for 0 < slope <= 1 (inv = 1 / slope): y > 0 -> y, y < 0 -> y / slope
__device__ __forceinline__ float lrelu_inv(float y, float inv) { return fminf(y, y * inv); }
__device__ __forceinline__ float2 h2f2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
// one 16-byte unit of each plane of the split-fp16 residual stream -> the eight fp32 residual values
__device__ __forceinline__ void split_decode8(const uint4& h, const uint4& lo, float inv, float* x) {
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = h2f2(hw[i]), b = h2f2(lw[i]);
        x[2 * i] = lrelu_inv(a.x + b.x, inv);
        x[2 * i + 1] = lrelu_inv(a.y + b.y, inv);
    }
}

// PF selects the residual code of the epilogue: 0 = fp32 residuals (if any) loaded when an item is processed — the launches without residuals;
// 1 = fp32 residuals prefetched one work item ahead; 2 = the split-fp16 stream (res_hi / res_lo), prefetched (raw words, one job per tile);
// 3 = the split-fp16 stream loaded when the item is processed (summed groups: their decode-and-add needs the registers a prefetch would hold).
// Each variant contains only its own residual code: with all of it in one kernel the prefetching instantiations spilled and the others —
// the conv1 and ConvTranspose launches — ran 5-10 % slower.
template <int NT, int MT, int PF>
__global__ void __launch_bounds__(NTHREADS, 1) conv1d_tc_kernel(const Params p) {
    constexpr int TM = BM * MT;
    constexpr int NI = MT >= 2 ? 2 : 1;                     // MMA issuer warps
    constexpr int MPI = MT / NI;                            // sub-tiles per issuer
    constexpr int SCH = (STAGE_BYTES / (NT * 16)) & ~1;    // 16-byte chunks per weight stage (even)
    constexpr int NG = (512 / (MT * NT)) >= MAXNG ? MAXNG : (512 / (MT * NT));   // accumulator groups (>= 2)
    static_assert(NG >= 2, "need at least two accumulator groups");
    constexpr int ACC_COLS = NG * MT * NT;
    constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t a_bytes = (uint32_t)p.ach * p.rows_max * 16;
    uint8_t* a_smem = smem;                                         // [nabuf][ach][rows_max][16]
    uint8_t* w_smem = smem + (size_t)p.nabuf * a_bytes;             // [NSTAGE][STAGE_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + NSTAGE * STAGE_BYTES);
    const uint32_t bar0 = tc::smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    auto A_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
    auto A_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + MAXABUF + s); };
    auto ACC_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * MAXABUF + s); };
    auto ACC_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * MAXABUF + MAXNG + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2 * MAXABUF + 2 * MAXNG);
    float* bias_s = reinterpret_cast<float*>(bars + 2 * NSTAGE + 2 * MAXABUF + 2 * MAXNG + 2);   // [njobs][Cout], 16-byte aligned
    // fused noise conv of job 0: taps TAP-major, [4][Cout] (zero padded), then the bias [Cout] — the epilogue reads four channels per LDS.128
    // (8 + 8 kn loads per 32-channel item instead of 64)
    float* nzw_s = bias_s + (((p.njobs / p.G) * p.Cout + 3) & ~3);
    float* nzb_s = nzw_s + 4 * p.Cout;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // NI MMA issuer warps (sub-tiles split between them): every "consumed" barrier collects NI commits
        for (int s = 0; s < NSTAGE; ++s) { tc::mbar_init(W_FULL(s), 1); tc::mbar_init(W_EMPTY(s), NI); }
        for (int s = 0; s < MAXABUF; ++s) { tc::mbar_init(A_FULL(s), 1); tc::mbar_init(A_EMPTY(s), NI); }
        for (int s = 0; s < MAXNG; ++s) { tc::mbar_init(ACC_FULL(s), NI); tc::mbar_init(ACC_EMPTY(s), 4 * NEG); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
    if (warp >= 4)   // biases of all jobs -> shared memory (read back with broadcast LDS.128 in the epilogue)
        for (int i = threadIdx.x - 128; i < (p.njobs / p.G) * p.Cout; i += NTHREADS - 128) {   // one (summed) bias row per job group
            const int g = i / p.Cout, c = i - g * p.Cout;
            float v = 0.f;
            for (int jj = 0; jj < p.G; ++jj) v += p.jobs[g * p.G + jj].bias ? p.jobs[g * p.G + jj].bias[c] : 0.f;
            bias_s[i] = v;
        }
    if (warp >= 4 && p.jobs[0].nz_har) {
        const rvc_tc_job& j0 = p.jobs[0];
        for (int c = threadIdx.x - 128; c < (p.ucols > 1 ? (1 << p.cr_shift) : p.Cout); c += NTHREADS - 128) {
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < j0.nz_kn; ++j) t[j] = j0.nz_w[(size_t)c * j0.nz_kn + j];
            for (int j = 0; j < 4; ++j) nzw_s[j * p.Cout + c] = t[j];
            nzb_s[c] = j0.nz_b ? j0.nz_b[c] : 0.f;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: everything above (barrier init, TMEM allocation, bias staging — constant data only)
    // may overlap the tail of the previous kernel in the stream; activations are only touched after the wait.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // tile id -> (job, batch, L tile, N tile, phase); phase is the fastest index so that the u phases of one input tile run on
    // neighbouring CTAs at the same time and share the activation tile through L2
    auto decode = [&](int tile, int& j, int& b, int& lt, int& nt, int& ph) {
        uint32_t q, r, t;
        p.fd_tpj.divmod((uint32_t)tile, q, r);
        j = (int)q;
        p.fd_up.divmod(r, q, t);
        ph = (int)t;
        p.fd_nnt.divmod(q, r, t);
        nt = (int)t;
        p.fd_nlt.divmod(r, q, t);
        lt = (int)t;
        b = (int)q;
    };

    // Tile walk.  Tile ids are ordered heavy jobs first (the caller lists the k = 11 ResBlock before k = 7 and k = 3); CTA c takes
    // tile it*P + c in even waves and it*P + (P-1-c) in odd ones ("snake"): dealing a cost-sorted list back and forth balances the
    // per-CTA load almost like longest-processing-time-first (C = 256 stage, 288 tiles on 148 CTAs: makespan 19.0 -> 15.0 tap units
    // against 14.6 ideal; C = 128: 79 -> 75 against 71) where plain round-robin hands the same CTAs the heavy tile of every wave.
    const int P = (int)gridDim.x, cta = (int)blockIdx.x;
    const int n_waves = p.total_tiles / P, rem = p.total_tiles - n_waves * P;
    const int n_my = n_waves + ((((n_waves & 1) ? P - 1 - cta : cta) < rem) ? 1 : 0);
    auto tile_of = [&](int it) { return it * P + ((it & 1) ? P - 1 - cta : cta); };

    // register rebalancing (setmaxnreg): the producer / issuer warpgroup needs few registers, the epilogue warpgroups many
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ======================= activation producer =======================
        uint32_t ast = 0, aph = 0;
        for (int it = 0; it < n_my; ++it) {
            const int tile = tile_of(it);
            int j, b, lt, nt, ph;
            decode(tile, j, b, lt, nt, ph);
            for (int jj = 0; jj < p.G; ++jj) {
            const rvc_tc_job& jb = p.jobs[j * p.G + jj];
            const int pad = jb.pad;
            const int row0 = PADL + lt * TM - pad;                           // >= 0 (host checks pad <= PADL)
            // the last tile must not read past the chunk's Lp rows (ragged L, TM = 512, the extra tile of the transposed mode):
            // rows of the stage that are not loaded only reach outputs the epilogue discards (lin >= L)
            const int rows = min(TM + (jb.K - 1) * jb.dil, p.Lp - row0);
            for (int cs = 0; cs < p.ncs; ++cs) {
                const int c0 = cs * p.ach;
                const int nc = min(p.ach, p.nch - c0);
                if (lane == 0) {
                    tc::mbar_wait_relaxed(A_EMPTY(ast), aph ^ 1);
                    if (p.dbg & 4) tc::mbar_arrive(A_FULL(ast));
                    else tc::mbar_arrive_expect_tx(A_FULL(ast), (uint32_t)nc * rows * 16);
                }
                __syncwarp();
                if (p.dbg & 4) { if (++ast == (uint32_t)p.nabuf) { ast = 0; aph ^= 1; } continue; }
                const size_t a_bs = jb.a_bs ? (size_t)jb.a_bs : (size_t)p.nch * p.Lp * 16;
                const uint8_t* asrc = reinterpret_cast<const uint8_t*>(jb.a) + (size_t)b * a_bs +
                                      ((size_t)c0 * p.Lp + (size_t)row0) * 16;
                for (int c = lane; c < nc; c += 32)
                    tc::bulk_g2s(tc::smem_u32(a_smem + (size_t)ast * a_bytes + (size_t)c * p.rows_max * 16), asrc + (size_t)c * p.Lp * 16,
                                 (uint32_t)rows * 16, A_FULL(ast));
                if (++ast == (uint32_t)p.nabuf) { ast = 0; aph ^= 1; }
            }
            }
        }
    } else if (warp == 2) {
        // ======================= weight producer =======================
        uint32_t wst = 0, wph = 0;
        for (int it = 0; it < n_my; ++it) {
            const int tile = tile_of(it);
            int j, b, lt, nt, ph;
            decode(tile, j, b, lt, nt, ph);
            for (int jj = 0; jj < p.G; ++jj) {
            const rvc_tc_job& jb = p.jobs[j * p.G + jj];
            const size_t w_cs = jb.w_cs ? (size_t)jb.w_cs : (size_t)p.Cout;      // rows between consecutive weight chunks
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(jb.w) + (size_t)b * jb.w_bs + (size_t)ph * jb.K * p.nch * w_cs * 16;
            for (int cs = 0; cs < p.ncs; ++cs) {
                const int ca = cs * p.ach;
                const int na = min(p.ach, p.nch - ca);
                // flat unit index f = tap * na + chunk; a ring slot holds SCH consecutive units (several taps when C is small)
                const int nflat = jb.K * na;
                const bool contiguous = (NT == p.Cout) && (w_cs == (size_t)p.Cout) && (na == p.nch);
                for (int f0 = 0; f0 < nflat; f0 += SCH) {
                    const int nu = min(SCH, nflat - f0);
                    if (lane == 0) {
                        tc::mbar_wait_relaxed(W_EMPTY(wst), wph ^ 1);
                        if (p.dbg & 4) tc::mbar_arrive(W_FULL(wst));
                        else tc::mbar_arrive_expect_tx(W_FULL(wst), (uint32_t)nu * NT * 16);
                    }
                    __syncwarp();
                    if (p.dbg & 4) { if (++wst == NSTAGE) { wst = 0; wph ^= 1; } continue; }
                    const uint32_t dst = tc::smem_u32(w_smem + wst * STAGE_BYTES);
                    if (contiguous) {
                        if (lane == 0) tc::bulk_g2s(dst, wsrc + (size_t)f0 * NT * 16, (uint32_t)nu * NT * 16, W_FULL(wst));
                    } else {
                        for (int u = lane; u < nu; u += 32) {
                            const int f = f0 + u, tap = f / na, ch = f - tap * na;
                            tc::bulk_g2s(dst + u * NT * 16, wsrc + (((size_t)tap * p.nch + ca + ch) * w_cs + (size_t)nt * NT) * 16, NT * 16, W_FULL(wst));
                        }
                    }
                    if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
                }
            }
            }
        }
    } else if (warp == 1 || (NI == 2 && warp == 3)) {
        // ======================= MMA issuers: warp 1 -> sub-tiles 0 (and 2), warp 3 -> sub-tiles 1 (and 3).  The whole warp runs the (uniform)
        // loop so that descriptors live in uniform registers; one elected lane issues tcgen05.mma / tcgen05.commit.
        // Descriptors are updated incrementally: +2*rows_max (16-B units) per k-step, +dil per tap for A; +2*NT for W. =====
        const int m = (warp == 1) ? 0 : 1;
        const uint32_t idesc = tc::make_idesc(p.fp16 ? 0u : 1u, NT);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                  // SBO = 128 B, descriptor version 1, no swizzle
        const uint32_t a_lo_hi = ((uint32_t)p.rows_max & 0x3FFFu) << 16;   // LBO(A) = rows_max * 16 B
        const uint32_t b_lo_hi = ((uint32_t)NT & 0x3FFFu) << 16;           // LBO(W) = NT * 16 B
        const uint32_t rows2 = 2u * (uint32_t)p.rows_max;
        uint32_t wst = 0, wph = 0, ast = 0, aph = 0;
        for (int it = 0; it < n_my; ++it) {
            const int tile = tile_of(it);
            int j, b, lt, nt, ph;
            decode(tile, j, b, lt, nt, ph);
            const int g = it % NG;
            const uint32_t gph = (it / NG) & 1;
            tc::mbar_wait(ACC_EMPTY(g), gph ^ 1);
            tc::fence_after_sync();
            const uint32_t d_tmem = tmem_base + (uint32_t)((g * MT + m) * NT);   // + 2*NT for this issuer's second sub-tile
            uint32_t accumulate = 0;
            for (int jj = 0; jj < p.G; ++jj) {
            const int K = p.jobs[j * p.G + jj].K;
            const uint32_t dil = (uint32_t)p.jobs[j * p.G + jj].dil;
            for (int cs = 0; cs < p.ncs; ++cs) {
                const int na = min(p.ach, p.nch - cs * p.ach);
                tc::mbar_wait(A_FULL(ast), aph);
                tc::fence_after_sync();
                // descriptor low words: [13:0] start address / 16, [29:16] LBO / 16.  Per k-step: A += 2 chunks (2*rows_max rows),
                // W += 2 chunks (2*NT rows); per tap: A restarts at the tap's row shift.  No field overflows: smem < 256 KB.
                const uint32_t da_base = a_lo_hi | (((tc::smem_u32(a_smem + (size_t)ast * a_bytes) >> 4) + (uint32_t)(m * BM)) & 0x3FFFu);
                const int nflat = K * na;
                int ch = 0;
                uint32_t tap_off = 0, da_lo = da_base;
                for (int f0 = 0; f0 < nflat; f0 += SCH) {
                    const int nu = min(SCH, nflat - f0);
                    tc::mbar_wait(W_FULL(wst), wph);
                    tc::fence_after_sync();
                    uint32_t db_lo = b_lo_hi | ((tc::smem_u32(w_smem + wst * STAGE_BYTES) >> 4) & 0x3FFFu);
                    int u = 0;
                    while (u < nu) {
                        const int run = min(nu - u, na - ch);   // k-steps until the stage ends or the tap wraps (even)
                        for (int r = 0; r < run; r += 2) {
                            if (!(p.dbg & 2) && tc::elect_one()) {
                                tc::umma<false>(d_tmem, ((uint64_t)desc_hi << 32) | da_lo, ((uint64_t)desc_hi << 32) | db_lo, idesc, accumulate);
                                if constexpr (MPI == 2)   // sub-tile m + 2: same weights, activation rows 2*128 further down
                                    tc::umma<false>(d_tmem + 2u * NT, ((uint64_t)desc_hi << 32) | (da_lo + 2u * BM), ((uint64_t)desc_hi << 32) | db_lo,
                                                    idesc, accumulate);
                            }
                            accumulate = 1;
                            da_lo += rows2;
                            db_lo += 2u * NT;
                        }
                        u += run;
                        ch += run;
                        if (ch == na) { ch = 0; tap_off += dil; da_lo = da_base + tap_off; }
                    }
                    if (tc::elect_one()) tc::umma_commit(W_EMPTY(wst));
                    if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
                }
                if (tc::elect_one()) tc::umma_commit(A_EMPTY(ast));
                if (++ast == (uint32_t)p.nabuf) { ast = 0; aph ^= 1; }
            }
            }
            if (tc::elect_one()) tc::umma_commit(ACC_FULL(g));
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        // ======================= epilogue: NEG warpgroups x 4 warps; warp%4 selects the TMEM lane quarter, the warpgroup
        // takes every NEG-th (sub-tile, 32-column block) work item =======================
        const int q = warp & 3;
        const int eg = (warp - 4) >> 2;   // 0 .. NEG-1
        constexpr int EBP = NT < 32 ? 16 : 32;
        // Residual prefetch (PF instantiation), one work item ahead, also across tiles: the fp32 residual loads of item n+1 are
        // in flight while item n waits for its accumulator, does its math and stores (3-6 % on the conv2 launches).  Launches
        // without residuals use the PF = false instantiation: the 32 extra live registers cost the conv1 epilogue ~10 %.
        float vn[(PF == 1 || PF == 2) ? EBP : 1];
        auto load_res = [&](const rvc_tc_job& jn, int b_, int lt_, int nt_, int ph_, int item_, float* vn) {
            const int nblk_ = jn.gated ? (NT / EBP) / 2 : NT / EBP;
            const int m_ = item_ / nblk_, cb_ = item_ - m_ * nblk_;
            const int lin_ = lt_ * TM + m_ * BM + q * 32 + lane;
            const int l_ = lin_ * p.up + ph_ - p.pshift;
            if (!(lin_ < p.L && l_ >= 0 && l_ < p.Lout) || (p.dbg & 9)) return;
            const int co_ = jn.gated ? nt_ * (NT / 2) + cb_ * EBP : nt_ * NT + cb_ * EBP;
            if constexpr (PF >= 2) {
                if (!jn.res_hi) return;
                // split-fp16 stream (hi = an earlier launch's y_act, lo = its y_lo).  One job: the raw words now, decoded when the item is
                // processed, so the loads stay in flight; a summed group decodes as it adds.
                const size_t poff = (size_t)b_ * p.nch_out * p.Lp_out + (size_t)(co_ >> 3) * p.Lp_out + PADL + l_;      // 16-byte units
                if (PF == 2 || !p.sum_jobs) {
                    const uint4* ph4 = reinterpret_cast<const uint4*>(jn.res_hi) + poff;
                    const uint4* pl4 = reinterpret_cast<const uint4*>(jn.res_lo) + poff;
#pragma unroll
                    for (int u = 0; u < EBP / 8; ++u) {
                        const uint4 h = tc::ldg128(ph4 + (size_t)u * p.Lp_out), lo = tc::ldg128(pl4 + (size_t)u * p.Lp_out);
                        vn[8 * u + 0] = __uint_as_float(h.x); vn[8 * u + 1] = __uint_as_float(h.y);
                        vn[8 * u + 2] = __uint_as_float(h.z); vn[8 * u + 3] = __uint_as_float(h.w);
                        vn[8 * u + 4] = __uint_as_float(lo.x); vn[8 * u + 5] = __uint_as_float(lo.y);
                        vn[8 * u + 6] = __uint_as_float(lo.z); vn[8 * u + 7] = __uint_as_float(lo.w);
                    }
                } else if constexpr (PF == 3) {
                    const float inv = 1.f / jn.act_slope;
#pragma unroll
                    for (int u = 0; u < EBP / 8; ++u) {
                        uint4 h[3], lo[3];      // (groups of at most three jobs: the ResBlocks of a stage)
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj)
                            if (jj < p.G && (&jn)[jj].res_hi) {
                                h[jj] = tc::ldg128(reinterpret_cast<const uint4*>((&jn)[jj].res_hi) + poff + (size_t)u * p.Lp_out);
                                lo[jj] = tc::ldg128(reinterpret_cast<const uint4*>((&jn)[jj].res_lo) + poff + (size_t)u * p.Lp_out);
                            }
                        split_decode8(h[0], lo[0], inv, vn + 8 * u);
#pragma unroll
                        for (int jj = 1; jj < 3; ++jj)
                            if (jj < p.G && (&jn)[jj].res_hi) {
                                float t8[8];
                                split_decode8(h[jj], lo[jj], inv, t8);
#pragma unroll
                                for (int i = 0; i < 8; ++i) vn[8 * u + i] += t8[i];
                            }
                    }
                }
                return;
            }
            if (p.blocked & 1) {   // channel-blocked fp32: 32 contiguous bytes per 8 channels
                const float* rp = jn.res + (size_t)b_ * jn.res_bs + (size_t)(co_ >> 3) * jn.res_cs + (size_t)l_ * 8;
#pragma unroll
                for (int u = 0; u < EBP / 8; ++u) tc::ldg256(rp + (size_t)u * jn.res_cs, vn + 8 * u);
                if (p.sum_jobs)
                    for (int jj = 1; jj < p.G; ++jj) {
                        const rvc_tc_job& j2 = (&jn)[jj];
                        if (!j2.res) continue;
                        const float* rp2 = j2.res + (size_t)b_ * j2.res_bs + (size_t)(co_ >> 3) * j2.res_cs + (size_t)l_ * 8;
#pragma unroll
                        for (int u = 0; u < EBP / 8; ++u) {
                            float t8[8];
                            tc::ldg256(rp2 + (size_t)u * j2.res_cs, t8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) vn[8 * u + i] += t8[i];
                        }
                    }
                return;
            }
            const float* rp = jn.res + (size_t)b_ * jn.res_bs + (size_t)co_ * jn.res_cs + l_;
#pragma unroll
            for (int i = 0; i < EBP; ++i) vn[i] = rp[(size_t)i * jn.res_cs];
            if (p.sum_jobs)
                for (int jj = 1; jj < p.G; ++jj) {
                    const rvc_tc_job& j2 = (&jn)[jj];
                    if (!j2.res) continue;
                    const float* rp2 = j2.res + (size_t)b_ * j2.res_bs + (size_t)co_ * j2.res_cs + l_;
#pragma unroll
                    for (int i = 0; i < EBP; ++i) vn[i] += rp2[(size_t)i * j2.res_cs];
                }
        };
        // first item of a tile (for the very first tile and, from the last item of a tile, for the next one)
        auto prefetch_tile = [&](int tile_) {
            if (tile_ >= p.total_tiles) return;
            int j_, b_, lt_, nt_, ph_;
            decode(tile_, j_, b_, lt_, nt_, ph_);
            const rvc_tc_job& jn = p.jobs[j_ * p.G];
            if ((PF >= 2 ? jn.res_hi != nullptr : jn.res != nullptr) && eg < MT * (jn.gated ? (NT / EBP) / 2 : NT / EBP)) load_res(jn, b_, lt_, nt_, ph_, eg, vn);
        };
        constexpr bool PREFETCH = PF == 1 || PF == 2;
        if constexpr (PREFETCH) {
#pragma unroll
            for (int i = 0; i < EBP; ++i) vn[i] = 0.f;
            prefetch_tile(n_my > 0 ? tile_of(0) : p.total_tiles);
        }
        for (int it = 0; it < n_my; ++it) {
            const int tile = tile_of(it);
            int j, b, lt, nt, ph;
            decode(tile, j, b, lt, nt, ph);
            const int jl = j * p.G;                 // the group's leader: outputs, bias slot, epilogue options
            const rvc_tc_job& jb = p.jobs[jl];
            const int g = it % NG;
            const uint32_t gph = (it / NG) & 1;
            const float* bb = jb.bias_b ? jb.bias_b + (size_t)b * jb.bias_b_bs : nullptr;
            constexpr int EB = NT < 32 ? 16 : 32;     // accumulator columns handled per work item
            constexpr int NBLK = NT / EB;
            const int nblk = jb.gated ? NBLK / 2 : NBLK;
            tc::mbar_wait(ACC_FULL(g), gph);
            tc::fence_after_sync();
#pragma unroll 1
            for (int item = eg; item < MT * nblk; item += NEG) {
                const int m = item / nblk, cb = item - m * nblk;
                const int lin = lt * TM + m * BM + q * 32 + lane;
                const int col0 = nt * NT + cb * EB;               // weight-column index (bias order)
                int co0 = jb.gated ? nt * (NT / 2) + cb * EB : col0;         // output-channel index
                int l = lin * p.up + ph - p.pshift;                // output position
                if (p.ucols > 1) {   // phase columns: block r of Cr columns is phase r of a stride-ucols ConvTranspose
                    const int r_ = col0 >> p.cr_shift;
                    co0 = col0 - (r_ << p.cr_shift);
                    l = lin * p.ucols + r_ - p.pshift;
                }
                const bool ok = lin < p.L && l >= 0 && l < p.Lout && !(p.dbg & 1);
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g * MT + m) * NT);
                uint32_t r[EB];
                float v[EB];
                // this item's residual was prefetched one item ago; start the next item's loads before touching the accumulator
                if constexpr (PREFETCH) {   // this item's residual was requested one item ago; request the next one before the accumulator
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] = vn[i];
                    if (item + NEG < MT * nblk) { if (PF >= 2 ? jb.res_hi != nullptr : jb.res != nullptr) load_res(jb, b, lt, nt, ph, item + NEG, vn); }
                    else prefetch_tile(it + 1 < n_my ? tile_of(it + 1) : p.total_tiles);
                } else {
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] = 0.f;
                    if (PF >= 2 ? jb.res_hi != nullptr : jb.res != nullptr) load_res(jb, b, lt, nt, ph, item, v);
                    else if (jb.nz_har && ok) {
                        // fused noise conv (taps in the reference kernel's order: bias, then fma per tap)
                        const int64_t h0 = (int64_t)l * jb.nz_sn - jb.nz_pn, hlen = (int64_t)p.Lout * jb.nz_sn;
                        const float* hb = jb.nz_har + (size_t)b * jb.nz_har_bs;
                        float hv[4];
#pragma unroll
                        for (int jx = 0; jx < 4; ++jx) hv[jx] = (jx < jb.nz_kn && h0 + jx >= 0 && h0 + jx < hlen) ? hb[h0 + jx] : 0.f;
                        const int kn = jb.nz_kn;
#pragma unroll
                        for (int i4 = 0; i4 < EB / 4; ++i4) {      // (reference order per channel: bias, then one fma per tap)
                            float4 a = *reinterpret_cast<const float4*>(nzb_s + co0 + 4 * i4);
#pragma unroll
                            for (int jx = 0; jx < 4; ++jx)
                                if (jx < kn) {
                                    const float4 w4 = *reinterpret_cast<const float4*>(nzw_s + jx * p.Cout + co0 + 4 * i4);
                                    a.x = fmaf(w4.x, hv[jx], a.x); a.y = fmaf(w4.y, hv[jx], a.y); a.z = fmaf(w4.z, hv[jx], a.z); a.w = fmaf(w4.w, hv[jx], a.w);
                                }
                            v[4 * i4 + 0] = a.x; v[4 * i4 + 1] = a.y; v[4 * i4 + 2] = a.z; v[4 * i4 + 3] = a.w;
                        }
                    }
                }
                const float mk = (jb.mask && ok) ? jb.mask[(size_t)b * p.Lout + l] : 1.f;
                tc::tmem_ld<EB>(t_row + cb * EB, r);
                tc::tmem_ld_wait();
                const float4* bs4 = reinterpret_cast<const float4*>(bias_s + j * p.Cout + col0);
                if (jb.gated) {
                    uint32_t r2[EB];
                    tc::tmem_ld<EB>(t_row + (cb + NBLK / 2) * EB, r2);
                    tc::tmem_ld_wait();
                    const int col1 = col0 + (NBLK / 2) * EB;
                    const float* bs1 = bias_s + j * p.Cout + col1;
#pragma unroll
                    for (int i = 0; i < EB; ++i) {
                        float t = __uint_as_float(r[i]) + bias_s[j * p.Cout + col0 + i];
                        float sg = __uint_as_float(r2[i]) + bs1[i];
                        if (bb) { t += __ldg(bb + col0 + i); sg += __ldg(bb + col1 + i); }
                        r[i] = __float_as_uint(gate_fast(t, sg));
                    }
                } else {
#pragma unroll
                    for (int i4 = 0; i4 < EB / 4; ++i4) {
                        const float4 b4 = bs4[i4];
                        r[4 * i4 + 0] = __float_as_uint(__uint_as_float(r[4 * i4 + 0]) + b4.x);
                        r[4 * i4 + 1] = __float_as_uint(__uint_as_float(r[4 * i4 + 1]) + b4.y);
                        r[4 * i4 + 2] = __float_as_uint(__uint_as_float(r[4 * i4 + 2]) + b4.z);
                        r[4 * i4 + 3] = __float_as_uint(__uint_as_float(r[4 * i4 + 3]) + b4.w);
                    }
                    if (bb) {
#pragma unroll
                        for (int i = 0; i < EB; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __ldg(bb + col0 + i));
                    }
                }
                if (!ok) continue;
                const bool has_r = PF >= 2 ? jb.res_hi != nullptr : jb.res != nullptr;
                if (PF >= 2 && jb.res_hi && (PF == 2 || !p.sum_jobs)) {   // raw hi / lo words of the split-fp16 stream -> residual values
                    const float inv = 1.f / jb.act_slope;
#pragma unroll
                    for (int u = 0; u < EB / 8; ++u) {
                        const uint4 h = make_uint4(__float_as_uint(v[8 * u + 0]), __float_as_uint(v[8 * u + 1]), __float_as_uint(v[8 * u + 2]), __float_as_uint(v[8 * u + 3]));
                        const uint4 lo = make_uint4(__float_as_uint(v[8 * u + 4]), __float_as_uint(v[8 * u + 5]), __float_as_uint(v[8 * u + 6]), __float_as_uint(v[8 * u + 7]));
                        split_decode8(h, lo, inv, v + 8 * u);
                    }
                }
                if (jb.relu == 1) {
#pragma unroll
                    for (int i = 0; i < EB; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]), 0.f));
                }
                const float alpha = jb.alpha;
                if (p.sum_jobs) {   // v = alpha * (sum_j conv_j + sum_j bias_j + sum_j res_j)
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] = alpha * ((has_r ? v[i] : 0.f) + __uint_as_float(r[i]));
                } else if (has_r || jb.nz_har) {
                    if (alpha == 1.f) {
#pragma unroll
                        for (int i = 0; i < EB; ++i) v[i] = v[i] + __uint_as_float(r[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < EB; ++i) v[i] = fmaf(alpha, __uint_as_float(r[i]), v[i]);
                    }
                } else if (alpha == 1.f) {
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] = __uint_as_float(r[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] = alpha * __uint_as_float(r[i]);
                }
                float sl = jb.act_slope;
                if (jb.relu == 2) {   // leaky_relu on the final value itself (both outputs carry it)
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] = lrelu_max(v[i], sl);
                    sl = 1.f;
                }
                if (jb.mask) {
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] *= mk;
                }
                if (jb.y_f32 && !(p.dbg & 16)) {
                    if (p.blocked & 2) {
                        float* yp = jb.y_f32 + (size_t)b * jb.y_bs + (size_t)(co0 >> 3) * jb.y_cs + (size_t)l * 8;
#pragma unroll
                        for (int u = 0; u < EB / 8; ++u) tc::stg256(yp + (size_t)u * jb.y_cs, v + 8 * u);
                    } else {
                        float* yp = jb.y_f32 + (size_t)b * jb.y_bs + (size_t)co0 * jb.y_cs + l;
#pragma unroll
                        for (int i = 0; i < EB; ++i) yp[(size_t)i * jb.y_cs] = v[i];
                    }
                }
                if (jb.y_act && !(p.dbg & 32)) {
                    const size_t ya_bs = jb.y_act_bs ? (size_t)jb.y_act_bs : (size_t)p.nch_out * p.Lp_out * 16;
                    uint8_t* yb = reinterpret_cast<uint8_t*>(jb.y_act) + (size_t)b * ya_bs + (((size_t)co0 / 8) * p.Lp_out + PADL + l) * 16;
#pragma unroll
                    for (int u = 0; u < EB / 8; ++u) {
                        uint4 pk;
                        pk.x = pack2(lrelu_max(v[8 * u + 0], sl), lrelu_max(v[8 * u + 1], sl), p.fp16);
                        pk.y = pack2(lrelu_max(v[8 * u + 2], sl), lrelu_max(v[8 * u + 3], sl), p.fp16);
                        pk.z = pack2(lrelu_max(v[8 * u + 4], sl), lrelu_max(v[8 * u + 5], sl), p.fp16);
                        pk.w = pack2(lrelu_max(v[8 * u + 6], sl), lrelu_max(v[8 * u + 7], sl), p.fp16);
                        *reinterpret_cast<uint4*>(yb + (size_t)u * p.Lp_out * 16) = pk;
                        if (jb.y_lo) {   // low half of the split-fp16 residual stream: fp16(value - float(hi))  (fp16 launches only)
                            const uint32_t hw[4] = {pk.x, pk.y, pk.z, pk.w};
                            uint32_t lw[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
                                lw[i] = pack2(lrelu_max(v[8 * u + 2 * i], sl) - hf.x, lrelu_max(v[8 * u + 2 * i + 1], sl) - hf.y, 1);
                            }
                            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(jb.y_lo) + (yb - reinterpret_cast<uint8_t*>(jb.y_act)) + (size_t)u * p.Lp_out * 16) =
                                make_uint4(lw[0], lw[1], lw[2], lw[3]);
                        }
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(ACC_EMPTY(g));
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ packing kernels
// fp32 [B, C, L] (strided) -> packed [B][C/8][Lp][8]; value = lrelu(x, slope) * mask
__global__ void __launch_bounds__(256) pack_act_kernel(const float* __restrict__ x, const float* __restrict__ x2, const float* __restrict__ x3,
                                                       float in_div, int64_t x_bs, int64_t x_cs, const float* __restrict__ mask,
                                                       uint8_t* __restrict__ y, int nch, int L, int Lp, float slope, int fp16) {
    rvc::pdl_prologue();
    const int l = blockIdx.x * 256 + threadIdx.x;
    const int c = blockIdx.y, b = blockIdx.z;
    if (l >= L) return;
    const int64_t off = (int64_t)b * x_bs + (int64_t)c * 8 * x_cs + l;
    const float mk = mask ? mask[(int64_t)b * L + l] : 1.f;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float t = x[off + (int64_t)i * x_cs];
        if (x2) t = ((t + x2[off + (int64_t)i * x_cs]) + x3[off + (int64_t)i * x_cs]) / in_div;
        v[i] = rvc::lrelu(t, slope) * mk;
    }
    uint4 pk;
    pk.x = pack2(v[0], v[1], fp16); pk.y = pack2(v[2], v[3], fp16);
    pk.z = pack2(v[4], v[5], fp16); pk.w = pack2(v[6], v[7], fp16);
    *reinterpret_cast<uint4*>(y + (((size_t)b * nch + c) * Lp + PADL + l) * 16) = pk;
}

// folded fp32 [Cout][Cin][K] -> packed [K][Cin/8][Cout][8]
__global__ void pack_weight_kernel(const float* __restrict__ w, uint8_t* __restrict__ y, int Cout, int Cin, int K, int fp16) {
    const int nch = Cin / 8;
    const int64_t n = (int64_t)K * nch * Cout;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int co = (int)(idx % Cout);
    const int c = (int)((idx / Cout) % nch);
    const int tap = (int)(idx / ((int64_t)Cout * nch));
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = w[((int64_t)co * Cin + c * 8 + i) * K + tap];
    uint4 pk;
    pk.x = pack2(v[0], v[1], fp16); pk.y = pack2(v[2], v[3], fp16);
    pk.z = pack2(v[4], v[5], fp16); pk.w = pack2(v[6], v[7], fp16);
    *reinterpret_cast<uint4*>(y + idx * 16) = pk;
}

// per-device SM count, queried once (idempotent cache; no other global state)
int sm_count_cached() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = sms;
    }
    return cache[dev];
}

template <int NT, int MT, int PF>
int launch_tc_pf(const Params& p, size_t smem, int grid, cudaStream_t stream);

template <int NT, int MT>
int launch_tc(const Params& p, size_t smem, int grid, cudaStream_t stream) {
    if (p.res_split) {
        if constexpr (NT >= 64) {      // (the prefetching split-fp16 variant is built for the tiles the generator stages use)
            if (!p.sum_jobs && !(p.dbg & 64)) return launch_tc_pf<NT, MT, 2>(p, smem, grid, stream);
        }
        return launch_tc_pf<NT, MT, 3>(p, smem, grid, stream);
    }
    return (p.has_res && !(p.dbg & 64)) ? launch_tc_pf<NT, MT, 1>(p, smem, grid, stream) : launch_tc_pf<NT, MT, 0>(p, smem, grid, stream);
}

template <int NT, int MT, int PF>
int launch_tc_pf(const Params& p, size_t smem, int grid, cudaStream_t stream) {
    auto kern = conv1d_tc_kernel<NT, MT, PF>;
    // opt in to the maximum dynamic shared memory once per device and instantiation (the attribute is sticky)
    static bool opted[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_conv1d_tc: smem opt-in: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) opted[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
    if (le != cudaSuccess) return rvc::fail((int)le, "rvc_conv1d_tc: launch: %s", cudaGetErrorString(le));
    return rvc::check_launch("rvc_conv1d_tc");
}

template <int NT>
int launch_tc_mt(int mt, const Params& p, size_t smem, int grid, cudaStream_t stream) {
    if constexpr (NT <= 64) {
        if (mt == 4) return launch_tc<NT, 4>(p, smem, grid, stream);
    }
    return mt == 1 ? launch_tc<NT, 1>(p, smem, grid, stream) : launch_tc<NT, 2>(p, smem, grid, stream);
}

}  // namespace

RVC_API int64_t rvc_packed_rows(int32_t L) { return packed_rows(L); }

RVC_API int rvc_pack_act(const float* x, const float* x2, const float* x3, float in_div, int64_t x_bs, int64_t x_cs, const float* mask, void* y,
                         int32_t B, int32_t C, int32_t L, int32_t dtype, float slope, void* stream) {
    if (!x || !y) return rvc::fail(RVC_EINVAL, "rvc_pack_act: null pointer");
    if ((x2 == nullptr) != (x3 == nullptr)) return rvc::fail(RVC_EINVAL, "rvc_pack_act: x2 and x3 go together");
    if (dtype != RVC_DT_BF16 && dtype != RVC_DT_FP16) return rvc::fail(RVC_EINVAL, "rvc_pack_act: dtype must be RVC_DT_BF16 or RVC_DT_FP16");
    if (B <= 0 || B > 65535 || C <= 0 || C % 8 || L < 0) return rvc::fail(RVC_ESHAPE, "rvc_pack_act: bad shape (C=%d must be a multiple of 8)", C);
    if (L == 0) return 0;
    dim3 grid(rvc::cdiv(L, 256), C / 8, B);
    rvc::launch_pdl(pack_act_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, x, x2, x3, in_div, x_bs, x_cs, mask, (uint8_t*)y, C / 8, L, packed_rows(L), slope,
                                                            dtype == RVC_DT_FP16);
    return rvc::check_launch("rvc_pack_act");
}

RVC_API int rvc_pack_weight(const float* w, void* y, int32_t Cout, int32_t Cin, int32_t K, int32_t dtype, void* stream) {
    if (!w || !y) return rvc::fail(RVC_EINVAL, "rvc_pack_weight: null pointer");
    if (dtype != RVC_DT_BF16 && dtype != RVC_DT_FP16) return rvc::fail(RVC_EINVAL, "rvc_pack_weight: dtype must be RVC_DT_BF16 or RVC_DT_FP16");
    if (Cout <= 0 || Cin <= 0 || Cin % 8 || K <= 0) return rvc::fail(RVC_ESHAPE, "rvc_pack_weight: bad shape");
    const int64_t n = (int64_t)K * (Cin / 8) * Cout;
    pack_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (uint8_t*)y, Cout, Cin, K, dtype == RVC_DT_FP16);
    return rvc::check_launch("rvc_pack_weight");
}

RVC_API int rvc_conv1d_tc(const rvc_conv1d_tc_desc* dp, void* stream) {
    if (!dp) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: null descriptor");
    const rvc_conv1d_tc_desc& d = *dp;
    if (d.njobs < 1 || d.njobs > MAXJOBS) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: njobs must be 1..%d", MAXJOBS);
    const int G = d.sum_jobs <= 0 ? 1 : (d.sum_jobs == 1 ? d.njobs : d.sum_jobs);     // jobs per tile group
    if (d.njobs % G) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: sum_jobs = %d does not divide njobs = %d", d.sum_jobs, d.njobs);
    if (d.dtype != RVC_DT_BF16 && d.dtype != RVC_DT_FP16) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: dtype must be RVC_DT_BF16 or RVC_DT_FP16");
    if (d.B <= 0 || d.L <= 0 || d.Cin <= 0 || d.Cout <= 0) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: bad shape");
    if (d.Cin % 16) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: Cin=%d must be a multiple of 16", d.Cin);
    if (d.Cout % 16) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: Cout=%d must be a multiple of 16", d.Cout);
    Params p;
    int maxpad = 0, gated = d.jobs[0].gated ? 1 : 0;
    for (int j = 0; j < d.njobs; ++j) {
        const rvc_tc_job& jb = d.jobs[j];
        if (!jb.a || !jb.w) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: job %d has null a/w", j);
        if (jb.K <= 0 || jb.dil <= 0 || jb.pad < 0 || jb.pad > (jb.K - 1) * jb.dil)
            return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: job %d needs K > 0, dil > 0, 0 <= pad <= (K-1)*dil", j);
        if ((jb.gated ? 1 : 0) != gated) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: jobs of one launch must agree on `gated`");
        if (jb.gated && !jb.bias) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: gated job needs a bias");
        if (!(jb.act_slope >= 0.f && jb.act_slope <= 1.f)) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: act_slope must be in [0, 1]");
        if (jb.y_lo && (!jb.y_act || d.dtype != RVC_DT_FP16 || ((uintptr_t)jb.y_lo & 15)))
            return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: y_lo goes with y_act, fp16 operands and 16-byte alignment");
        if (jb.res_hi || jb.res_lo) {
            if (!jb.res_hi || !jb.res_lo || jb.res || jb.nz_har || gated || d.up > 1 || d.dtype != RVC_DT_FP16 || !(jb.act_slope > 0.f) ||
                (((uintptr_t)jb.res_hi | (uintptr_t)jb.res_lo) & 15) || jb.y_act_bs)
                return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: job %d: res_hi / res_lo come as a pair, replace res, and need fp16 operands, act_slope > 0, a plain conv "
                                             "and the default y_act layout", j);
        }
        if (jb.nz_har) {
            if (j != 0 || d.sum_jobs || gated || jb.res) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: the fused noise conv belongs to job 0, replaces res, and excludes sum_jobs / gated");
            if (!jb.nz_w || jb.nz_kn < 1 || jb.nz_kn > 4 || jb.nz_sn < 1 || jb.nz_pn < 0) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: fused noise conv needs 1 <= kn <= 4, sn >= 1, pn >= 0");
        }
        const int halo = (jb.K - 1) * jb.dil;
        if (jb.pad > PADL || halo - jb.pad > PADR)
            return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: halo %d/%d exceeds the %d zero rows of the packed layout", jb.pad, halo - jb.pad, PADL);
        maxpad = halo > maxpad ? halo : maxpad;
        p.jobs[j] = jb;
    }
    const int sms = sm_count_cached();
    // tile shape: the widest N tile (and 2 sub-tiles) that still yields >= one wave of CTAs; gated needs NT % 64 == 0
    int mt = d.mt, NT = d.nt;
    const int up = d.up > 0 ? d.up : 1;
    if (up > 1 && (d.pshift < 0 || d.pshift >= up + PADR)) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: bad pshift");
    // phase columns (up_cols = u > 1): a stride-u ConvTranspose as ONE plain conv whose Cout = u * Cr weight columns hold the u phases side by side
    const int ucols = d.up_cols > 1 ? d.up_cols : 1;
    int cr_shift = 0;
    if (ucols > 1) {
        const int Cr = d.Cout / ucols;
        while ((1 << cr_shift) < Cr) ++cr_shift;
        if (up > 1 || gated || G > 1 || d.Cout % ucols || (1 << cr_shift) != Cr || Cr < 32 || d.pshift < 0 || d.pshift >= ucols + PADR)
            return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: up_cols = %d needs up <= 1, no gate, no job groups, Cout = up_cols * (a power of two >= 32) and a valid pshift", ucols);
        for (int j = 0; j < d.njobs; ++j)
            if (d.jobs[j].res || d.jobs[j].res_hi) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: up_cols excludes residual inputs (the fused noise conv is allowed)");
    }
    const int ups = ucols > 1 ? ucols : up;       // output positions per input position
    // transposed modes: input positions q in [0, L + ceil(pshift/ups)) reach valid outputs (rows past L are zero in the packed buffer)
    const int Lq = ups > 1 ? d.L + rvc::cdiv(d.pshift, ups) : d.L;
    const int Lout = ups > 1 ? d.L * ups : d.L;
    const int njt = d.njobs / G;                // job groups multiply the tile count
    if (G > 1) {
        if (gated || up > 1) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: sum_jobs excludes gated / transposed mode");
        for (int j = 0; j < d.njobs; ++j)      // (a member's residual is added when it has one; a leader without one takes none at all)
            if (d.jobs[j].res && !d.jobs[j / G * G].res) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: sum_jobs: a member has res but its group leader has none");
        for (int j = 0; j < d.njobs; ++j)
            if (d.jobs[j].res_hi && (!d.jobs[j / G * G].res_hi || G > 3))
                return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: sum_jobs: split-fp16 residuals need a group leader that has one too and groups of at most three jobs");
        for (int j = 0; j < d.njobs; ++j)
            if (d.jobs[j / G * G].res_hi && d.jobs[j].res) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: sum_jobs: a group takes its residuals in one format");
    }
    auto ntiles = [&](int nt_, int mt_) { return (long)njt * d.B * rvc::cdiv(Lq, BM * mt_) * (d.Cout / nt_) * up; };
    // wide layers on short sequences (the C = 256 generator stage: L = 10-12 k) are bound by streaming the weights of each tile
    // from L2, not by the tensor pipe: 512-row x 64-column tiles re-use a weight stage four times and win even below one wave
    // (measured, 3 jobs k = 11/7/3, L = 12000: conv 50 -> 45 us, stage-average launch 101 -> 54 us)
    if (NT == 0 && mt == 0 && up == 1 && !gated && d.Cin >= 256 && d.Cout >= 256 && d.Cout % 64 == 0 && 2 * ntiles(64, 4) >= sms &&
        ntiles(128, 2) < 4L * sms) {
        NT = 64;
        mt = 4;
    }
    if (NT == 0) {
        const int cands[4] = {128, 64, 32, 16};
        for (int i = 0; i < 4; ++i) {
            const int c = cands[i];
            if (d.Cout % c || (gated && c % 64)) continue;
            if (c == 16 && NT != 0) break;          // N = 16 MMAs waste 60 % of the pipe: only when nothing wider divides Cout
            const long t = ntiles(c, 1);
            // one wave of wider tiles on >= 60 % of the SMs beats 1.x waves of narrower ones (192 -> 768, k = 3, T = 1000:
            // 96 tiles of 64 columns 7.0 us, 192 tiles of 32 columns 8.9 us)
            if (NT != 0 && t >= sms && 2 * t < 3L * sms && 5 * ntiles(NT, 1) >= 3L * sms) break;
            NT = c;
            if (t >= sms) break;
        }
        if (NT == 0) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: no N tile for Cout=%d (gated=%d)", d.Cout, gated);
    }
    if (NT != 128 && NT != 64 && NT != 32 && NT != 16) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: nt must be 16, 32, 64 or 128");
    if (d.Cout % NT || (gated && NT % 64)) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: nt=%d does not tile Cout=%d (gated=%d)", NT, d.Cout, gated);
    // sub-tiles per CTA tile: as many as still leave >= 2 waves of tiles (fewer, larger tiles amortise the per-tile handshakes);
    // 4 sub-tiles need NT <= 64 (TMEM: 2 groups x 4 x NT columns)
    // (weight bytes streamed per output halve with every doubling, so 2 sub-tiles are preferred as soon as they fill the GPU once)
    if (mt == 0) mt = (NT <= 64 && up == 1 && ntiles(NT, 4) >= 2L * sms) ? 4 : (ntiles(NT, 2) >= (long)sms ? 2 : 1);
    if (d.f32_blocked & ~3) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: f32_blocked must be 0..3");
    for (int j = 0; j < d.njobs && d.f32_blocked; ++j) {
        const rvc_tc_job& jb = d.jobs[j];
        if ((d.f32_blocked & 1) && jb.res && (((uintptr_t)jb.res & 31) || (jb.res_bs & 7) || (jb.res_cs & 7)))
            return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: blocked res needs a 32-byte aligned pointer and strides in multiples of 8");
        if ((d.f32_blocked & 2) && jb.y_f32 && (((uintptr_t)jb.y_f32 & 31) || (jb.y_bs & 7) || (jb.y_cs & 7)))
            return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: blocked y_f32 needs a 32-byte aligned pointer and strides in multiples of 8");
    }
    if (mt != 1 && mt != 2 && mt != 4) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: mt must be 1, 2 or 4");
    if (mt == 4 && NT > 64) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: mt = 4 needs nt <= 64");
    const int TM = BM * mt;
    p.njobs = d.njobs; p.B = d.B; p.Cin = d.Cin; p.Cout = d.Cout; p.L = Lq; p.Lp = packed_rows(d.L);
    p.up = up; p.pshift = ups > 1 ? d.pshift : 0; p.Lout = Lout; p.Lp_out = packed_rows(Lout);
    p.ucols = ucols; p.cr_shift = cr_shift;
    p.nch = d.Cin / 8;
    p.nch_out = (gated ? d.Cout / 2 : ucols > 1 ? d.Cout / ucols : d.Cout) / 8;
    p.rows_max = TM + maxpad;
    p.n_lt = rvc::cdiv(Lq, TM);
    p.n_nt = d.Cout / NT;
    p.tiles_per_job = d.B * p.n_lt * p.n_nt * up;
    p.total_tiles = p.tiles_per_job * njt;
    p.fd_tpj.init((uint32_t)p.tiles_per_job); p.fd_up.init((uint32_t)up); p.fd_nnt.init((uint32_t)p.n_nt); p.fd_nlt.init((uint32_t)p.n_lt);
    p.fp16 = d.dtype == RVC_DT_FP16;
    p.sum_jobs = d.sum_jobs > 0 ? 1 : 0;
    p.G = G;
    p.dbg = d.jobs[0].reserved;
    p.blocked = d.f32_blocked;
    p.has_res = 0;
    p.res_split = 0;
    for (int j = 0; j < d.njobs; ++j) {
        p.has_res |= d.jobs[j].res != nullptr || d.jobs[j].res_hi != nullptr;
        p.res_split |= d.jobs[j].res_hi != nullptr;
    }
    if (p.res_split)
        for (int j = 0; j < d.njobs; ++j)
            if (d.jobs[j].res) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: a launch takes its residuals in one format (res or res_hi / res_lo)");
    const size_t fixed = (size_t)NSTAGE * STAGE_BYTES + 512 + (size_t)(d.njobs / G) * d.Cout * sizeof(float) + (d.jobs[0].nz_har ? 16 + (size_t)d.Cout * 20 : 0);
    const size_t limit = 227 * 1024;
    // activation stage: as many channel chunks as fit twice (double buffer); even count
    int ach = p.nch;
    while (2 * (size_t)ach * p.rows_max * 16 + fixed > limit && ach > 2) ach = ((ach + 1) / 2 + 1) & ~1;
    if (2 * (size_t)ach * p.rows_max * 16 + fixed > limit) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: activation stage does not fit shared memory");
    p.ach = ach;
    p.ncs = rvc::cdiv(p.nch, ach);
    const size_t a_stage = (size_t)ach * p.rows_max * 16;
    int nabuf = (int)((limit - fixed) / a_stage);
    nabuf = nabuf > MAXABUF ? MAXABUF : nabuf;
    p.nabuf = nabuf;
    const size_t smem = (size_t)nabuf * a_stage + fixed;
    const int cap = rvc::grid_limit() > 0 && rvc::grid_limit() < sms ? rvc::grid_limit() : sms;
    const int grid = p.total_tiles < cap ? p.total_tiles : cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (NT == 128) return launch_tc_mt<128>(mt, p, smem, grid, st);
    if (NT == 64) return launch_tc_mt<64>(mt, p, smem, grid, st);
    if (NT == 32) return launch_tc_mt<32>(mt, p, smem, grid, st);
    return launch_tc_mt<16>(mt, p, smem, grid, st);
}
