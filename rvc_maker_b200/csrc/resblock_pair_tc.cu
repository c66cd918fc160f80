// One ResBlock step  x <- c2(lrelu(c1_d(lrelu(x)))) + x  (residuals.py:32-36) as ONE persistent tcgen05 kernel for the narrow
// generator stages (C = 32, 64), where the two separate conv launches are bound by the epilogue's HBM traffic, not the tensor pipe.
//
// Per tile of R = 256 rows (512 at C = 32) the intermediate t = lrelu(c1(lrelu(x)) + b1) never leaves the SM: phase 1 accumulates c1 into TMEM, the
// epilogue rounds t to fp16 straight into a shared-memory tile in the packed operand layout ([c/8][row][8], the K-major
// no-swizzle core-matrix layout), and phase 2 reads that tile as its A operand — a conv2 tap is again just a descriptor start
// address 16 B further down.  Saved against two launches: the packed write of t, its read back, one launch.  The tile computes
// t on rows [l0 - h2, l0 - h2 + R) and keeps the R - (K2-1) outputs that see only real t rows, so consecutive tiles overlap
// by K2-1 rows (4 % recompute at k = 11, R = 256).  t rows outside [0, L) are written as zeros: conv2's "same" padding.
//
// Roles as in conv1d_tc.cu (512 threads): warp 0 activation producer, warp 2 weight producer (w1 and w2 stream through one ring),
// warps 1/3 MMA issuers (one or two 128-row sub-tiles each), warps 4-15 epilogue.  Everything is double-buffered (activation tile, t tile,
// both accumulators) and the two phases are skewed by one tile — every role walks  P1(0), P1(1), P2(0), P1(2), P2(1), ... — so that
// the tensor pipe works on tile i+1's conv1 while the epilogue still stores tile i.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#ifndef RVC_PAIR_NSTAGE
#define RVC_PAIR_NSTAGE 4
#endif

namespace {

constexpr int BM = 128;
constexpr int PADL = 64;          // zero rows in front of position 0 in packed buffers
constexpr int ROWQ = 256;
constexpr int PADR = 64;
constexpr int RTX = 16;           // extra rows of the t tile: >= K2max - 1
constexpr int NSTAGE = RVC_PAIR_NSTAGE;
constexpr int STAGE_BYTES = 16384;
constexpr int NTHREADS = 512;
constexpr int NEG = 3;
constexpr int MAXJOBS = 3;

inline int packed_rows(int L) { return PADL + ((L + ROWQ - 1) / ROWQ) * ROWQ + PADR; }

struct PParams {
    rvc_pair_job jobs[MAXJOBS];
    int njobs, B, C, L, Lp, fp16;
    int ra;                          // rows per chunk of an activation stage: 256 + max_j (K1-1)*dil1
    int tiles_before[MAXJOBS + 1];   // prefix sums of the tiles per job
    int n_lt[MAXJOBS];
    int tmo[MAXJOBS];                // outputs kept per tile: 256 - (K2-1)
    int total_tiles;
    int sum_jobs;                    // every tile runs ALL jobs; conv2 of all jobs accumulates into one tile (stage average)
    int y_ncl;                       // sum mode: y is plain [B, C, L] instead of channel-blocked
    int dbg;                         // timing experiments only: 1 = no MMA issue, 2 = no epilogue-2 global I/O, 4 = no residual loads
    int split_in, split_out;         // split-fp16 residual stream: residual = inv_lrelu(a + a_lo);  outputs y_act (hi) + y_lo
    float slope, alpha, inv_slope;
};

// Debug timeline (tools/pair_trace.py), compiled in only with -DRVC_PAIR_TRACE (python -m rvc_maker_b200.build --trace): when a buffer is
// registered (rvc_debug_pair_trace), CTA 0 appends (event id, unit, SM clock) records — one lane per role — so that the producer / MMA /
// epilogue hand-offs of the fused step can be laid out on one time axis.  The product build compiles trace() to nothing.
#ifdef RVC_PAIR_TRACE
__device__ unsigned long long* g_trace = nullptr;
__device__ unsigned int g_trace_cap = 0;
// no atomics and no global reads per record (either costs a memory round trip and distorts what is measured): the kernel reads the buffer pointer
// once, record (ev, unit) has its own slot 80 * unit + 16 * role + k  (role 0 producer [ev 1], 1 MMA warp [ev 2..7], 2 + eg epilogue group
// [ev 10 + 3k + eg]); the store is fire-and-forget
#define TRACE_SETUP unsigned long long* const tr_buf = blockIdx.x == 0 ? g_trace : nullptr; const unsigned int tr_cap = g_trace_cap;
#define trace(ev, unit) do { if (tr_buf) { const int ev_ = (ev); const int role_ = ev_ == 1 ? 0 : ev_ < 10 ? 1 : 2 + (ev_ - 10) % 3; \
    const int k_ = ev_ == 1 ? 0 : ev_ < 10 ? ev_ - 2 : (ev_ - 10) / 3; const unsigned int i_ = 80u * (unsigned)(unit) + 16u * role_ + k_; \
    if (i_ < tr_cap) tr_buf[i_] = (unsigned long long)clock64(); } } while (0)
// a record that is taken only once register `x` has arrived (a branch on its value makes the warp wait for the load / tcgen05.ld behind it)
#define trace_after(ev, unit, x) do { if (tr_buf) { if (__float_as_uint(x) == 0x7fc12345u) tr_buf[0] = 0; trace(ev, unit); } } while (0)
#else
#define TRACE_SETUP
#define trace(ev, unit) do {} while (0)
#define trace_after(ev, unit, x) do {} while (0)
#endif

__device__ __forceinline__ uint32_t pack2h(float a, float b, int fp16) {
    uint32_t r;
    if (fp16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float lrelu_max(float v, float slope) { return fmaxf(v, v * slope); }
// inverse of leaky_relu for 0 < slope <= 1 (inv = 1 / slope >= 1): y > 0 -> y, y < 0 -> y / slope
__device__ __forceinline__ float lrelu_inv(float y, float inv) { return fminf(y, y * inv); }
__device__ __forceinline__ float2 h2f2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
using tc::ldg128;

// MT = 128-row sub-tiles per tile: 2, or 4 for C = 32 where TMEM holds twice the rows and the per-tile handshakes (about 2.3 us
// of barrier round trips per tile with no MMA and no stores at all) were the bound rather than HBM
// SUM / SPLIT (stage-average mode / split-fp16 residual stream) are template parameters: with the four modes behind run-time flags one kernel
// was 75 KB of code whose hot path hopped between distant fragments (tools/pair_trace.py: the epilogue warps lost 1-2 us per unit at the loop head)
template <int C, int MT, bool SUM, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS, 1) resblock_pair_kernel(const PParams p) {
#ifdef RVC_PAIR_TRACE
    const int dbg = dbg;                              // timing experiments exist in the trace build only
#else
    constexpr int dbg = 0;
#endif
    constexpr int TMR = BM * MT;                        // rows per tile
    constexpr int RT = TMR + RTX;                       // rows of the t tile
    constexpr int MPI = MT / 2;                         // sub-tiles per issuer warp
    constexpr int NCH = C / 8;                          // 16-byte channel chunks
    constexpr int SCH = STAGE_BYTES / (C * 16);         // chunk-rows per weight stage (even)
    constexpr int EB = 32;                              // accumulator columns per epilogue work item
    constexpr int NBLK = C / EB;
    constexpr int NITEM = MT * NBLK;                    // work items per phase and tile
    constexpr int TMEM_COLS = 4 * MT * C;               // 2 phases x 2 buffers x MT sub-tiles x C columns
    static_assert(TMEM_COLS <= 512 && (SCH % 2) == 0, "tile shape");

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t a_bytes = (uint32_t)NCH * p.ra * 16;
    constexpr uint32_t t_bytes = (uint32_t)NCH * RT * 16;
    uint8_t* a_smem = smem;                             // [2][NCH][ra][16]
    uint8_t* t_smem = smem + 2 * (size_t)a_bytes;       // [2][NCH][RT][16]
    uint8_t* w_smem = t_smem + 2 * (size_t)t_bytes;     // [NSTAGE][STAGE_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + NSTAGE * STAGE_BYTES);
    const uint32_t bar0 = tc::smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    auto A_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
    auto A_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 + s); };
    auto ACC1_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 4 + s); };
    auto ACC1_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 6 + s); };
    auto T_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 8 + s); };
    auto T_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 10 + s); };
    auto ACC2_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 12 + s); };
    auto ACC2_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 14 + s); };
    constexpr int NBAR = 2 * NSTAGE + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    float* bias_s = reinterpret_cast<float*>(bars + NBAR + 2);     // [njobs][2][C]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    TRACE_SETUP

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { tc::mbar_init(W_FULL(s), 1); tc::mbar_init(W_EMPTY(s), 2); }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(A_FULL(s), 1);      tc::mbar_init(A_EMPTY(s), 2);
            tc::mbar_init(ACC1_FULL(s), 2);   tc::mbar_init(ACC1_EMPTY(s), 4 * NEG);
            tc::mbar_init(T_FULL(s), 4 * NEG); tc::mbar_init(T_EMPTY(s), 2);
            tc::mbar_init(ACC2_FULL(s), 2);   tc::mbar_init(ACC2_EMPTY(s), 4 * NEG);
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
    if (warp >= 4) {
        for (int i = threadIdx.x - 128; i < p.njobs * 2 * C; i += NTHREADS - 128) {
            const int j = i / (2 * C), r = i - j * 2 * C;
            const float* src = r < C ? p.jobs[j].b1 : p.jobs[j].b2;
            float v = src ? src[r < C ? r : r - C] : 0.f;
            if (SUM && j == 0 && r >= C)       // the stage average adds every job's conv2 bias
                for (int jj = 1; jj < p.njobs; ++jj) v += p.jobs[jj].b2 ? p.jobs[jj].b2[r - C] : 0.f;
            bias_s[i] = v;
        }
        // rows 256.. of both t tiles are never written by the epilogue: keep them finite (they only reach discarded outputs)
        for (int i = threadIdx.x - 128; i < 2 * NCH * (RT - TMR); i += NTHREADS - 128) {
            const int buf = i / (NCH * (RT - TMR)), r = i - buf * NCH * (RT - TMR);
            const int ch = r / (RT - TMR), row = TMR + r - ch * (RT - TMR);
            *reinterpret_cast<uint4*>(t_smem + (size_t)buf * t_bytes + ((size_t)ch * RT + row) * 16) = make_uint4(0, 0, 0, 0);
        }
        tc::fence_proxy_async();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // tile -> (job, batch item, position tile).  Everything the decode needs sits in registers: the earlier version (a loop over p.tiles_before with
    // indexed constant loads, called twice per unit by every role) cost the epilogue warps 0.45 us per unit (tools/pair_trace.py)
    const int tb1 = p.tiles_before[1], tb2 = p.tiles_before[2];
    const int nl0 = p.n_lt[0], nl1 = p.n_lt[1], nl2 = p.n_lt[2];
    auto decode = [&](int tile, int& j, int& b, int& lt) {
        j = (tile >= tb1 ? 1 : 0) + (tile >= tb2 ? 1 : 0);
        const int r = tile - (j == 0 ? 0 : j == 1 ? tb1 : tb2);
        if (p.B == 1) { b = 0; lt = r; }
        else { const int nl = j == 0 ? nl0 : j == 1 ? nl1 : nl2; b = r / nl; lt = r - b * nl; }
    };
    // snake walk over the cost-sorted tile list (heavy jobs first), see conv1d_tc.cu: even waves deal tiles to CTA 0..P-1, odd ones back
    const int P = (int)gridDim.x, cta = (int)blockIdx.x;
    const int n_waves = p.total_tiles / P, rem = p.total_tiles - n_waves * P;
    const int n_my = n_waves + ((((n_waves & 1) ? P - 1 - cta : cta) < rem) ? 1 : 0);
    auto tile_of = [&](int it) { return it * P + ((it & 1) ? P - 1 - cta : cta); };
    // a *unit* is one (tile, job) conv pair: in the plain mode every tile is one unit of its own job; with sum_jobs a tile is
    // njobs consecutive units whose phase 2 accumulates into the same accumulator
    const int J = SUM ? p.njobs : 1;
    const int n_units = n_my * J;
    auto unit = [&](int n, int& it, int& jj, int& j, int& b, int& lt) {
        it = J == 1 ? n : J == 2 ? (n >> 1) : (int)(__umulhi((unsigned)n, 0xAAAAAAABu) >> 1);     // n / J, J <= 3
        jj = n - it * J;
        decode(tile_of(it), j, b, lt);
        if (SUM) j = jj;
    };

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ======================= activation producer: rows [l0 - h2 - pad1, + 256 + (K1-1)*dil1) of packed lrelu(x) =======================
        for (int n = 0; n < n_units; ++n) {
            int it, jj, j, b, lt;
            unit(n, it, jj, j, b, lt);
            const rvc_pair_job& jb = p.jobs[j];
            const int ast = n & 1;
            const uint32_t aph = (n >> 1) & 1;
            const int h2 = (jb.K2 - 1) / 2, pad1 = (jb.K1 - 1) / 2 * jb.dil1;
            const int row0 = PADL + lt * p.tmo[j] - h2 - pad1;                       // >= 0 (host checks h2 + pad1 <= PADL)
            int rows = TMR + (jb.K1 - 1) * jb.dil1;
            rows = min(rows, p.Lp - row0);                                           // the last tile must not read past the buffer
            if (lane == 0) {
                tc::mbar_wait_relaxed(A_EMPTY(ast), aph ^ 1);
                trace(1, n);
                tc::mbar_arrive_expect_tx(A_FULL(ast), (uint32_t)NCH * rows * 16);
            }
            __syncwarp();
            const uint8_t* asrc = reinterpret_cast<const uint8_t*>(jb.a) + (size_t)b * NCH * p.Lp * 16 + (size_t)row0 * 16;
            for (int c = lane; c < NCH; c += 32)
                tc::bulk_g2s(tc::smem_u32(a_smem + (size_t)ast * a_bytes + (size_t)c * p.ra * 16), asrc + (size_t)c * p.Lp * 16, (uint32_t)rows * 16,
                             A_FULL(ast));
        }
    } else if (warp == 2) {
        // ======================= weight producer: w1(0), w1(1), w2(0), w1(2), w2(1), ... through one ring =======================
        uint32_t wst = 0, wph = 0;
        auto stream = [&](const void* w, int K) {
            const int nflat = K * NCH;
            for (int f0 = 0; f0 < nflat; f0 += SCH) {
                const int nu = min(SCH, nflat - f0);
                if (lane == 0) {
                    tc::mbar_wait_relaxed(W_EMPTY(wst), wph ^ 1);
                    tc::mbar_arrive_expect_tx(W_FULL(wst), (uint32_t)nu * C * 16);
                    tc::bulk_g2s(tc::smem_u32(w_smem + wst * STAGE_BYTES), reinterpret_cast<const uint8_t*>(w) + (size_t)f0 * C * 16, (uint32_t)nu * C * 16,
                                 W_FULL(wst));
                }
                __syncwarp();
                if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
            }
        };
        int it = 0, jj = 0, j = 0, b = 0, lt = 0;
        for (int n = 0; n <= n_units; ++n) {
            const int pj = j;                           // job of unit n - 1
            if (n < n_units) { unit(n, it, jj, j, b, lt); stream(p.jobs[j].w1, p.jobs[j].K1); }
            if (n > 0) stream(p.jobs[pj].w2, p.jobs[pj].K2);
        }
    } else {
        // ======================= MMA issuers: warp 1 -> rows 0..127, warp 3 -> rows 128..255 of the tile =======================
        const int m = (warp == 1) ? 0 : 1;
        const uint32_t idesc = tc::make_idesc(p.fp16 ? 0u : 1u, C);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t b_lo_hi = ((uint32_t)C & 0x3FFFu) << 16;
        uint32_t wst = 0, wph = 0;
        // one conv phase: K taps x NCH/2 k-steps; A rows advance by `dil` per tap and 2*lbo_rows per k-step
        auto run_phase = [&](uint32_t d_tmem, uint32_t a_addr16, uint32_t lbo_rows, int K, uint32_t dil, uint32_t accumulate) {
            const uint32_t a_lo_hi = (lbo_rows & 0x3FFFu) << 16;
            const uint32_t da_base = a_lo_hi | ((a_addr16 + (uint32_t)(m * BM)) & 0x3FFFu);
            const int nflat = K * NCH;
            int ch = 0;
            uint32_t tap_off = 0, da_lo = da_base;
            for (int f0 = 0; f0 < nflat; f0 += SCH) {
                const int nu = min(SCH, nflat - f0);
                tc::mbar_wait(W_FULL(wst), wph);
                tc::fence_after_sync();
                uint32_t db_lo = b_lo_hi | ((tc::smem_u32(w_smem + wst * STAGE_BYTES) >> 4) & 0x3FFFu);
                int u = 0;
                while (u < nu) {
                    const int run = min(nu - u, NCH - ch);
                    for (int r = 0; r < run; r += 2) {
                        if (!(dbg & 1) && tc::elect_one()) {
                            tc::umma<false>(d_tmem, ((uint64_t)desc_hi << 32) | da_lo, ((uint64_t)desc_hi << 32) | db_lo, idesc, accumulate);
                            if constexpr (MPI == 2)   // sub-tile m + 2: same weights, rows 2*128 further down
                                tc::umma<false>(d_tmem + 2u * C, ((uint64_t)desc_hi << 32) | (da_lo + 2u * BM), ((uint64_t)desc_hi << 32) | db_lo, idesc, accumulate);
                        }
                        accumulate = 1;
                        da_lo += 2u * lbo_rows;
                        db_lo += 2u * C;
                    }
                    u += run;
                    ch += run;
                    if (ch == NCH) { ch = 0; tap_off += dil; da_lo = da_base + tap_off; }
                }
                if (tc::elect_one()) tc::umma_commit(W_EMPTY(wst));
                if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
            }
        };
        int it = 0, jj = 0, j = 0, b = 0, lt = 0;
        for (int n = 0; n <= n_units; ++n) {
            const int pit = it, pjj = jj, pj = j;       // unit n - 1: decoded in the previous round
            if (n < n_units) {   // ---- phase 1 of unit n
                unit(n, it, jj, j, b, lt);
                const int g = n & 1;
                const uint32_t gph = (n >> 1) & 1;
                tc::mbar_wait(ACC1_EMPTY(g), gph ^ 1);
                if (m == 0 && lane == 0) trace(2, n);
                tc::mbar_wait(A_FULL(g), gph);
                if (m == 0 && lane == 0) trace(3, n);
                tc::fence_after_sync();
                run_phase(tmem_base + (uint32_t)((g * MT + m) * C), tc::smem_u32(a_smem + (size_t)g * a_bytes) >> 4, (uint32_t)p.ra, p.jobs[j].K1,
                          (uint32_t)p.jobs[j].dil1, 0u);
                if (tc::elect_one()) { tc::umma_commit(A_EMPTY(g)); tc::umma_commit(ACC1_FULL(g)); }
                if (m == 0 && lane == 0) trace(4, n);
                __syncwarp();
            }
            if (n > 0) {         // ---- phase 2 of unit n - 1 (accumulates over the jobs of a tile with sum_jobs)
                const int g = (n - 1) & 1;
                const uint32_t gph = ((n - 1) >> 1) & 1;
                const int g2 = pit & 1;
                if (pjj == 0) tc::mbar_wait(ACC2_EMPTY(g2), ((pit >> 1) & 1) ^ 1);
                if (m == 0 && lane == 0) trace(5, n - 1);
                tc::mbar_wait(T_FULL(g), gph);
                if (m == 0 && lane == 0) trace(6, n - 1);
                tc::fence_after_sync();
                run_phase(tmem_base + (uint32_t)(2 * MT * C + (g2 * MT + m) * C), tc::smem_u32(t_smem + (size_t)g * t_bytes) >> 4, (uint32_t)RT, p.jobs[pj].K2, 1u,
                          pjj ? 1u : 0u);
                if (tc::elect_one()) {
                    tc::umma_commit(T_EMPTY(g));
                    if (pjj == J - 1) tc::umma_commit(ACC2_FULL(g2));
                }
                if (m == 0 && lane == 0) trace(7, n - 1);
                __syncwarp();
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        // ======================= epilogue warps =======================
        // Step `it` holds the work items of epilogue 1 of tile it (k = 0 .. NITEM-1) and of epilogue 2 of tile it-1
        // (k = NITEM .. 2*NITEM-1) in ONE list shared round-robin by the three warpgroups (k = eg, eg+3, ...): 8 items over 3
        // warpgroups cost 3 item times, two separate 4-item phases would cost 4.  The residual of a warp's next epilogue-2 item is
        // requested before it turns to its epilogue-1 items (and before the accumulator wait), so its latency is hidden.
        const int q = warp & 3;
        const int eg = (warp - 4) >> 2;
        const size_t cs = (size_t)p.L * 8;                               // stride between 8-channel blocks of the fp32 streams
        const int first2 = eg + NEG * ((NITEM - eg + NEG - 1) / NEG);    // this warpgroup's first epilogue-2 item index (>= NITEM)
        float vn[EB];
        int it1 = 0, jj1 = 0, j1 = 0, b1 = 0, lt1 = 0;
        for (int n = 0; n <= n_units; ++n) {
            if (q == 0 && lane == 0) trace(28 + eg, n);
            const bool has1 = n < n_units;
            int it2 = it1, jj2 = jj1, j2 = j1, b2 = b1, lt2 = lt1;      // unit n - 1: decoded in the previous round
            if (has1) unit(n, it1, jj1, j1, b1, lt1);
            const bool has2 = n > 0 && jj2 == J - 1;      // unit n-1 closed its tile: the accumulator of tile it2 is complete
            if (SUM) j2 = 0;                       // outputs / summed bias of the stage average live in job 0
            const rvc_pair_job& jb2 = p.jobs[j2];
            const int l02 = lt2 * p.tmo[j2];
            if (q == 0 && lane == 0) trace_after(40 + eg, n, __int_as_float(l02 + lt1));
            auto res_load = [&](int k, float* dst) {      // residual(s) of epilogue-2 item k - NITEM of tile it2
                const int item = k - NITEM;
                const int m = item / NBLK, cb = item - m * NBLK;
                const int r2 = m * BM + q * 32 + lane;
                const int l = l02 + r2;
                if (!(r2 < p.tmo[j2] && l < p.L) || (dbg & 6)) return;
                if (SPLIT) {
                    // split-fp16 stream: hi (= the conv operand plane, usually still in L2 from this tile's activation load) + lo
                    const size_t poff = ((size_t)b2 * NCH + (size_t)(cb * (EB / 8))) * p.Lp + PADL + l;     // 16-byte units
                    if (!SUM) {   // raw bits now, decoded when the item is processed: the loads stay in flight meanwhile
                        const uint4* ah = reinterpret_cast<const uint4*>(jb2.a) + poff;
                        const uint4* al = reinterpret_cast<const uint4*>(jb2.a_lo) + poff;
#pragma unroll
                        for (int u = 0; u < EB / 8; ++u) {
                            const uint4 h = ldg128(ah + (size_t)u * p.Lp), lo = ldg128(al + (size_t)u * p.Lp);
                            dst[8 * u + 0] = __uint_as_float(h.x); dst[8 * u + 1] = __uint_as_float(h.y);
                            dst[8 * u + 2] = __uint_as_float(h.z); dst[8 * u + 3] = __uint_as_float(h.w);
                            dst[8 * u + 4] = __uint_as_float(lo.x); dst[8 * u + 5] = __uint_as_float(lo.y);
                            dst[8 * u + 6] = __uint_as_float(lo.z); dst[8 * u + 7] = __uint_as_float(lo.w);
                        }
                    } else {             // stage average: x_0 + x_1 + x_2, all six loads of a chunk in flight before the first add
                        const int j1i = 1 < p.njobs ? 1 : 0, j2i = 2 < p.njobs ? 2 : 0;
                        const float s1 = p.njobs > 1 ? 1.f : 0.f, s2 = p.njobs > 2 ? 1.f : 0.f;
                        const float inv = p.inv_slope;
#pragma unroll
                        for (int u = 0; u < EB / 8; ++u) {
                            const size_t o = poff + (size_t)u * p.Lp;
                            const uint4 h0 = ldg128(reinterpret_cast<const uint4*>(p.jobs[0].a) + o), l0 = ldg128(reinterpret_cast<const uint4*>(p.jobs[0].a_lo) + o);
                            const uint4 h1 = ldg128(reinterpret_cast<const uint4*>(p.jobs[j1i].a) + o), l1 = ldg128(reinterpret_cast<const uint4*>(p.jobs[j1i].a_lo) + o);
                            const uint4 h2 = ldg128(reinterpret_cast<const uint4*>(p.jobs[j2i].a) + o), l2 = ldg128(reinterpret_cast<const uint4*>(p.jobs[j2i].a_lo) + o);
                            const uint32_t hw0[4] = {h0.x, h0.y, h0.z, h0.w}, lw0[4] = {l0.x, l0.y, l0.z, l0.w};
                            const uint32_t hw1[4] = {h1.x, h1.y, h1.z, h1.w}, lw1[4] = {l1.x, l1.y, l1.z, l1.w};
                            const uint32_t hw2[4] = {h2.x, h2.y, h2.z, h2.w}, lw2[4] = {l2.x, l2.y, l2.z, l2.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 a0 = h2f2(hw0[i]), b0 = h2f2(lw0[i]), a1 = h2f2(hw1[i]), b1_ = h2f2(lw1[i]), a2 = h2f2(hw2[i]), b2_ = h2f2(lw2[i]);
                                dst[8 * u + 2 * i] = (lrelu_inv(a0.x + b0.x, inv) + s1 * lrelu_inv(a1.x + b1_.x, inv)) + s2 * lrelu_inv(a2.x + b2_.x, inv);
                                dst[8 * u + 2 * i + 1] = (lrelu_inv(a0.y + b0.y, inv) + s1 * lrelu_inv(a1.y + b1_.y, inv)) + s2 * lrelu_inv(a2.y + b2_.y, inv);
                            }
                        }
                    }
                    return;
                }
                const size_t off = (size_t)b2 * C * p.L + (size_t)(cb * (EB / 8)) * cs + (size_t)l * 8;
#pragma unroll
                for (int u = 0; u < EB / 8; ++u) tc::ldg256(jb2.res + off + (size_t)u * cs, dst + 8 * u);
                if (SUM) {
                    // the other jobs' residuals: all loads of a chunk are issued before the first add (a per-job loop serialised
                    // three DRAM round trips per item and made the stage-average epilogue the critical path)
                    const float* r1 = p.jobs[1 < p.njobs ? 1 : 0].res + off;
                    const float* r2 = p.jobs[2 < p.njobs ? 2 : 0].res + off;
                    const float s1 = p.njobs > 1 ? 1.f : 0.f, s2 = p.njobs > 2 ? 1.f : 0.f;
#pragma unroll
                    for (int u = 0; u < EB / 8; ++u) {
                        float a8[8], b8[8];
                        tc::ldg256(r1 + (size_t)u * cs, a8);
                        tc::ldg256(r2 + (size_t)u * cs, b8);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dst[8 * u + i] = (dst[8 * u + i] + s1 * a8[i]) + s2 * b8[i];
                    }
                }
            };
            if constexpr (SUM) {      // (the stage average converts its six planes as they arrive; that wait overlaps best ahead of epilogue 1)
                if (has2 && first2 < 2 * NITEM) res_load(first2, vn);
            }
            if (has1) {
                // ---- epilogue 1 of unit n: t = lrelu(acc1 + b1) -> fp16 -> shared-memory operand tile (zeros outside [0, L))
                const rvc_pair_job& jb = p.jobs[j1];
                const int g = n & 1;
                const uint32_t gph = (n >> 1) & 1;
                const int l0 = lt1 * p.tmo[j1] - (jb.K2 - 1) / 2;        // position of t row 0
                if (q == 0 && lane == 0) trace(10 + eg, n);
                tc::mbar_wait(ACC1_FULL(g), gph);
                if (q == 0 && lane == 0) trace(13 + eg, n);
                tc::mbar_wait(T_EMPTY(g), gph ^ 1);
                if (q == 0 && lane == 0) trace(16 + eg, n);
                tc::fence_after_sync();
                uint8_t* tb = t_smem + (size_t)g * t_bytes;
#pragma unroll 1
                for (int item = eg; item < NITEM; item += NEG) {
                    const int m = item / NBLK, cb = item - m * NBLK;
                    const int row = m * BM + q * 32 + lane;
                    const int pos = l0 + row;
                    const bool inside = pos >= 0 && pos < p.L;
                    uint32_t r[EB];
                    tc::tmem_ld<EB>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g * MT + m) * C + cb * EB), r);
                    tc::tmem_ld_wait();
                    const float4* bs4 = reinterpret_cast<const float4*>(bias_s + (j1 * 2) * C + cb * EB);
#pragma unroll
                    for (int u = 0; u < EB / 8; ++u) {
                        const float4 c0 = bs4[2 * u], c1 = bs4[2 * u + 1];
                        uint4 pk;
                        pk.x = pack2h(lrelu_max(__uint_as_float(r[8 * u + 0]) + c0.x, p.slope), lrelu_max(__uint_as_float(r[8 * u + 1]) + c0.y, p.slope), p.fp16);
                        pk.y = pack2h(lrelu_max(__uint_as_float(r[8 * u + 2]) + c0.z, p.slope), lrelu_max(__uint_as_float(r[8 * u + 3]) + c0.w, p.slope), p.fp16);
                        pk.z = pack2h(lrelu_max(__uint_as_float(r[8 * u + 4]) + c1.x, p.slope), lrelu_max(__uint_as_float(r[8 * u + 5]) + c1.y, p.slope), p.fp16);
                        pk.w = pack2h(lrelu_max(__uint_as_float(r[8 * u + 6]) + c1.z, p.slope), lrelu_max(__uint_as_float(r[8 * u + 7]) + c1.w, p.slope), p.fp16);
                        if (!inside) pk = make_uint4(0, 0, 0, 0);
                        *reinterpret_cast<uint4*>(tb + ((size_t)(cb * (EB / 8) + u) * RT + row) * 16) = pk;
                    }
                }
                tc::fence_before_sync();
                tc::fence_proxy_async();          // the t tile is read by the tensor core through the async proxy
                __syncwarp();
                if (lane == 0) { tc::mbar_arrive(T_FULL(g)); tc::mbar_arrive(ACC1_EMPTY(g)); }
                if (q == 0 && lane == 0) trace(19 + eg, n);
            }
            // The residual of this warp's first epilogue-2 item is requested AFTER its epilogue-1 items: right after the store burst of the previous
            // epilogue 2 the load/store path is backed up and the eight loads take 1-2 us to issue (tools/pair_trace.py) — time during which the
            // MMA warps would wait for the t tile; and the proxy fence that ends epilogue 1 (MEMBAR.ALL.CTA) would wait for the loads to return.
            if (!SUM && has2 && first2 < 2 * NITEM) {
                if (q == 0 && lane == 0) trace(46 + eg, n);
                res_load(first2, vn);
                if (q == 0 && lane == 0) trace(49 + eg, n);
            }
            if (has2) {
                // ---- epilogue 2 of tile it2: x' = alpha * (acc2 + b2 + x)  ->  fp32 and packed lrelu(x')
                const int g = it2 & 1;
                const uint32_t gph = (it2 >> 1) & 1;
                tc::mbar_wait(ACC2_FULL(g), gph);
                if (q == 0 && lane == 0) trace(22 + eg, n - 1);
                tc::fence_after_sync();
#pragma unroll 1
                for (int k = first2; k < 2 * NITEM; k += NEG) {
                    const int item = k - NITEM;
                    const int m = item / NBLK, cb = item - m * NBLK;
                    const int r2 = m * BM + q * 32 + lane;
                    const int l = l02 + r2;
                    const bool ok = r2 < p.tmo[j2] && l < p.L && !(dbg & 2);
                    float v[EB];
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] = vn[i];
                    if (k == first2 && q == 0 && lane == 0) trace_after(31 + eg, n - 1, v[EB - 1] + v[EB - 5]);      // hi and lo words of the last chunk
                    if (k + NEG < 2 * NITEM) res_load(k + NEG, vn);
                    uint32_t r[EB];
                    tc::tmem_ld<EB>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * MT * C + (g * MT + m) * C + cb * EB), r);
                    tc::tmem_ld_wait();
                    if (k == first2 && q == 0 && lane == 0) trace_after(34 + eg, n - 1, __uint_as_float(r[EB - 1]));
                    if (!ok) continue;
                    if (SPLIT && !SUM) {   // raw hi / lo words -> x = inv_lrelu(hi + lo)
                        const float inv = p.inv_slope;
#pragma unroll
                        for (int u = 0; u < EB / 8; ++u) {
                            uint32_t w[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[8 * u + i]);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 h = h2f2(w[i]), lo = h2f2(w[4 + i]);
                                v[8 * u + 2 * i] = lrelu_inv(h.x + lo.x, inv);
                                v[8 * u + 2 * i + 1] = lrelu_inv(h.y + lo.y, inv);
                            }
                        }
                    }
                    const float4* bs4 = reinterpret_cast<const float4*>(bias_s + (j2 * 2 + 1) * C + cb * EB);
#pragma unroll
                    for (int i4 = 0; i4 < EB / 4; ++i4) {
                        const float4 b4 = bs4[i4];
                        v[4 * i4 + 0] += __uint_as_float(r[4 * i4 + 0]) + b4.x;
                        v[4 * i4 + 1] += __uint_as_float(r[4 * i4 + 1]) + b4.y;
                        v[4 * i4 + 2] += __uint_as_float(r[4 * i4 + 2]) + b4.z;
                        v[4 * i4 + 3] += __uint_as_float(r[4 * i4 + 3]) + b4.w;
                    }
                    if (SUM) {
#pragma unroll
                        for (int i = 0; i < EB; ++i) v[i] *= p.alpha;
                    }
                    if (jb2.y) {
                        if (p.y_ncl) {
                            float* yp = jb2.y + (size_t)b2 * C * p.L + (size_t)(cb * EB) * p.L + l;
#pragma unroll
                            for (int i = 0; i < EB; ++i) yp[(size_t)i * p.L] = v[i];
                        } else {
                            float* yp = jb2.y + (size_t)b2 * C * p.L + (size_t)(cb * (EB / 8)) * cs + (size_t)l * 8;
#pragma unroll
                            for (int u = 0; u < EB / 8; ++u) tc::stg256(yp + (size_t)u * cs, v + 8 * u);
                        }
                    }
                    if ((SPLIT && !SUM)) {
                        // split-fp16 stream out: hi = fp16(lrelu(x')) (the next conv1's operand), lo = fp16(lrelu(x') - hi)
                        const size_t poff = ((size_t)b2 * NCH * p.Lp + ((size_t)cb * (EB / 8)) * p.Lp + PADL + l) * 16;
                        uint8_t* yh = reinterpret_cast<uint8_t*>(jb2.y_act) + poff;
                        uint8_t* yl = reinterpret_cast<uint8_t*>(jb2.y_lo) + poff;
#pragma unroll
                        for (int u = 0; u < EB / 8; ++u) {
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float f0 = lrelu_max(v[8 * u + 2 * i], p.slope), f1 = lrelu_max(v[8 * u + 2 * i + 1], p.slope);
                                hw[i] = pack2h(f0, f1, 1);
                                const float2 hf = h2f2(hw[i]);
                                lw[i] = pack2h(f0 - hf.x, f1 - hf.y, 1);
                            }
                            tc::stg128(yh + (size_t)u * p.Lp * 16, make_uint4(hw[0], hw[1], hw[2], hw[3]));
                            tc::stg128(yl + (size_t)u * p.Lp * 16, make_uint4(lw[0], lw[1], lw[2], lw[3]));
                        }
                    } else if (jb2.y_act) {
                        uint8_t* yb = reinterpret_cast<uint8_t*>(jb2.y_act) + (size_t)b2 * NCH * p.Lp * 16 + (((size_t)cb * (EB / 8)) * p.Lp + PADL + l) * 16;
#pragma unroll
                        for (int u = 0; u < EB / 8; ++u) {
                            uint4 pk;
                            pk.x = pack2h(lrelu_max(v[8 * u + 0], p.slope), lrelu_max(v[8 * u + 1], p.slope), p.fp16);
                            pk.y = pack2h(lrelu_max(v[8 * u + 2], p.slope), lrelu_max(v[8 * u + 3], p.slope), p.fp16);
                            pk.z = pack2h(lrelu_max(v[8 * u + 4], p.slope), lrelu_max(v[8 * u + 5], p.slope), p.fp16);
                            pk.w = pack2h(lrelu_max(v[8 * u + 6], p.slope), lrelu_max(v[8 * u + 7], p.slope), p.fp16);
                            tc::stg128(yb + (size_t)u * p.Lp * 16, pk);
                        }
                    }
                    if (k == first2 && q == 0 && lane == 0) trace(37 + eg, n - 1);
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(ACC2_EMPTY(g));
                if (q == 0 && lane == 0) trace(25 + eg, n - 1);
            }
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

int sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev] > 0 ? sms[dev] : 148;
}

template <int C, int MT, bool SUM, bool SPLIT>
int launch_pair(const PParams& p, size_t smem, int grid, cudaStream_t stream) {
    auto kern = resblock_pair_kernel<C, MT, SUM, SPLIT>;
    static bool opted[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_resblock_pair_tc: smem opt-in: %s", cudaGetErrorString(e));
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
    if (le != cudaSuccess) return rvc::fail((int)le, "rvc_resblock_pair_tc: launch: %s", cudaGetErrorString(le));
    return rvc::check_launch("rvc_resblock_pair_tc");
}

}  // namespace

RVC_API int rvc_resblock_pair_tc(const rvc_resblock_pair_desc* dp, void* stream) {
    if (!dp) return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: null descriptor");
    const rvc_resblock_pair_desc& d = *dp;
    if (d.njobs < 1 || d.njobs > MAXJOBS) return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: njobs must be 1..3");
    if (d.dtype != RVC_DT_BF16 && d.dtype != RVC_DT_FP16) return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: dtype must be RVC_DT_BF16 or RVC_DT_FP16");
    if (d.C != 32 && d.C != 64) return rvc::fail(RVC_ESHAPE, "rvc_resblock_pair_tc: C=%d (the fused pair covers C = 32 and 64)", d.C);
    if (d.B <= 0 || d.L <= 0) return rvc::fail(RVC_ESHAPE, "rvc_resblock_pair_tc: bad shape");
    if (!(d.slope >= 0.f && d.slope <= 1.f)) return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: slope must be in [0, 1]");
    const int TMR = BM * (d.C == 32 ? 4 : 2), RT = TMR + RTX;     // C = 32: four sub-tiles per tile (TMEM: 4 * MT * C columns)
    PParams p = {};
    p.njobs = d.njobs; p.B = d.B; p.C = d.C; p.L = d.L; p.Lp = packed_rows(d.L); p.fp16 = d.dtype == RVC_DT_FP16; p.slope = d.slope;
    p.sum_jobs = d.sum_jobs ? 1 : 0; p.y_ncl = d.y_ncl ? 1 : 0; p.alpha = d.sum_jobs ? d.alpha : 1.f;
    p.dbg = d.reserved;
    p.split_in = d.jobs[0].a_lo != nullptr;
    p.split_out = !d.sum_jobs && d.jobs[0].y_lo != nullptr;
    p.inv_slope = d.slope > 0.f ? 1.f / d.slope : 1.f;
    int halo1 = 0, k2max = 1;
    long total = 0;
    for (int j = 0; j < d.njobs; ++j) {
        const rvc_pair_job& jb = d.jobs[j];
        const bool outs = !d.sum_jobs || j == 0;      // the stage average writes through job 0 only
        if (!jb.a || !jb.w1 || !jb.w2 || !(jb.res || jb.a_lo)) return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: job %d has a null pointer", j);
        if ((jb.a_lo != nullptr) != (d.jobs[0].a_lo != nullptr) || (!d.sum_jobs && (jb.y_lo != nullptr) != (d.jobs[0].y_lo != nullptr)))
            return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: the jobs of one launch must agree on the split-fp16 stream (a_lo / y_lo)");
        if ((jb.a_lo || (!d.sum_jobs && jb.y_lo)) && (d.dtype != RVC_DT_FP16 || !(d.slope > 0.f)))
            return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: the split-fp16 stream needs dtype = RVC_DT_FP16 and slope > 0");
        if (!d.sum_jobs && (!(jb.y || jb.y_lo) || !jb.y_act)) return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: job %d needs y_act and y (or y_lo)", j);
        if (!d.sum_jobs && ((jb.a_lo != nullptr) != (jb.y_lo != nullptr)))
            return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: job %d: a pair step keeps its stream format (a_lo and y_lo go together)", j);
        // (y_lo may alias a_lo: the lo plane is read only at a tile's own output rows, by the thread that rewrites them)
        if (outs && !d.sum_jobs && jb.y_lo && ((uintptr_t)jb.y_lo & 15))
            return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: job %d: y_lo must be 16-byte aligned", j);
        if (jb.a_lo && ((uintptr_t)jb.a_lo & 15)) return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: job %d: a_lo needs 16-byte alignment", j);
        if (d.sum_jobs && j == 0 && !jb.y && !jb.y_act) return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: sum_jobs needs y or y_act on job 0");
        if (outs && jb.a == jb.y_act)
            return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: job %d: y_act must not alias a (tiles read a with halos while others write y_act)", j);
        if (jb.K1 <= 0 || !(jb.K1 & 1) || jb.K2 <= 0 || !(jb.K2 & 1) || jb.dil1 <= 0)
            return rvc::fail(RVC_ESHAPE, "rvc_resblock_pair_tc: job %d needs odd kernel sizes and dil1 > 0", j);
        if (jb.K2 - 1 > RTX) return rvc::fail(RVC_ESHAPE, "rvc_resblock_pair_tc: K2=%d exceeds %d", jb.K2, RTX + 1);
        if ((jb.K1 - 1) / 2 * jb.dil1 + (jb.K2 - 1) / 2 > PADL)
            return rvc::fail(RVC_ESHAPE, "rvc_resblock_pair_tc: halo %d exceeds the %d zero rows of the packed layout", (jb.K1 - 1) / 2 * jb.dil1 + (jb.K2 - 1) / 2, PADL);
        if ((((uintptr_t)jb.res | (uintptr_t)(outs ? jb.y : nullptr)) & 31) ||
            (((uintptr_t)jb.a | (uintptr_t)jb.w1 | (uintptr_t)jb.w2 | (uintptr_t)(outs ? jb.y_act : nullptr)) & 15))
            return rvc::fail(RVC_EINVAL, "rvc_resblock_pair_tc: job %d: fp32 streams need 32-byte, packed buffers 16-byte alignment", j);
        halo1 = (jb.K1 - 1) * jb.dil1 > halo1 ? (jb.K1 - 1) * jb.dil1 : halo1;
        k2max = jb.K2 > k2max ? jb.K2 : k2max;
        p.jobs[j] = jb;
    }
    for (int j = 0; j < d.njobs; ++j) {
        p.tmo[j] = TMR - ((d.sum_jobs ? k2max : d.jobs[j].K2) - 1);     // with sum_jobs every job shares the tile grid of the widest conv2
        p.n_lt[j] = rvc::cdiv(d.L, p.tmo[j]);
        p.tiles_before[j] = (int)total;
        if (!d.sum_jobs || j == 0) total += (long)d.B * p.n_lt[j];
    }
    if (total > 0x7fffffffL) return rvc::fail(RVC_ESHAPE, "rvc_resblock_pair_tc: too many tiles");
    for (int j = d.njobs; j <= MAXJOBS; ++j) p.tiles_before[j] = (int)total;
    if (d.sum_jobs)
        for (int j = 1; j <= MAXJOBS; ++j) p.tiles_before[j] = (int)total;   // decode() then always lands on job 0
    p.total_tiles = (int)total;
    p.ra = TMR + halo1;
    const int nch = d.C / 8;
    const size_t smem = 2 * (size_t)nch * p.ra * 16 + 2 * (size_t)nch * RT * 16 + (size_t)NSTAGE * STAGE_BYTES + 512 + (size_t)d.njobs * 2 * d.C * sizeof(float);
    if (smem > 227 * 1024) return rvc::fail(RVC_ESHAPE, "rvc_resblock_pair_tc: tile does not fit shared memory (%zu bytes)", smem);
    const int sms = sm_count();
    const int cap = rvc::grid_limit() > 0 && rvc::grid_limit() < sms ? rvc::grid_limit() : sms;
    const int grid = p.total_tiles < cap ? p.total_tiles : cap;
    const cudaStream_t st = (cudaStream_t)stream;
    const int mode = (p.sum_jobs ? 2 : 0) | (p.split_in ? 1 : 0);
    if (d.C == 64) {
        switch (mode) {
            case 0: return launch_pair<64, 2, false, false>(p, smem, grid, st);
            case 1: return launch_pair<64, 2, false, true>(p, smem, grid, st);
            case 2: return launch_pair<64, 2, true, false>(p, smem, grid, st);
            default: return launch_pair<64, 2, true, true>(p, smem, grid, st);
        }
    }
    switch (mode) {
        case 0: return launch_pair<32, 4, false, false>(p, smem, grid, st);
        case 1: return launch_pair<32, 4, false, true>(p, smem, grid, st);
        case 2: return launch_pair<32, 4, true, false>(p, smem, grid, st);
        default: return launch_pair<32, 4, true, true>(p, smem, grid, st);
    }
}

// Debug only (tools/pair_trace.py): register a device buffer of `capacity` 64-bit records for the timeline of CTA 0 (NULL / 0 switches it off);
// slot 30 * unit + 6 * role + k holds the SM clock of that event (0 = not reached).  Returns 0, or RVC_EINVAL when the library was built without -DRVC_PAIR_TRACE.
RVC_API int rvc_debug_pair_trace(void* dev_buf, int32_t capacity) {
#ifdef RVC_PAIR_TRACE
    unsigned int cap = dev_buf ? (unsigned int)capacity : 0u;
    unsigned long long* ptr = (unsigned long long*)dev_buf;
    cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
    cudaMemcpyToSymbol(g_trace, &ptr, sizeof(ptr));
    return 0;
#else
    (void)dev_buf; (void)capacity;
    return rvc::fail(RVC_EINVAL, "rvc_debug_pair_trace: this build has no trace points (python -m rvc_maker_b200.build --trace)");
#endif
}
