// A CHAIN of dependent tensor-core Conv1d launches (conv1d_tc.cu) as ONE persistent kernel with a grid-wide barrier between the
// phases — for the short-sequence layers (the flow's 32 convs, the FFN pairs of the TextEncoder) whose launches are bound by
// per-launch latency, not by throughput: at T = 1000 a layer is 24-96 tiles, each launch costs 8-9 us inside the captured graph
// (launch + prologue + TMEM allocation + weight-stream latency + drain) for ~1-2 us of tensor work.  Here
//   * barriers, TMEM and the phase table are set up once;
//   * the weight producer is never stopped by the barrier: it streams the NEXT phase's (constant) weights through the ring while
//     the epilogue of the current phase is still storing, so the weight-stream latency leaves the critical path;
//   * a phase boundary costs one grid barrier (atomic counter in L2) instead of a kernel boundary.
// Same data layout, warp roles and epilogue arithmetic as conv1d_tc.cu with the tile shape fixed at 128 positions x 64 columns
// (results are bit-identical to the separate launches with nt = 64, mt = 1).  Not supported in a chain: transposed (phase) mode,
// the fused noise conv, per-batch weights.  Cout is padded to a multiple of 64 by the caller's weight packing; `cout_valid`
// channels are stored.
//
// Grid barrier: every epilogue thread fences its global writes, the 12 epilogue warps meet on a named barrier, one thread adds
// 1 to a global counter and spins (ld.acquire.gpu) until all CTAs of the grid arrived, fences again (gpu scope: drops this
// SM's L1 lines) and publishes the phase number in shared memory; the activation producer (and every epilogue warp, whose
// residual loads read what other CTAs wrote) waits for that number, the producer adds a generic->async proxy fence before its
// bulk copies.  All CTAs are co-resident (grid <= SM count, one CTA per SM), spins are bounded (about 1 s) so that a logic error
// shows up as a wrong result and a set flag instead of a hung GPU.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <string.h>

namespace {

constexpr int BM = 128;
constexpr int NT = 64;
constexpr int PADL = 64;
constexpr int PADR = 64;
constexpr int ROWQ = 256;
constexpr int NSTAGE = 4;
constexpr int STAGE_BYTES = 16384;
constexpr int NTHREADS = 512;
constexpr int NEG = 3;
constexpr int CJOBS = 3;
constexpr int MAXABUF = 8;
constexpr int NG = 4;                       // accumulator groups: 4 x 64 TMEM columns
constexpr int TMEM_COLS = NG * NT;
constexpr int SCH = STAGE_BYTES / (NT * 16);   // 16-byte chunks per weight stage
constexpr int EB = 32;
constexpr int NBLK = NT / EB;
constexpr int MAXPHASES = 48;

inline int packed_rows(int L) { return PADL + ((L + ROWQ - 1) / ROWQ) * ROWQ + PADR; }

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

struct CPhase {
    rvc_tc_job jobs[CJOBS];
    int njobs, G, B, Cin, Cout, cout_valid, L, Lp;
    int nch, nch_out, ach, ncs, rows_max, n_lt, n_nt, tiles_per_job, total_tiles;
    FastDiv fd_tpj, fd_nnt, fd_nlt;
    int fp16, sum_jobs, blocked, pad_[3];
};
static_assert(sizeof(CPhase) % 16 == 0, "phase table rows are copied as 16-byte units");

__device__ __forceinline__ uint32_t pack2(float a, float b, int fp16) {
    uint32_t r;
    if (fp16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// same arithmetic as conv1d_tc.cu (bit-identical results)
__device__ __forceinline__ float gate_fast(float t, float s) {
    const float e2 = __expf(2.f * fminf(fmaxf(t, -15.f), 15.f));
    const float es = __expf(-fmaxf(s, -30.f));
    return __fdividef(e2 - 1.f, (e2 + 1.f) * (1.f + es));
}
__device__ __forceinline__ float lrelu_max(float v, float slope) { return fmaxf(v, v * slope); }

__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_cta_shared(uint32_t addr) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t addr, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
constexpr long long SPIN_LIMIT = 1ll << 31;   // cycles (about 1 s): a barrier that never completes ends as an error flag, not a hang

__global__ void __launch_bounds__(NTHREADS, 1) conv1d_tc_chain_kernel(const CPhase* __restrict__ gphases, int nph, int a_slot_bytes, int nabuf,
                                                                      int bias_cap, unsigned* __restrict__ sync) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_smem = smem;                                          // [nabuf][a_slot_bytes]
    uint8_t* w_smem = smem + (size_t)nabuf * a_slot_bytes;           // [NSTAGE][STAGE_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + NSTAGE * STAGE_BYTES);
    const uint32_t bar0 = tc::smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    auto A_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
    auto A_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + MAXABUF + s); };
    auto ACC_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * MAXABUF + s); };
    auto ACC_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * MAXABUF + NG + s); };
    constexpr int NBAR = 2 * NSTAGE + 2 * MAXABUF + 2 * NG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);
    const uint32_t go_addr = tc::smem_u32(tmem_slot + 1);            // number of completed phases, published after each grid barrier
    float* bias_s = reinterpret_cast<float*>(bars + NBAR + 2);       // [2][bias_cap]
    CPhase* phases = reinterpret_cast<CPhase*>(bias_s + 2 * bias_cap);   // [nph] (16-byte aligned: bias_cap % 4 == 0)

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    {   // phase table -> shared memory (constant data: may overlap the previous kernel's tail)
        const uint4* src = reinterpret_cast<const uint4*>(gphases);
        uint4* dst = reinterpret_cast<uint4*>(phases);
        const int n16 = nph * (int)(sizeof(CPhase) / 16);
        for (int i = threadIdx.x; i < n16; i += NTHREADS) dst[i] = src[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { tc::mbar_init(W_FULL(s), 1); tc::mbar_init(W_EMPTY(s), 1); }
        for (int s = 0; s < MAXABUF; ++s) { tc::mbar_init(A_FULL(s), 1); tc::mbar_init(A_EMPTY(s), 1); }
        for (int s = 0; s < NG; ++s) { tc::mbar_init(ACC_FULL(s), 1); tc::mbar_init(ACC_EMPTY(s), 4 * NEG); }
        tmem_slot[1] = 0;
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int P = (int)gridDim.x, cta = (int)blockIdx.x;
    auto decode = [&](const CPhase& p, int tile, int& j, int& b, int& lt, int& nt) {
        uint32_t q, r, t;
        p.fd_tpj.divmod((uint32_t)tile, q, r);
        j = (int)q;
        p.fd_nnt.divmod(r, q, t);
        nt = (int)t;
        p.fd_nlt.divmod(q, r, t);
        lt = (int)t;
        b = (int)r;
    };
    // snake walk over a phase's tiles (see conv1d_tc.cu)
    auto n_mine = [&](const CPhase& p) {
        const int n_waves = p.total_tiles / P, rem = p.total_tiles - n_waves * P;
        return n_waves + ((((n_waves & 1) ? P - 1 - cta : cta) < rem) ? 1 : 0);
    };
    auto tile_of = [&](int it) { return it * P + ((it & 1) ? P - 1 - cta : cta); };
    // wait until `k` phases have completed on the whole grid
    auto wait_go = [&](int k) {
        if (k <= 0) return;
        const long long t0 = clock64();
        while (ld_acquire_cta_shared(go_addr) < k)
            if (clock64() - t0 > SPIN_LIMIT) break;
    };

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ======================= activation producer =======================
        uint32_t ast = 0, aph = 0;
        for (int ph = 0; ph < nph; ++ph) {
            const CPhase& p = phases[ph];
            const int n_my = n_mine(p);
            if (n_my == 0) continue;
            if (lane == 0) {
                wait_go(ph);                                          // the previous phase's outputs are complete grid-wide ...
                asm volatile("fence.proxy.async;" ::: "memory");      // ... and ordered before this thread's async-proxy (bulk copy) reads
            }
            __syncwarp();
            for (int it = 0; it < n_my; ++it) {
                int j, b, lt, nt;
                decode(p, tile_of(it), j, b, lt, nt);
                for (int jj = 0; jj < p.G; ++jj) {
                    const rvc_tc_job& jb = p.jobs[j * p.G + jj];
                    const int row0 = PADL + lt * BM - jb.pad;
                    const int rows = min(BM + (jb.K - 1) * jb.dil, p.Lp - row0);
                    for (int cs = 0; cs < p.ncs; ++cs) {
                        const int c0 = cs * p.ach;
                        const int nc = min(p.ach, p.nch - c0);
                        if (lane == 0) {
                            tc::mbar_wait(A_EMPTY(ast), aph ^ 1);
                            tc::mbar_arrive_expect_tx(A_FULL(ast), (uint32_t)nc * rows * 16);
                        }
                        __syncwarp();
                        const size_t a_bs = jb.a_bs ? (size_t)jb.a_bs : (size_t)p.nch * p.Lp * 16;
                        const uint8_t* asrc = reinterpret_cast<const uint8_t*>(jb.a) + (size_t)b * a_bs + ((size_t)c0 * p.Lp + (size_t)row0) * 16;
                        for (int c = lane; c < nc; c += 32)
                            tc::bulk_g2s(tc::smem_u32(a_smem + (size_t)ast * a_slot_bytes + (size_t)c * p.rows_max * 16), asrc + (size_t)c * p.Lp * 16,
                                         (uint32_t)rows * 16, A_FULL(ast));
                        if (++ast == (uint32_t)nabuf) { ast = 0; aph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ======================= weight producer: constant data, runs ahead of the grid barrier =======================
        uint32_t wst = 0, wph = 0;
        for (int ph = 0; ph < nph; ++ph) {
            const CPhase& p = phases[ph];
            const int n_my = n_mine(p);
            for (int it = 0; it < n_my; ++it) {
                int j, b, lt, nt;
                decode(p, tile_of(it), j, b, lt, nt);
                for (int jj = 0; jj < p.G; ++jj) {
                    const rvc_tc_job& jb = p.jobs[j * p.G + jj];
                    const size_t w_cs = jb.w_cs ? (size_t)jb.w_cs : (size_t)p.Cout;
                    const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(jb.w);
                    for (int cs = 0; cs < p.ncs; ++cs) {
                        const int ca = cs * p.ach;
                        const int na = min(p.ach, p.nch - ca);
                        const int nflat = jb.K * na;
                        for (int f0 = 0; f0 < nflat; f0 += SCH) {
                            const int nu = min(SCH, nflat - f0);
                            if (lane == 0) {
                                tc::mbar_wait(W_EMPTY(wst), wph ^ 1);
                                tc::mbar_arrive_expect_tx(W_FULL(wst), (uint32_t)nu * NT * 16);
                            }
                            __syncwarp();
                            const uint32_t dst = tc::smem_u32(w_smem + wst * STAGE_BYTES);
                            for (int u = lane; u < nu; u += 32) {
                                const int f = f0 + u, tap = f / na, ch = f - tap * na;
                                tc::bulk_g2s(dst + u * NT * 16, wsrc + (((size_t)tap * p.nch + ca + ch) * w_cs + (size_t)nt * NT) * 16, NT * 16, W_FULL(wst));
                            }
                            if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (uniform loop, one elected lane issues) =======================
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t b_lo_hi = ((uint32_t)NT & 0x3FFFu) << 16;
        uint32_t wst = 0, wph = 0, ast = 0, aph = 0;
        int itg = 0;                                                   // tiles issued so far (all phases): accumulator group counter
        for (int ph = 0; ph < nph; ++ph) {
            const CPhase& p = phases[ph];
            const int n_my = n_mine(p);
            const uint32_t idesc = tc::make_idesc(p.fp16 ? 0u : 1u, NT);
            const uint32_t a_lo_hi = ((uint32_t)p.rows_max & 0x3FFFu) << 16;
            const uint32_t rows2 = 2u * (uint32_t)p.rows_max;
            for (int it = 0; it < n_my; ++it, ++itg) {
                int j, b, lt, nt;
                decode(p, tile_of(it), j, b, lt, nt);
                const int g = itg % NG;
                const uint32_t gph = (itg / NG) & 1;
                tc::mbar_wait(ACC_EMPTY(g), gph ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(g * NT);
                uint32_t accumulate = 0;
                for (int jj = 0; jj < p.G; ++jj) {
                    const int K = p.jobs[j * p.G + jj].K;
                    const uint32_t dil = (uint32_t)p.jobs[j * p.G + jj].dil;
                    for (int cs = 0; cs < p.ncs; ++cs) {
                        const int na = min(p.ach, p.nch - cs * p.ach);
                        tc::mbar_wait(A_FULL(ast), aph);
                        tc::fence_after_sync();
                        const uint32_t da_base = a_lo_hi | ((tc::smem_u32(a_smem + (size_t)ast * a_slot_bytes) >> 4) & 0x3FFFu);
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
                                const int run = min(nu - u, na - ch);
                                for (int r = 0; r < run; r += 2) {
                                    if (tc::elect_one())
                                        tc::umma<false>(d_tmem, ((uint64_t)desc_hi << 32) | da_lo, ((uint64_t)desc_hi << 32) | db_lo, idesc, accumulate);
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
                        if (++ast == (uint32_t)nabuf) { ast = 0; aph ^= 1; }
                    }
                }
                if (tc::elect_one()) tc::umma_commit(ACC_FULL(g));
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        // ======================= epilogue warps (+ the grid barrier) =======================
        const int q = warp & 3;
        const int eg = (warp - 4) >> 2;
        const int et = threadIdx.x - 128;      // 0 .. 383
        int itg = 0;
        for (int ph = 0; ph < nph; ++ph) {
            const CPhase& p = phases[ph];
            const int n_my = n_mine(p);
            float* bs = bias_s + (ph & 1) * bias_cap;
            // this phase's (group-summed) biases: constant data, staged before the wait.  Buffer ph & 1 was last read in phase ph - 2,
            // which every epilogue warp left before the named barrier at the end of phase ph - 1.
            for (int i = et; i < (p.njobs / p.G) * p.Cout; i += NTHREADS - 128) {
                const int g = i / p.Cout, c = i - g * p.Cout;
                float v = 0.f;
                if (c < p.cout_valid)
                    for (int jj = 0; jj < p.G; ++jj) v += p.jobs[g * p.G + jj].bias ? p.jobs[g * p.G + jj].bias[c] : 0.f;
                bs[i] = v;
            }
            asm volatile("bar.sync 1, 384;" ::: "memory");
            wait_go(ph);       // residual / mask loads below read what other CTAs wrote in the previous phase
            for (int it = 0; it < n_my; ++it, ++itg) {
                int j, b, lt, nt;
                decode(p, tile_of(it), j, b, lt, nt);
                const int jl = j * p.G;
                const rvc_tc_job& jb = p.jobs[jl];
                const int g = itg % NG;
                const uint32_t gph = (itg / NG) & 1;
                const float* bb = jb.bias_b ? jb.bias_b + (size_t)b * jb.bias_b_bs : nullptr;
                const int nblk = jb.gated ? NBLK / 2 : NBLK;
                tc::mbar_wait(ACC_FULL(g), gph);
                tc::fence_after_sync();
#pragma unroll 1
                for (int cb = eg; cb < nblk; cb += NEG) {
                    const int l = lt * BM + q * 32 + lane;
                    const int col0 = nt * NT + cb * EB;
                    const int co0 = jb.gated ? nt * (NT / 2) + cb * EB : col0;
                    const bool ok = l < p.L && co0 < p.cout_valid;
                    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * NT);
                    uint32_t r[EB];
                    float v[EB];
#pragma unroll
                    for (int i = 0; i < EB; ++i) v[i] = 0.f;
                    if (jb.res && ok) {
                        if (p.blocked & 1) {
                            const float* rp = jb.res + (size_t)b * jb.res_bs + (size_t)(co0 >> 3) * jb.res_cs + (size_t)l * 8;
#pragma unroll
                            for (int u = 0; u < EB / 8; ++u) tc::ldg256(rp + (size_t)u * jb.res_cs, v + 8 * u);
                        } else {
                            const float* rp = jb.res + (size_t)b * jb.res_bs + (size_t)co0 * jb.res_cs + l;
#pragma unroll
                            for (int i = 0; i < EB; ++i) v[i] = rp[(size_t)i * jb.res_cs];
                        }
                        for (int jj = 1; jj < p.G; ++jj) {
                            const rvc_tc_job& j2 = (&jb)[jj];
                            if (!j2.res) continue;
                            if (p.blocked & 1) {
                                const float* rp2 = j2.res + (size_t)b * j2.res_bs + (size_t)(co0 >> 3) * j2.res_cs + (size_t)l * 8;
#pragma unroll
                                for (int u = 0; u < EB / 8; ++u) {
                                    float t8[8];
                                    tc::ldg256(rp2 + (size_t)u * j2.res_cs, t8);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) v[8 * u + i] += t8[i];
                                }
                            } else {
                                const float* rp2 = j2.res + (size_t)b * j2.res_bs + (size_t)co0 * j2.res_cs + l;
#pragma unroll
                                for (int i = 0; i < EB; ++i) v[i] += rp2[(size_t)i * j2.res_cs];
                            }
                        }
                    }
                    const float mk = (jb.mask && ok) ? jb.mask[(size_t)b * p.L + l] : 1.f;
                    tc::tmem_ld<EB>(t_row + cb * EB, r);
                    tc::tmem_ld_wait();
                    const float4* bs4 = reinterpret_cast<const float4*>(bs + j * p.Cout + col0);
                    if (jb.gated) {
                        uint32_t r2[EB];
                        tc::tmem_ld<EB>(t_row + (cb + NBLK / 2) * EB, r2);
                        tc::tmem_ld_wait();
                        const int col1 = col0 + (NBLK / 2) * EB;
                        const float* bs1 = bs + j * p.Cout + col1;
#pragma unroll
                        for (int i = 0; i < EB; ++i) {
                            float t = __uint_as_float(r[i]) + bs[j * p.Cout + col0 + i];
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
                    if (jb.relu == 1) {
#pragma unroll
                        for (int i = 0; i < EB; ++i) r[i] = __float_as_uint(fmaxf(__uint_as_float(r[i]), 0.f));
                    }
                    const float alpha = jb.alpha;
                    if (p.sum_jobs) {
#pragma unroll
                        for (int i = 0; i < EB; ++i) v[i] = alpha * ((jb.res ? v[i] : 0.f) + __uint_as_float(r[i]));
                    } else if (jb.res) {
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
                    if (jb.relu == 2) {
#pragma unroll
                        for (int i = 0; i < EB; ++i) v[i] = lrelu_max(v[i], sl);
                        sl = 1.f;
                    }
                    if (jb.mask) {
#pragma unroll
                        for (int i = 0; i < EB; ++i) v[i] *= mk;
                    }
                    if (jb.y_f32) {
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
                    if (jb.y_act) {
                        const size_t ya_bs = jb.y_act_bs ? (size_t)jb.y_act_bs : (size_t)p.nch_out * p.Lp * 16;
                        uint8_t* yb = reinterpret_cast<uint8_t*>(jb.y_act) + (size_t)b * ya_bs + (((size_t)co0 / 8) * p.Lp + PADL + l) * 16;
#pragma unroll
                        for (int u = 0; u < EB / 8; ++u) {
                            uint4 pk;
                            pk.x = pack2(lrelu_max(v[8 * u + 0], sl), lrelu_max(v[8 * u + 1], sl), p.fp16);
                            pk.y = pack2(lrelu_max(v[8 * u + 2], sl), lrelu_max(v[8 * u + 3], sl), p.fp16);
                            pk.z = pack2(lrelu_max(v[8 * u + 4], sl), lrelu_max(v[8 * u + 5], sl), p.fp16);
                            pk.w = pack2(lrelu_max(v[8 * u + 6], sl), lrelu_max(v[8 * u + 7], sl), p.fp16);
                            *reinterpret_cast<uint4*>(yb + (size_t)u * p.Lp * 16) = pk;
                            if (jb.y_lo) {
                                const uint32_t hw[4] = {pk.x, pk.y, pk.z, pk.w};
                                uint32_t lw[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
                                    lw[i] = pack2(lrelu_max(v[8 * u + 2 * i], sl) - hf.x, lrelu_max(v[8 * u + 2 * i + 1], sl) - hf.y, 1);
                                }
                                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(jb.y_lo) + (yb - reinterpret_cast<uint8_t*>(jb.y_act)) + (size_t)u * p.Lp * 16) =
                                    make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            }
                        }
                    }
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(ACC_EMPTY(g));
            }
            // ---- end of phase: publish this CTA's writes, meet the rest of the grid (nothing follows the last phase but the kernel's end)
            if (ph + 1 < nph) {
                __threadfence();
                asm volatile("bar.sync 1, 384;" ::: "memory");
                if (et == 0) {
                    atomicAdd(sync, 1u);
                    const unsigned target = (unsigned)(ph + 1) * (unsigned)P;
                    const long long t0 = clock64();
                    while (ld_acquire_gpu(sync) < target)
                        if (clock64() - t0 > SPIN_LIMIT) { atomicExch(sync + 2, 1u); break; }
                    __threadfence();                                  // gpu-scope fence: later loads of this SM do not hit stale L1 lines
                    st_release_cta_shared(go_addr, ph + 1);
                }
            }
        }
        // the last CTA through resets the counter for the next launch (every CTA has passed every barrier by then)
        if (et == 0 && nph > 1) {
            __threadfence();
            if (atomicAdd(sync + 1, 1u) == (unsigned)P - 1) {
                atomicExch(sync, 0u);
                atomicExch(sync + 1, 0u);
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

struct ChainHeader {   // first 64 bytes of the blob
    int32_t magic, nph, grid, smem, a_slot, nabuf, bias_cap, pad_[9];
};
static_assert(sizeof(ChainHeader) == 64, "header");
constexpr int32_t CHAIN_MAGIC = 0x52564343;

}  // namespace

RVC_API int64_t rvc_conv1d_tc_chain_blob_bytes(int32_t n) { return (int64_t)sizeof(ChainHeader) + (int64_t)(n > 0 ? n : 0) * (int64_t)sizeof(CPhase); }

RVC_API int rvc_conv1d_tc_chain_build(const rvc_conv1d_tc_desc* descs, const int32_t* cout_valid, int32_t n, void* blob, int64_t blob_cap) {
    if (!descs || !blob) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc_chain_build: null pointer");
    if (n < 1 || n > MAXPHASES) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc_chain_build: 1..%d phases", MAXPHASES);
    if (blob_cap < rvc_conv1d_tc_chain_blob_bytes(n)) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc_chain_build: blob too small");
    ChainHeader h;
    memset(&h, 0, sizeof(h));
    CPhase* ph = reinterpret_cast<CPhase*>(reinterpret_cast<uint8_t*>(blob) + sizeof(ChainHeader));
    const int sms = sm_count();
    int max_tiles = 0, max_astage = 0, bias_cap = 4;
    for (int i = 0; i < n; ++i) {
        const rvc_conv1d_tc_desc& d = descs[i];
        CPhase& p = ph[i];
        memset(&p, 0, sizeof(p));
        if (d.njobs < 1 || d.njobs > CJOBS) return rvc::fail(RVC_EINVAL, "chain phase %d: njobs must be 1..%d", i, CJOBS);
        if (d.dtype != RVC_DT_BF16 && d.dtype != RVC_DT_FP16) return rvc::fail(RVC_EINVAL, "chain phase %d: dtype must be RVC_DT_BF16 or RVC_DT_FP16", i);
        if (d.B <= 0 || d.L <= 0 || d.Cin <= 0 || d.Cout <= 0 || d.Cin % 16 || d.Cout % NT)
            return rvc::fail(RVC_ESHAPE, "chain phase %d: needs Cin %% 16 == 0 and Cout %% 64 == 0 (pad the weights), got %d -> %d", i, d.Cin, d.Cout);
        if (d.up > 1 || d.up_cols > 1 || (d.nt && d.nt != NT) || (d.mt && d.mt != 1)) return rvc::fail(RVC_EINVAL, "chain phase %d: plain convs with 128 x 64 tiles only", i);
        if (d.f32_blocked & ~3) return rvc::fail(RVC_EINVAL, "chain phase %d: f32_blocked must be 0..3", i);
        const int G = d.sum_jobs <= 0 ? 1 : (d.sum_jobs == 1 ? d.njobs : d.sum_jobs);
        if (d.njobs % G) return rvc::fail(RVC_EINVAL, "chain phase %d: sum_jobs does not divide njobs", i);
        const int cv = cout_valid && cout_valid[i] > 0 ? cout_valid[i] : d.Cout;
        if (cv > d.Cout || cv % EB) return rvc::fail(RVC_ESHAPE, "chain phase %d: cout_valid must be a multiple of 32 and <= Cout", i);
        int maxpad = 0;
        const int gated = d.jobs[0].gated ? 1 : 0;
        for (int j = 0; j < d.njobs; ++j) {
            const rvc_tc_job& jb = d.jobs[j];
            if (!jb.a || !jb.w) return rvc::fail(RVC_EINVAL, "chain phase %d: job %d has null a/w", i, j);
            if (jb.w_bs) return rvc::fail(RVC_EINVAL, "chain phase %d: per-batch weights are not supported in a chain", i);
            if (jb.nz_har) return rvc::fail(RVC_EINVAL, "chain phase %d: the fused noise conv is not supported in a chain", i);
            if (jb.K <= 0 || jb.dil <= 0 || jb.pad < 0 || jb.pad > (jb.K - 1) * jb.dil) return rvc::fail(RVC_ESHAPE, "chain phase %d: bad K / dil / pad", i);
            if ((jb.gated ? 1 : 0) != gated || (jb.gated && (!jb.bias || G > 1))) return rvc::fail(RVC_EINVAL, "chain phase %d: gated jobs need a bias, agree, and exclude sum_jobs", i);
            if (!(jb.act_slope >= 0.f && jb.act_slope <= 1.f)) return rvc::fail(RVC_EINVAL, "chain phase %d: act_slope must be in [0, 1]", i);
            if (jb.y_lo && (!jb.y_act || d.dtype != RVC_DT_FP16)) return rvc::fail(RVC_EINVAL, "chain phase %d: y_lo goes with y_act and fp16", i);
            if (jb.res && !d.jobs[j / G * G].res) return rvc::fail(RVC_EINVAL, "chain phase %d: a member has res but its group leader has none", i);
            const int halo = (jb.K - 1) * jb.dil;
            if (jb.pad > PADL || halo - jb.pad > PADR) return rvc::fail(RVC_ESHAPE, "chain phase %d: halo exceeds the zero rows of the packed layout", i);
            maxpad = halo > maxpad ? halo : maxpad;
            p.jobs[j] = jb;
        }
        p.njobs = d.njobs; p.G = G; p.B = d.B; p.Cin = d.Cin; p.Cout = d.Cout; p.cout_valid = cv; p.L = d.L; p.Lp = packed_rows(d.L);
        p.nch = d.Cin / 8;
        p.nch_out = (gated ? cv / 2 : cv) / 8;
        p.rows_max = BM + maxpad;
        p.n_lt = rvc::cdiv(d.L, BM);
        p.n_nt = d.Cout / NT;
        p.tiles_per_job = d.B * p.n_lt * p.n_nt;
        p.total_tiles = p.tiles_per_job * (d.njobs / G);
        p.fd_tpj.init((uint32_t)p.tiles_per_job); p.fd_nnt.init((uint32_t)p.n_nt); p.fd_nlt.init((uint32_t)p.n_lt);
        p.fp16 = d.dtype == RVC_DT_FP16;
        p.sum_jobs = d.sum_jobs > 0 ? 1 : 0;
        p.blocked = d.f32_blocked;
        max_tiles = p.total_tiles > max_tiles ? p.total_tiles : max_tiles;
        const int bc = (d.njobs / G) * d.Cout;
        bias_cap = bc > bias_cap ? bc : bias_cap;
    }
    bias_cap = (bias_cap + 3) & ~3;
    const size_t fixed = (size_t)NSTAGE * STAGE_BYTES + 512 + 2 * (size_t)bias_cap * sizeof(float) + (size_t)n * sizeof(CPhase);
    const size_t limit = 227 * 1024;
    // activation stage per phase: all channel chunks when two slots fit, otherwise channel-staged (even chunk count); one slot size for the chain
    for (int i = 0; i < n; ++i) {
        CPhase& p = ph[i];
        int ach = p.nch;
        while (2 * (size_t)ach * p.rows_max * 16 + fixed > limit && ach > 2) ach = ((ach + 1) / 2 + 1) & ~1;
        if (2 * (size_t)ach * p.rows_max * 16 + fixed > limit) return rvc::fail(RVC_ESHAPE, "chain phase %d: activation stage does not fit shared memory", i);
        p.ach = ach;
        p.ncs = rvc::cdiv(p.nch, ach);
        const int st = ach * p.rows_max * 16;
        max_astage = st > max_astage ? st : max_astage;
    }
    int nabuf = (int)((limit - fixed) / (size_t)max_astage);
    nabuf = nabuf > MAXABUF ? MAXABUF : nabuf;
    if (nabuf < 2) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc_chain_build: activation ring does not fit");
    h.magic = CHAIN_MAGIC; h.nph = n; h.grid = max_tiles < sms ? max_tiles : sms; h.a_slot = max_astage; h.nabuf = nabuf; h.bias_cap = bias_cap;
    h.smem = (int32_t)((size_t)nabuf * max_astage + fixed);
    if (h.smem < 120 * 1024) h.smem = 120 * 1024;     // one CTA per SM: the grid barrier needs every CTA resident, each with its TMEM columns
    memcpy(blob, &h, sizeof(h));
    return 0;
}

RVC_API int rvc_conv1d_tc_chain_run(const void* host_header, const void* dev_blob, uint32_t* dev_sync, void* stream) {
    if (!host_header || !dev_blob || !dev_sync) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc_chain_run: null pointer");
    ChainHeader h;
    memcpy(&h, host_header, sizeof(h));
    if (h.magic != CHAIN_MAGIC) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc_chain_run: not a chain blob");
    if (((uintptr_t)dev_blob & 15) || ((uintptr_t)dev_sync & 3)) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc_chain_run: misaligned blob / counter");
    auto kern = conv1d_tc_chain_kernel;
    static bool opted[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return rvc::fail((int)e, "rvc_conv1d_tc_chain_run: smem opt-in: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) opted[dev] = true;
    }
    if (h.grid > sm_count()) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc_chain_run: the chain was built for a larger GPU");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h.grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = (size_t)h.smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const CPhase* phases = reinterpret_cast<const CPhase*>(reinterpret_cast<const uint8_t*>(dev_blob) + sizeof(ChainHeader));
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, phases, (int)h.nph, (int)h.a_slot, (int)h.nabuf, (int)h.bias_cap, (unsigned*)dev_sync);
    if (le != cudaSuccess) return rvc::fail((int)le, "rvc_conv1d_tc_chain_run: launch: %s", cudaGetErrorString(le));
    return rvc::check_launch("rvc_conv1d_tc_chain_run");
}
