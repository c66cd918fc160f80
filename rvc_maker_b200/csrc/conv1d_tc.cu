// ResBlock Conv1d (C -> C, kernel K, dilation d, "same" padding) as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[128 positions, NT out-channels] += A_tap[128 positions, 16|8 in-channels] * W_tap[NT, 16|8]^T   (tcgen05.mma, fp32 in TMEM)
//
// Data layout ("C8/C4 packed", DT = bf16 or tf32-rounded fp32):  act[b][c / CK][PADL + l][CK],  CK = 16 B / sizeof(DT).
// One 16-byte unit therefore holds CK consecutive channels of ONE position, and consecutive positions are 16 B apart,
// which is exactly the K-major no-swizzle UMMA core-matrix layout (8 rows x 16 B contiguous).  A conv tap that reads
// position l + j*dil - pad is the SAME shared-memory tile with the descriptor start address advanced by 16*j*dil bytes,
// so the activation tile (256 + halo positions, all input channels) is staged ONCE per CTA tile with plain 1-D bulk
// async copies (one per 16-byte channel chunk; zero rows around the sequence live in the global buffer, so there is no
// boundary logic) and reused by every tap.  Weights, pre-packed [K][C/CK][Cout][CK], stream through a 4-slot ring.
//
// Warp roles (192 threads): warp 0 = bulk-copy producer, warp 1 = MMA issuer (+TMEM owner), warps 2-5 = epilogue
// (TMEM -> registers -> bias / residual / leaky_relu -> fp32 NCL and packed-DT stores, coalesced along positions).
// Accumulators are double-buffered in TMEM (2 groups x 2 sub-tiles x NT columns) so the epilogue of tile t overlaps
// the MMAs of tile t+1; CTAs are persistent over a static round-robin tile list spanning up to three independent
// conv jobs (the three ResBlocks of a stage).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;          // positions per accumulator (TMEM lanes)
constexpr int MT = 2;            // accumulators (sub-tiles) per CTA tile
constexpr int TM = BM * MT;      // positions per CTA tile
constexpr int PADL = 64;         // zero rows in front of position 0 in packed buffers
constexpr int PADR = 64;
constexpr int NSTAGE = 4;        // weight ring slots
constexpr int STAGE_BYTES = 16384;
constexpr int NTHREADS = 192;
constexpr int MAXJOBS = 3;

__host__ __device__ inline int packed_rows(int L) { return PADL + ((L + TM - 1) / TM) * TM + PADR; }

struct Params {
    rvc_tc_job jobs[MAXJOBS];
    int njobs, B, C, L, Lp;
    int nch;          // C / CK: 16-byte channel chunks
    int rows_max;     // A-tile rows allocated per chunk (TM + max halo)
    int n_lt, n_nt;   // tiles along L and along Cout
    int tiles_per_job, total_tiles;
    int abuf;         // number of A buffers (1 or 2)
};

template <bool TF32, int NT>
__global__ void __launch_bounds__(NTHREADS, 1) conv1d_tc_kernel(const Params p) {
    constexpr int SCH = STAGE_BYTES / (NT * 16);          // 16-byte chunks per weight stage
    constexpr uint32_t IDESC = tc::make_idesc(TF32 ? 2u : 1u, NT);
    constexpr int TMEM_COLS = 2 * MT * NT;                // two accumulator groups
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two <= 512");
    constexpr int CK = TF32 ? 4 : 8;

    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t a_bytes = (uint32_t)p.nch * p.rows_max * 16;
    uint8_t* a_smem = smem;                                         // [abuf][nch][rows_max][16]
    uint8_t* w_smem = smem + (size_t)p.abuf * a_bytes;              // [NSTAGE][STAGE_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + NSTAGE * STAGE_BYTES);
    // barrier map: w_full[4] w_empty[4] a_full[2] a_empty[2] acc_full[2] acc_empty[2]
    const uint32_t bar0 = tc::smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    auto A_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
    auto A_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 + s); };
    auto ACC_FULL = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 4 + s); };
    auto ACC_EMPTY = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 6 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { tc::mbar_init(W_FULL(s), 1); tc::mbar_init(W_EMPTY(s), 1); }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(A_FULL(s), 1); tc::mbar_init(A_EMPTY(s), 1);
            tc::mbar_init(ACC_FULL(s), 1); tc::mbar_init(ACC_EMPTY(s), 4);
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tc::smem_u32(tmem_slot), TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int spt = (p.nch + SCH - 1) / SCH;   // weight stages per tap

    auto decode = [&](int tile, int& j, int& b, int& lt, int& nt) {
        j = tile / p.tiles_per_job;
        int r = tile - j * p.tiles_per_job;
        nt = r % p.n_nt; r /= p.n_nt;
        lt = r % p.n_lt;
        b = r / p.n_lt;
    };

    if (warp == 0) {
        // ======================= producer: A tiles + weight stages via bulk async copies =======================
        uint32_t wst = 0, wph = 0;
        int it = 0;
        // stage the activation tile of local iteration `i` (tile id `t`) into buffer i % abuf
        auto issue_a = [&](int i, int t) {
            int j, b, lt, nt;
            decode(t, j, b, lt, nt);
            const rvc_tc_job& jb = p.jobs[j];
            const int pad = (jb.K - 1) / 2 * jb.dil;
            const int rows = TM + 2 * pad;
            const int ab = i % p.abuf;
            const uint32_t aph = (i / p.abuf) & 1;
            if (lane == 0) {
                tc::mbar_wait(A_EMPTY(ab), aph ^ 1);
                tc::mbar_arrive_expect_tx(A_FULL(ab), (uint32_t)p.nch * rows * 16);
            }
            __syncwarp();
            const uint8_t* asrc = reinterpret_cast<const uint8_t*>(jb.a) +
                                  ((size_t)b * p.nch * p.Lp + (size_t)(PADL + lt * TM - pad)) * 16;
            for (int c = lane; c < p.nch; c += 32)
                tc::bulk_g2s(tc::smem_u32(a_smem + (size_t)ab * a_bytes + (size_t)c * p.rows_max * 16), asrc + (size_t)c * p.Lp * 16,
                             (uint32_t)rows * 16, A_FULL(ab));
        };
        if ((int)blockIdx.x < p.total_tiles) issue_a(0, blockIdx.x);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            int j, b, lt, nt;
            decode(tile, j, b, lt, nt);
            const rvc_tc_job& jb = p.jobs[j];
            const int next = tile + gridDim.x;
            bool next_issued = next >= p.total_tiles;
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(jb.w);
            int sidx = 0;
            for (int tap = 0; tap < jb.K; ++tap) {
                for (int s = 0; s < spt; ++s, ++sidx) {
                    // prefetch the next tile's activations once this tile's MMAs are known to be running (double-buffered A only)
                    if (!next_issued && p.abuf == 2 && sidx == NSTAGE) { issue_a(it + 1, next); next_issued = true; }
                    const int c0 = s * SCH;
                    const int nc = min(SCH, p.nch - c0);
                    if (lane == 0) {
                        tc::mbar_wait(W_EMPTY(wst), wph ^ 1);
                        tc::mbar_arrive_expect_tx(W_FULL(wst), (uint32_t)nc * NT * 16);
                    }
                    __syncwarp();
                    const uint32_t dst = tc::smem_u32(w_smem + wst * STAGE_BYTES);
                    if (NT == p.C) {
                        if (lane == 0)
                            tc::bulk_g2s(dst, wsrc + ((size_t)(tap * p.nch + c0) * p.C) * 16, (uint32_t)nc * NT * 16, W_FULL(wst));
                    } else {
                        for (int c = lane; c < nc; c += 32)
                            tc::bulk_g2s(dst + c * NT * 16, wsrc + ((size_t)(tap * p.nch + c0 + c) * p.C + nt * NT) * 16, NT * 16, W_FULL(wst));
                    }
                    if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
                }
            }
            if (!next_issued) issue_a(it + 1, next);
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (single thread) =======================
        if (lane == 0) {
            uint32_t wst = 0, wph = 0;
            int it = 0;
            const uint32_t lbo_a = (uint32_t)p.rows_max * 16, lbo_b = NT * 16;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                int j, b, lt, nt;
                decode(tile, j, b, lt, nt);
                const rvc_tc_job& jb = p.jobs[j];
                const int ab = it % p.abuf;
                const uint32_t aph = (it / p.abuf) & 1;
                const int g = it & 1;
                const uint32_t gph = (it >> 1) & 1;
                tc::mbar_wait(ACC_EMPTY(g), gph ^ 1);
                tc::mbar_wait(A_FULL(ab), aph);
                tc::fence_after_sync();
                const uint32_t a_base = tc::smem_u32(a_smem + (size_t)ab * a_bytes);
                for (int tap = 0; tap < jb.K; ++tap) {
                    for (int s = 0; s < spt; ++s) {
                        const int c0 = s * SCH;
                        const int nc = min(SCH, p.nch - c0);
                        tc::mbar_wait(W_FULL(wst), wph);
                        tc::fence_after_sync();
                        const uint32_t w_base = tc::smem_u32(w_smem + wst * STAGE_BYTES);
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            const uint32_t d_tmem = tmem_base + (uint32_t)((g * MT + m) * NT);
                            for (int kk = 0; kk < nc; kk += 2) {
                                const uint32_t a_addr = a_base + (uint32_t)((c0 + kk) * p.rows_max + m * BM + tap * jb.dil) * 16;
                                const uint32_t b_addr = w_base + (uint32_t)(kk * NT) * 16;
                                tc::umma<TF32>(d_tmem, tc::make_desc(a_addr, lbo_a, 128), tc::make_desc(b_addr, lbo_b, 128), IDESC,
                                               (tap | s | kk) != 0);
                            }
                        }
                        tc::umma_commit(W_EMPTY(wst));
                        if (++wst == NSTAGE) { wst = 0; wph ^= 1; }
                    }
                }
                tc::umma_commit(A_EMPTY(ab));
                tc::umma_commit(ACC_FULL(g));
            }
        }
    } else {
        // ======================= epilogue: 4 warps, warp%4 selects the TMEM lane quarter =======================
        const int q = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            int j, b, lt, nt;
            decode(tile, j, b, lt, nt);
            const rvc_tc_job& jb = p.jobs[j];
            const int g = it & 1;
            const uint32_t gph = (it >> 1) & 1;
            tc::mbar_wait(ACC_FULL(g), gph);
            tc::fence_after_sync();
#pragma unroll 1
            for (int m = 0; m < MT; ++m) {
                const int l = lt * TM + m * BM + q * 32 + lane;
                const bool ok = l < p.L;
#pragma unroll 1
                for (int cb = 0; cb < NT / 32; ++cb) {
                    uint32_t r[32];
                    tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g * MT + m) * NT + cb * 32), r);
                    tc::tmem_ld_wait();
                    const int co0 = nt * NT + cb * 32;
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + __ldg(jb.bias + co0 + i);
                    if (ok) {
                        const size_t nbase = ((size_t)b * p.C + co0) * p.L + l;
                        if (jb.res) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] += jb.res[nbase + (size_t)i * p.L];
                        }
                        if (jb.y_f32) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) jb.y_f32[nbase + (size_t)i * p.L] = v[i];
                        }
                        if (jb.acc_out) {
                            if (jb.beta != 0.f) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    float* o = jb.acc_out + nbase + (size_t)i * p.L;
                                    *o = jb.beta * (*o) + jb.out_scale * v[i];
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i) jb.acc_out[nbase + (size_t)i * p.L] = jb.out_scale * v[i];
                            }
                        }
                        if (jb.y_act) {
                            uint8_t* yb = reinterpret_cast<uint8_t*>(jb.y_act) +
                                          (((size_t)b * p.nch + co0 / CK) * p.Lp + PADL + l) * 16;
#pragma unroll
                            for (int u = 0; u < 32 / CK; ++u) {
                                uint4 pk;
                                if constexpr (TF32) {
                                    pk.x = __float_as_uint(tc::round_tf32(rvc::lrelu(v[4 * u + 0], jb.act_slope)));
                                    pk.y = __float_as_uint(tc::round_tf32(rvc::lrelu(v[4 * u + 1], jb.act_slope)));
                                    pk.z = __float_as_uint(tc::round_tf32(rvc::lrelu(v[4 * u + 2], jb.act_slope)));
                                    pk.w = __float_as_uint(tc::round_tf32(rvc::lrelu(v[4 * u + 3], jb.act_slope)));
                                } else {
                                    pk.x = tc::pack_bf16x2(rvc::lrelu(v[8 * u + 0], jb.act_slope), rvc::lrelu(v[8 * u + 1], jb.act_slope));
                                    pk.y = tc::pack_bf16x2(rvc::lrelu(v[8 * u + 2], jb.act_slope), rvc::lrelu(v[8 * u + 3], jb.act_slope));
                                    pk.z = tc::pack_bf16x2(rvc::lrelu(v[8 * u + 4], jb.act_slope), rvc::lrelu(v[8 * u + 5], jb.act_slope));
                                    pk.w = tc::pack_bf16x2(rvc::lrelu(v[8 * u + 6], jb.act_slope), rvc::lrelu(v[8 * u + 7], jb.act_slope));
                                }
                                *reinterpret_cast<uint4*>(yb + (size_t)u * p.Lp * 16) = pk;
                            }
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
template <bool TF32>
__global__ void __launch_bounds__(256) pack_act_kernel(const float* __restrict__ x, int64_t x_bs, int64_t x_cs, uint8_t* __restrict__ y,
                                                       int nch, int L, int Lp, float slope) {
    constexpr int CK = TF32 ? 4 : 8;
    const int l = blockIdx.x * 256 + threadIdx.x;
    const int c = blockIdx.y, b = blockIdx.z;
    if (l >= L) return;
    const float* xp = x + (int64_t)b * x_bs + (int64_t)c * CK * x_cs + l;
    float v[CK];
#pragma unroll
    for (int i = 0; i < CK; ++i) v[i] = rvc::lrelu(xp[(int64_t)i * x_cs], slope);
    uint4 pk;
    if constexpr (TF32) {
        pk.x = __float_as_uint(tc::round_tf32(v[0])); pk.y = __float_as_uint(tc::round_tf32(v[1]));
        pk.z = __float_as_uint(tc::round_tf32(v[2])); pk.w = __float_as_uint(tc::round_tf32(v[3]));
    } else {
        pk.x = tc::pack_bf16x2(v[0], v[1]); pk.y = tc::pack_bf16x2(v[2], v[3]);
        pk.z = tc::pack_bf16x2(v[4], v[5]); pk.w = tc::pack_bf16x2(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(y + (((size_t)b * nch + c) * Lp + PADL + l) * 16) = pk;
}

// folded fp32 [Cout][Cin][K] -> packed [K][Cin/CK][Cout][CK]
template <bool TF32>
__global__ void pack_weight_kernel(const float* __restrict__ w, uint8_t* __restrict__ y, int Cout, int Cin, int K) {
    constexpr int CK = TF32 ? 4 : 8;
    const int nch = Cin / CK;
    const int64_t n = (int64_t)K * nch * Cout;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int co = (int)(idx % Cout);
    const int c = (int)((idx / Cout) % nch);
    const int tap = (int)(idx / ((int64_t)Cout * nch));
    float v[CK];
#pragma unroll
    for (int i = 0; i < CK; ++i) v[i] = w[((int64_t)co * Cin + c * CK + i) * K + tap];
    uint4 pk;
    if constexpr (TF32) {
        pk.x = __float_as_uint(tc::round_tf32(v[0])); pk.y = __float_as_uint(tc::round_tf32(v[1]));
        pk.z = __float_as_uint(tc::round_tf32(v[2])); pk.w = __float_as_uint(tc::round_tf32(v[3]));
    } else {
        pk.x = tc::pack_bf16x2(v[0], v[1]); pk.y = tc::pack_bf16x2(v[2], v[3]);
        pk.z = tc::pack_bf16x2(v[4], v[5]); pk.w = tc::pack_bf16x2(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(y + idx * 16) = pk;
}

template <bool TF32, int NT>
int launch_tc(const Params& p, size_t smem, int grid, cudaStream_t stream) {
    auto kern = conv1d_tc_kernel<TF32, NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return rvc::fail((int)e, "rvc_conv1d_tc: smem opt-in (%zu B): %s", smem, cudaGetErrorString(e));
    kern<<<grid, NTHREADS, smem, stream>>>(p);
    return rvc::check_launch("rvc_conv1d_tc");
}

}  // namespace

RVC_API int64_t rvc_packed_rows(int32_t L) { return packed_rows(L); }

RVC_API int rvc_pack_act(const float* x, int64_t x_bs, int64_t x_cs, void* y, int32_t B, int32_t C, int32_t L, int32_t dtype, float slope,
                         void* stream) {
    if (!x || !y) return rvc::fail(RVC_EINVAL, "rvc_pack_act: null pointer");
    const int CK = dtype == RVC_DT_TF32 ? 4 : 8;
    if (B <= 0 || B > 65535 || C <= 0 || C % CK || L < 0) return rvc::fail(RVC_ESHAPE, "rvc_pack_act: bad shape (C=%d must be a multiple of %d)", C, CK);
    if (L == 0) return 0;
    dim3 grid(rvc::cdiv(L, 256), C / CK, B);
    if (dtype == RVC_DT_TF32)
        pack_act_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_bs, x_cs, (uint8_t*)y, C / CK, L, packed_rows(L), slope);
    else
        pack_act_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_bs, x_cs, (uint8_t*)y, C / CK, L, packed_rows(L), slope);
    return rvc::check_launch("rvc_pack_act");
}

RVC_API int rvc_pack_weight(const float* w, void* y, int32_t Cout, int32_t Cin, int32_t K, int32_t dtype, void* stream) {
    if (!w || !y) return rvc::fail(RVC_EINVAL, "rvc_pack_weight: null pointer");
    const int CK = dtype == RVC_DT_TF32 ? 4 : 8;
    if (Cout <= 0 || Cin <= 0 || Cin % CK || K <= 0) return rvc::fail(RVC_ESHAPE, "rvc_pack_weight: bad shape");
    const int64_t n = (int64_t)K * (Cin / CK) * Cout;
    if (dtype == RVC_DT_TF32)
        pack_weight_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (uint8_t*)y, Cout, Cin, K);
    else
        pack_weight_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (uint8_t*)y, Cout, Cin, K);
    return rvc::check_launch("rvc_pack_weight");
}

RVC_API int rvc_conv1d_tc(const rvc_conv1d_tc_desc* dp, void* stream) {
    if (!dp) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: null descriptor");
    const rvc_conv1d_tc_desc& d = *dp;
    if (d.njobs < 1 || d.njobs > MAXJOBS) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: njobs must be 1..3");
    const bool tf32 = d.dtype == RVC_DT_TF32;
    const int CK = tf32 ? 4 : 8;
    if (d.B <= 0 || d.L <= 0 || d.C <= 0 || d.C % (2 * CK)) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: C=%d must be a multiple of %d", d.C, 2 * CK);
    const int NT = d.C >= 128 ? 128 : d.C;
    if (NT != 128 && NT != 64 && NT != 32) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: C=%d unsupported (32, 64 or a multiple of 128)", d.C);
    if (d.C % NT) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: C=%d must be a multiple of 128", d.C);
    Params p;
    int maxpad = 0;
    for (int j = 0; j < d.njobs; ++j) {
        const rvc_tc_job& jb = d.jobs[j];
        if (!jb.a || !jb.w || !jb.bias) return rvc::fail(RVC_EINVAL, "rvc_conv1d_tc: job %d has null a/w/bias", j);
        if (jb.K <= 0 || !(jb.K & 1) || jb.dil <= 0) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: job %d needs odd K, dil > 0", j);
        const int pad = (jb.K - 1) / 2 * jb.dil;
        if (pad > PADL) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: halo %d exceeds the %d zero rows of the packed layout", pad, PADL);
        maxpad = pad > maxpad ? pad : maxpad;
        p.jobs[j] = jb;
    }
    p.njobs = d.njobs; p.B = d.B; p.C = d.C; p.L = d.L; p.Lp = packed_rows(d.L);
    p.nch = d.C / CK;
    p.rows_max = TM + 2 * maxpad;
    p.n_lt = rvc::cdiv(d.L, TM);
    p.n_nt = d.C / NT;
    p.tiles_per_job = d.B * p.n_lt * p.n_nt;
    p.total_tiles = p.tiles_per_job * d.njobs;
    const size_t a_bytes = (size_t)p.nch * p.rows_max * 16;
    const size_t fixed = (size_t)NSTAGE * STAGE_BYTES + 256;
    const size_t limit = 227 * 1024;
    p.abuf = (2 * a_bytes + fixed <= limit) ? 2 : 1;
    const size_t smem = p.abuf * a_bytes + fixed;
    if (smem > limit) return rvc::fail(RVC_ESHAPE, "rvc_conv1d_tc: activation tile (%zu B) does not fit shared memory", a_bytes);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    cudaStream_t st = (cudaStream_t)stream;
    if (tf32) {
        if (NT == 128) return launch_tc<true, 128>(p, smem, grid, st);
        if (NT == 64) return launch_tc<true, 64>(p, smem, grid, st);
        return launch_tc<true, 32>(p, smem, grid, st);
    }
    if (NT == 128) return launch_tc<false, 128>(p, smem, grid, st);
    if (NT == 64) return launch_tc<false, 64>(p, smem, grid, st);
    return launch_tc<false, 32>(p, smem, grid, st);
}
