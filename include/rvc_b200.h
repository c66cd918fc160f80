/*
 * rvc_b200.h — C ABI of librvc_b200.so: sm_100a kernels for the RVC-MAKER `Synthesizer.infer` hot path.
 *
 * The reference (TheNeodev/RVC-MAKER) has no FFI of its own: the path is a Python class whose arithmetic is
 * executed by torch.nn.functional calls.  Each entry point below therefore replaces the torch call sites of one
 * reference function (cited as file:line under main/library/algorithm/).  All pointers are DEVICE pointers unless
 * marked host; `stream` is a cudaStream_t passed as void*.  No entry point allocates, synchronises or keeps mutable
 * global state; every function returns 0 on success, a negative RVC_E* code for argument errors or a positive
 * cudaError_t, and records a message retrievable with rvc_last_error() (thread-local).
 *
 * Activation layout: fp32 `[B, C, L]` with explicit batch/channel strides in ELEMENTS (L stride is 1) — the reference's
 * own NCL layout, so channel slices (the flow's x0/x1 halves) are addressed without copies.
 */
#ifndef RVC_B200_H
#define RVC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVC_ABI_VERSION 30

#define RVC_EINVAL (-1)   /* bad argument */
#define RVC_ESHAPE (-2)   /* shape not supported by this kernel */
#define RVC_EALIGN (-3)   /* pointer / stride alignment */

int rvc_abi_version(void);
const char* rvc_last_error(void);
/* fills {sm_count, cc_major, cc_minor, max_smem_optin}; returns 0 or cudaError_t */
int rvc_device_info(int device, int* out4);
/* Cap on the CTAs of the persistent tcgen05 launches (rvc_conv1d_tc, rvc_resblock_pair_tc); 0 = one per SM (default).  Process-wide; returns the
 * previous value.  Lets a caller that runs two independent chunks on two streams leave SMs to the short-sequence kernels of the other chunk. */
int rvc_set_grid_limit(int n);

/* post-activation codes */
#define RVC_ACT_NONE 0
#define RVC_ACT_RELU 1
#define RVC_ACT_TANH 2

/*
 * Generic stride-1 Conv1d, fp32 CUDA-core path, with the fusions the path needs.
 *   in(b,ci,l)  = pre_lrelu ? leaky_relu(x, pre_slope) : x ;  in *= mask[b,l] if in_mask ; zero outside [0,L)
 *   acc(b,co,l) = sum_{ci,k} w[ci][k][co] * in(b, ci, l - pad + k*dil)
 *   gated == 0:  v = act(acc + bias[co] + bias_b[b*bias_b_bs + co])
 *   gated == 1:  v = tanh(acc_t + bias[co] + bias_b[..co]) * sigmoid(acc_s + bias[co+Cout] + bias_b[..co+Cout])
 *                (acc_s uses weight columns co+Cout; weight tensor has 2*Cout columns)   [commons.py:35-41]
 *   v = res ? res(b,co,l) + alpha*v : alpha*v ;  v *= out_scale ;  v *= mask[b,l] if out_mask
 *   y(b,co,l) = beta != 0 ? beta*y(b,co,l) + v : v
 * Weights are pre-packed as [Cin][K][Cw] fp32 (Cw = Cout, or 2*Cout when gated), weight-norm already folded.
 * Replaces: F.conv1d call sites of TextEncoder/Encoder/FFN/MultiHeadAttention 1x1 (synthesizers.py:222,302-305,369),
 * WaveNet (modules.py:43-45), ResidualCouplingLayer pre/post (residuals.py:129), GeneratorNSF.conv_pre/cond
 * (synthesizers.py:146-147), ResBlock (residuals.py:32-36).
 */
typedef struct {
    const float* x;  int64_t x_bs, x_cs;
    const float* w;
    const float* bias;                     /* [Cw] or NULL */
    const float* bias_b; int64_t bias_b_bs;/* per-batch bias [B][>=Cw] or NULL */
    const float* mask;                     /* [B][L] (1.0 / 0.0) or NULL */
    const float* res; int64_t res_bs, res_cs; /* or NULL */
    float* y;        int64_t y_bs, y_cs;
    int32_t B, Cin, Cout, L, K, dil, pad;
    int32_t pre_lrelu; float pre_slope;
    int32_t in_mask, out_mask;
    int32_t act, gated;
    float alpha, out_scale, beta;
} rvc_conv1d_desc;
int rvc_conv1d_f32(const rvc_conv1d_desc* d, void* stream);

/*
 * ConvTranspose1d (stride u, kernel k, padding (k-u)/2, L_out = L_in*u) with fused pre-leaky_relu and the NSF
 * noise-conv injection:  y = convT(lrelu(x, slope)) + bias + [ Conv1d_{1->Cout}(har, kn, stride sn, pad pn) + nbias ].
 * w is the folded ConvTranspose1d weight re-packed [Cin][k][Cout]; nw is [Cout][kn] (NULL for the non-f0 Generator).
 * Replaces GeneratorNSF.forward's `ups(x) + noise_convs(har_source)` (synthesizers.py:150,156) and Generator (:44).
 */
typedef struct {
    const float* x;  int64_t x_bs, x_cs;
    const float* w;  const float* bias;
    const float* har; int64_t har_bs;      /* [B][L_in*u*sn] or NULL */
    const float* nw; const float* nbias;
    float* y;        int64_t y_bs, y_cs;
    int32_t B, Cin, Cout, Lin, k, u, kn, sn, pn;
    float pre_slope;
    /* optional: input = (x + x2 + x3) / in_div — the `xs / num_kernels` of the previous stage (synthesizers.py:157-159)
     * fused at load; x2/x3 share x's strides. */
    const float* x2; const float* x3; float in_div;
} rvc_convtr1d_desc;
int rvc_convtr1d_f32(const rvc_convtr1d_desc* d, void* stream);

/* NSF noise conv alone (synthesizers.py:134-136, :156): y[b,co,l] = bias[co] + sum_j w[co][j] * har[b, l*sn - pn + j]; har has
 * Lout*sn samples per batch row (har_bs apart), y is contiguous [B,C,Lout] with batch stride y_bs.  Used as the residual input
 * of the tensor-core ConvTranspose1d.  blocked = 1 writes y channel-blocked, [b][c/8][l][8] (C % 8 == 0), see rvc_conv1d_tc_desc. */
int rvc_noise_conv_f32(const float* har, int64_t har_bs, const float* w, const float* bias, float* y, int64_t y_bs, int32_t B,
                       int32_t C, int32_t Lout, int32_t kn, int32_t sn, int32_t pn, int32_t blocked, void* stream);

/* y[b,0,l] = tanh( sum_{ci,k} w[ci*7+k] * leaky_relu(in[b,ci,l-3+k], slope) ), no bias — conv_post (synthesizers.py:161);
 * in = x, or (x + x2 + x3) / in_div when x2/x3 are given. */
int rvc_conv_post_f32(const float* x, const float* x2, const float* x3, float in_div, int64_t x_bs, int64_t x_cs, const float* w,
                      float* y, int64_t y_bs, int32_t B, int32_t C, int32_t L, float slope, void* stream);

/*
 * TextEncoder front end (synthesizers.py:367-368): x[b,c,t] = leaky_relu((phone[b,t,:].W[c,:] + bias[c] + E[pitch[b,t],c])
 * * sqrt(H), 0.1) * mask ;  mask[b,t] = t < lengths[b].  wT is W transposed to [D][H]; pitch may be NULL (non-f0).
 * pitch indices are int64 and gathered exactly.
 */
int rvc_embed_f32(const float* phone, const int64_t* pitch, const int64_t* lengths, const float* wT, const float* bias,
                  const float* emb_pitch, float* x, float* mask, int32_t B, int32_t T, int32_t D, int32_t H, void* stream);

/* Tensor-core variant of the front end: rvc_embed_prep writes mask[b,t] = t < lengths[b] and, when pitch != NULL, the exact
 * gather pb[b,c,t] = emb_pitch[pitch[b,t]][c] * scale (fp32 [B,H,T]); rvc_pack_rows converts the channels-last phone features
 * [B,T,D] into the packed operand layout; the Linear itself then runs through rvc_conv1d_tc with res = pb, alpha = sqrt(H),
 * relu = 2 (leaky_relu on the final value) and the mask. */
int rvc_embed_prep(const int64_t* pitch, const int64_t* lengths, const float* emb_pitch, float scale, float* pb, float* mask,
                   int32_t B, int32_t T, int32_t H, void* stream);
int rvc_pack_rows(const float* x, void* y, int32_t B, int32_t T, int32_t D, int32_t dtype, void* stream);

/* y = LayerNorm_C(x + r [+ r2 + r3]) * gamma + beta over the channel axis of [B,C,T] (synthesizers.py:179-181, :346-347);
 * r2 / r3 (both or neither) are further partial sums of the residual branch — the FFN's second conv runs as three
 * input-channel slices in one launch when the sequence is short.  out_mask multiplies by mask[b,t] (the final `x * x_mask`,
 * :348).  y_act (optional): the MASKED result in the packed bf16/fp16 layout of the tensor-core convs (see below),
 * act_dtype = RVC_DT_BF16 / RVC_DT_FP16. */
int rvc_layernorm_f32(const float* x, const float* r, const float* r2, const float* r3, const float* gamma, const float* beta,
                      const float* mask, float* y, int32_t B, int32_t C, int32_t T, float eps, int32_t out_mask, void* y_act,
                      int32_t act_dtype, void* stream);

/*
 * Relative-position multi-head attention core (synthesizers.py:221-250), heads share the [2w+1, dk] tables:
 *   s[i,j] = (q_i/sqrt(dk)).k_j + [|j-i|<=w] (q_i/sqrt(dk)).Ek[j-i+w];  s = -1e4 where mask_i*mask_j == 0
 *   p = softmax_j(s);  out_i = sum_j p[i,j] (v_j + [|j-i|<=w] Ev[j-i+w])
 * qkv is [B, 3*H, T] (q rows 0..H-1, k rows H..2H-1, v rows 2H..3H-1), out is [B,H,T]; head h owns channels h*dk..
 */
int rvc_rel_attention_f32(const float* qkv, const float* mask, const float* emb_k, const float* emb_v, float* out,
                          int32_t B, int32_t H, int32_t n_heads, int32_t T, int32_t window, void* stream);

/* Speaker conditioning on a length-1 sequence (modules.py:39 cond_layer(g), synthesizers.py:147 cond(g)):
 * y[b][co] = bias[co] + sum_ci w[ci][co] * x[b][ci], w packed [Cin][Cout]. */
int rvc_cond_gemv_f32(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Cout, void* stream);

/* z_p = (m + exp(logs) * eps * 0.66666) * mask, stats = [B, 2*C, T] (m rows then logs rows) — synthesizers.py:449. */
int rvc_sample_zp_f32(const float* stats, const float* eps, const float* mask, float* zp, int32_t B, int32_t C, int32_t T,
                      void* stream);

/*
 * SineGen + SourceModuleHnNSF, harmonic_num = 0 (synthesizers.py:79-100, :111-112):
 *   har[b, t*upp+n-1] = tanh(lw * s + lb),  s = 0.1*sin(2*pi*(f0_t/sr*n + phase_t))*uv + (uv*0.003 + (1-uv)*0.1/3)*noise
 *   phase_t = fmod( sum_{t'<t} (fmod(f0_t'/sr*upp + 0.5, 1) - 0.5), 1 ), prefix sum in fp64 rounded to fp32 per frame.
 * scratch: B*T floats.
 */
int rvc_sine_source_f32(const float* f0, const float* noise, const float* lin_w, const float* lin_b, float* scratch,
                        float* har, int32_t B, int32_t T, int32_t upp, float sr, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05) Conv1d: Cin -> Cout, odd K, dilation, "same" padding, stride 1, fp32 accumulation in TMEM.
 * Used for the ResBlocks (residuals.py:32-36), the TextEncoder 1x1 / FFN convs (synthesizers.py:222,302-305,369), the flow's
 * pre / WaveNet / post convs (residuals.py:129, modules.py:43-45) and conv_pre (synthesizers.py:146).
 *
 * Packed activation layout: act[b][c / 8][rvc_packed_rows(L)][8] in bf16 or fp16, i.e. 16-byte units of 8 channels per
 * position; position l lives at row 64 + l and every other row is ZERO (allocate zero-filled; kernels never write pad
 * rows).  Packed weights: [K][Cin / 8][Cout][8].
 */
#define RVC_DT_BF16 0
#define RVC_DT_FP16 2
/* output format of rvc_pack_rows / rvc_layernorm_f32 only: TWO fp16 planes per batch item, [b][2*C/8][rows][8] — chunks [0, C/8) hold
 * hi = fp16(x), chunks [C/8, 2*C/8) hold lo = fp16(x - hi); the operands of the split-precision GEMMs (rvc_conv1d_tc_desc.sum_jobs) */
#define RVC_DT_FP16_SPLIT 3
int64_t rvc_packed_rows(int32_t L);
/* y_packed = DT(leaky_relu(in, slope) * mask) for fp32 x[B,C,L] with element strides (slope = 1 -> identity, mask may be NULL);
 * in = x, or (x + x2 + x3) / in_div when x2/x3 are given (the `xs / num_kernels` of a generator stage). */
int rvc_pack_act(const float* x, const float* x2, const float* x3, float in_div, int64_t x_bs, int64_t x_cs, const float* mask, void* y,
                 int32_t B, int32_t C, int32_t L, int32_t dtype, float slope, void* stream);
/* folded fp32 [Cout][Cin][K] -> packed [K][Cin/8][Cout][8] */
int rvc_pack_weight(const float* w, void* y, int32_t Cout, int32_t Cin, int32_t K, int32_t dtype, void* stream);

/*
 * One conv job.  With acc = conv(a) (weight-column order):
 *   gated == 0:  v[co] = acc[co] + bias[co] + bias_b[b][co]
 *   gated == 1:  v[c]  = tanh(acc[t] + bias[t] + bias_b[b][t]) * sigmoid(acc[s] + bias[s] + bias_b[b][s]), where within every
 *                N tile of `nt` weight columns the first nt/2 are the tanh columns of output channels [tile*nt/2, ...) and
 *                the last nt/2 the matching sigmoid columns (the caller packs weights / biases in that order)  [commons.py:35-41]
 *   relu == 1: v = max(v, 0);   res: v = res[b][c][l] + alpha * v;   relu == 2: v = leaky_relu(v, act_slope) (then y_act adds none);
 *   mask: v *= mask[b][l]
 *   y_f32[b][c][l] = v (element strides y_bs / y_cs);   y_act = DT(leaky_relu(v, act_slope)) in the packed layout.
 * Up to six jobs sharing (B, Cin, Cout, L, dtype) run in one persistent launch.  nt / mt = 0 let the library pick the tile.
 */
typedef struct {
    const void* a; const void* w;
    const float* bias;
    const float* bias_b; int64_t bias_b_bs;
    const float* mask;
    const float* res; int64_t res_bs, res_cs;
    float* y_f32; int64_t y_bs, y_cs;
    void* y_act;
    int32_t K, dil, pad, gated, relu;   /* pad = zero rows left of position 0 seen by tap 0 ("same": (K-1)/2*dil) */
    float alpha, act_slope;             /* without res: v *= alpha */
    /* optional layout overrides (0 = default), used when one operand is itself an activation (attention: S = q.k^T, O = P.v):
     * a_bs / y_act_bs = bytes between batch items of a / y_act; w_bs = bytes between per-batch weight sets (0: shared
     * weights); w_cs = rows between consecutive 16-byte weight chunks (default Cout). */
    int64_t a_bs, w_bs, y_act_bs;
    int32_t w_cs, reserved;
    /* optional, in place of `res`: the NSF noise conv of a generator stage (synthesizers.py:134-136, :156) computed in the
     * epilogue — v += nz_b[c] + sum_{j < nz_kn} nz_w[c*nz_kn + j] * nz_har[b*nz_har_bs + l*nz_sn - nz_pn + j] at output
     * position l (source samples outside [0, Lout*nz_sn) read as 0), nz_kn <= 4.  Same arithmetic as rvc_noise_conv_f32
     * followed by res: bit-identical, without the extra launch and the fp32 round trip. */
    const float* nz_har; const float* nz_w; const float* nz_b;
    int64_t nz_har_bs;
    int32_t nz_kn, nz_sn, nz_pn, reserved2;
    /* optional second packed output next to y_act (fp16 only): y_lo = fp16(leaky_relu(v, act_slope) - float(y_act)), the low half of
     * the split-fp16 residual stream the fused ResBlock steps read (rvc_pair_job.a_lo) */
    void* y_lo;
    /* optional, in place of `res` (fp16 launches, up <= 1, act_slope > 0): the residual arrives as that split-fp16 stream,
     * res = inv_leaky_relu(float(res_hi) + float(res_lo), act_slope); both planes in the packed layout and batch stride of y_act (they are the
     * y_act / y_lo an earlier launch of the same shape wrote).  y_act may alias res_hi and y_lo may alias res_lo: a plane is read only at a
     * thread's own output elements, before it writes them.  The wide generator stages keep their residual stream this way
     * (residuals.py:32-36): 10 instead of 12 bytes per element and conv2. */
    const void* res_hi; const void* res_lo;
} rvc_tc_job;
/*
 * up > 1 selects the ConvTranspose1d phase mode (GeneratorNSF.ups, synthesizers.py:133,156): the packed weights hold `up`
 * phase blocks [up][K][Cin/8][Cout][8] with K = ceil(k/up) taps, block r tap t = w_T[:, :, r + (K-1-t)*up] (zero past k), every
 * job uses pad = K-1, and the result of input position q, phase r lands at output position q*up + r - pshift
 * (pshift = (k-up)/2); res / mask / y_f32 / y_act are indexed at OUTPUT positions (length L*up).
 *
 * sum_jobs = 1 (stage average of the ResBlocks, synthesizers.py:157-159): every tile runs ALL jobs into one accumulator and
 * the epilogue of job 0 receives v = alpha * (sum_j conv_j + sum_j bias_j + sum_j res_j); outputs / mask / act_slope of job 0.
 * sum_jobs = g >= 2: the same per GROUP of g consecutive jobs (njobs % g == 0; up to 6 jobs): group leader = job k*g holds the
 * outputs and epilogue options, members contribute conv + bias (+ res when they have one).  Used for the split-precision GEMMs
 * of the TextEncoder: a.W ~= a_hi.W_hi + a_lo.W_hi + a_hi.W_lo with fp16 planes hi = fp16(x), lo = fp16(x - hi) of both operands
 * (three jobs, one accumulator): ~2^-21 operand precision on the fp16 tensor-core path.
 *
 * f32_blocked (bit 0: res, bit 1: y_f32 of every job): the fp32 tensor is channel-blocked, [b][c/8][l][8] — element (b, c, l)
 * at b*bs + (c/8)*cs + l*8 + c%8 with cs = the stride between 8-channel blocks (8*L when dense).  The epilogue then moves
 * 32 contiguous bytes per thread and 1 KB per warp instead of 32 scattered 128-byte lines; the generator's residual streams
 * use it (measured: 5.5 vs 4.4 TB/s for the conv2 epilogue's traffic, tools/epilogue_pattern_bench.cu).
 */
typedef struct {
    rvc_tc_job jobs[6];
    int32_t njobs, B, Cin, Cout, L, dtype, nt, mt, up, pshift, sum_jobs, f32_blocked;
    /* up_cols = u > 1 ("phase columns", the stride-2 stages of the generator): a stride-u ConvTranspose1d as ONE plain conv (up <= 1) whose
     * Cout = u * Cr weight columns are the u phases side by side — column block r holds block r of the `up` layout above, K = ceil(k/u) taps,
     * pad = K - 1; bias has Cout entries (the layer's bias once per phase).  The column (r, c) of input position q lands at output position
     * q*u + r - pshift, channel c: outputs (y_f32 / y_act / y_lo / mask / the fused noise conv) have Cr channels and length L*u.  Cr must be a
     * power of two >= 32; no residual inputs.  Against u separate phase tiles: half the operand loads and tile hand-offs, MMAs of twice the width. */
    int32_t up_cols, reserved;
} rvc_conv1d_tc_desc;
int rvc_conv1d_tc(const rvc_conv1d_tc_desc* d, void* stream);

/*
 * A chain of DEPENDENT rvc_conv1d_tc launches as ONE persistent kernel with a grid-wide barrier between the phases (the flow's
 * 32 convs, residuals.py:87-95,127-137 + modules.py:35-51; the FFN pairs of the TextEncoder, synthesizers.py:302-315): at T = 1000
 * such launches are latency-bound (8-9 us each for 1-2 us of tensor work); in a chain the weight stream of phase i+1 overlaps the
 * epilogue of phase i and a phase boundary costs a barrier instead of a kernel boundary.  Phase i = descs[i] with the tile fixed
 * at 128 x 64 (results bit-identical to rvc_conv1d_tc with nt = 64, mt = 1); restrictions: up <= 1, no per-batch weights (w_bs),
 * no fused noise conv, njobs <= 3, Cout % 64 == 0 — a layer with fewer channels pads its packed weights with zero rows and
 * passes the real count in cout_valid[i] (multiple of 32; NULL / 0 = Cout): only those channels are stored.
 *   rvc_conv1d_tc_chain_build: host only; writes a 64-byte header + the phase table into `blob` (rvc_conv1d_tc_chain_blob_bytes(n)
 *     bytes).  The caller copies the blob to 16-byte aligned DEVICE memory once and keeps the host copy of the header.
 *   rvc_conv1d_tc_chain_run: one launch.  dev_sync: 4 zero-initialised uint32 in device memory owned by this chain (barrier counter,
 *     exit counter, error flag [2] = 1 if a barrier timed out, spare); the kernel leaves the counters at zero.
 */
int64_t rvc_conv1d_tc_chain_blob_bytes(int32_t n);
int rvc_conv1d_tc_chain_build(const rvc_conv1d_tc_desc* descs, const int32_t* cout_valid, int32_t n, void* blob, int64_t blob_cap);
int rvc_conv1d_tc_chain_run(const void* host_header, const void* dev_blob, uint32_t* dev_sync, void* stream);

/*
 * The reverse ResidualCouplingBlock in ONE kernel (residuals.py:87-95, :127-137; WaveNet modules.py:35-51) for the standard RVC flow:
 * 4 couplings of 2 x 96 channels, WaveNet of 3 layers, hidden 192, kernel 5.  A cluster of six CTAs carries a tile of 128 frames (80 kept,
 * 24 frames of halo per side recomputed) through all 32 convs: operands are exchanged through distributed shared memory, the hidden stream
 * and the skip sum are accumulated in TMEM, weights stream from L2; global memory is touched for z_p (read once) and z (written once).
 * Per coupling f (index = position in flow.flows / 2; processed 3, 2, 1, 0; the four Flips folded into channel-reversed pre / post weights
 * by the caller): packed fp16 weights as for rvc_conv1d_tc — pre_w [192 <- 96], in_w[l] [384 <- 192, k = 5] with its columns gate-ordered
 * (per 64 columns: 32 tanh then the matching 32 sigmoid), res_w[l] / skip_w[l] [192 <- 192] (l = 2: all of res_skip_layers[2] is skip),
 * post_w [post_rows >= 96 <- 192] (zero rows past 96); fp32 biases: hb[l] = b_pre + sum_{l' < l} b_res[l'] (what the hidden stream lacks
 * when in_layers[l] reads it), in_b[l] gate-ordered, sb = sum_l b_skip[l], post_b; gc = cond_layer(g) [B][...] with row stride gc_bs,
 * gate-ordered, 3 x 384 values per coupling from gc_off.  x0 / x1 = the halves of z_p [B, 192, T] (coupling 3 and 1 read the second half
 * and update the first, 2 and 0 the other way round); z [B, 192, T] receives the result, za (optional) its masked packed fp16 copy.
 */
typedef struct {
    const void* pre_w; const void* in_w[3]; const void* res_w[2]; const void* skip_w[3]; const void* post_w;
    const float* hb[3]; const float* in_b[3]; const float* sb; const float* post_b;
    int32_t gc_off, reserved;
} rvc_flow_coupling;
typedef struct {
    rvc_flow_coupling flows[4];
    const float* zp; float* z; void* za; const float* mask; const float* gc; int64_t gc_bs;
    int32_t B, T, hidden, half, n_flows, n_layers, kernel_size, post_rows;
} rvc_flow_desc;
int rvc_flow_reverse_tc(const rvc_flow_desc* d, void* stream);

/*
 * One ResBlock step  x' = c2(lrelu(c1_d(lrelu(x)) + b1)) + b2 + x  (ResBlockBase.forward, residuals.py:32-36) in ONE kernel for
 * the narrow generator stages (C = 32 or 64): the intermediate stays in shared memory as the second conv's tensor-core operand.
 * Up to three jobs (the three ResBlocks of a stage) per persistent launch.  Per job: a = packed lrelu(x) [B][C/8][rows][8];
 * w1 / w2 packed [K][C/8][C][8] (rvc_pack_weight), odd K1 / K2, conv1 dilation dil1, conv2 dilation 1, "same" padding;
 * res, y: fp32 CHANNEL-BLOCKED [B][C/8][L][8] (see rvc_conv1d_tc_desc.f32_blocked; y may alias res); y_act = packed
 * lrelu(x').  Results are bit-identical to the two rvc_conv1d_tc launches they replace.
 */
typedef struct {
    const void* a;
    const void* w1; const float* b1;
    const void* w2; const float* b2;
    const float* res; float* y; void* y_act;
    int32_t K1, dil1, K2, reserved;
    /* Split-fp16 residual stream (fp16 only; optional, both NULL = the fp32 streams above).  The stream carries y = lrelu(x) as TWO
     * packed fp16 planes, hi = fp16(y) — which IS the conv operand `a` — and lo = fp16(y - hi): 22 significant bits in 4 bytes,
     * against 4 (fp32 x) + 2 (packed operand) bytes.  a_lo != NULL: the residual is x = inv_lrelu(float(a) + float(a_lo)) and `res`
     * is ignored;  y_lo != NULL: the result is written as y_act = fp16(lrelu(x')) and y_lo = fp16(lrelu(x') - float(y_act)) and `y`
     * is ignored (plain mode; in sum mode the average still leaves through y / y_act).  A step then moves 8 bytes per element
     * instead of 12; |x - x_fp32| <= 2^-22 |y| + 3e-8 (the lo plane's subnormal spacing), slope > 0 required. */
    const void* a_lo; void* y_lo;
} rvc_pair_job;
typedef struct {
    rvc_pair_job jobs[3];
    int32_t njobs, B, C, L, dtype;
    float slope;        /* leaky_relu slope of both activations (0.1) */
    /* sum_jobs = 1: the last step of a stage (synthesizers.py:157-159).  Every tile runs all jobs, their conv2 accumulate into one
     * tile and job 0's outputs receive alpha * sum_j (c2_j(...) + b2_j + res_j): y (optional; channel-blocked, or plain [B,C,L]
     * when y_ncl = 1) and y_act (optional; packed lrelu of it). */
    int32_t sum_jobs, y_ncl;
    float alpha;
    int32_t reserved;   /* 0 (timing experiments: 1 = no MMA issue, 2 = no epilogue-2 global I/O, 4 = no residual loads) */
} rvc_resblock_pair_desc;
int rvc_resblock_pair_tc(const rvc_resblock_pair_desc* d, void* stream);

/* Debug aid (tools/pair_trace.py; only in the -DRVC_PAIR_TRACE build, RVC_EINVAL otherwise): registers a device buffer of `capacity` 64-bit slots that
 * CTA 0 of rvc_resblock_pair_tc fills with the SM clock of the hand-offs between its producer / MMA / epilogue warps (slot 80 * unit + 16 * role + k);
 * NULL switches it off. */
int rvc_debug_pair_trace(void* dev_buf, int32_t capacity);

/*
 * Softmax stage of the tensor-core attention (MultiHeadAttention.attention, synthesizers.py:221-250).  The contractions run
 * through rvc_conv1d_tc: S[b][h][j/8][i][j%8] = q_i . k_j (fp32, channel-blocked over the keys — f32_blocked output of the
 * GEMM; Tk >= T key rows per head, Tk % 8 == 0; rows >= T ignored) and
 * O = P . v (+ R).  This call computes, per (b, head, query i):
 *   s[i,j] = S[j][i] / sqrt(dk) + [|j-i| <= w] (q_i / sqrt(dk)) . emb_k[j-i+w];   s = -1e4 where mask_i * mask_j == 0
 *   P = softmax_j(s) -> packed fp16 [b][h][Tk/8][rvc_packed_rows(T)][8] (positions = queries, channels = keys)
 *   R[b][h*dk + d][i] = sum_{|j-i|<=w} P[i,j] emb_v[j-i+w][d]                        (fp32 [B,H,T])
 *   vw = v^T packed over keys: fp16 [b][h][Tk/8][dk][8]                              (weights of the P.v GEMM)
 * qkv is the fp32 [B, 3H, T] projection (q, k, v row blocks).
 */
int rvc_attn_softmax(const float* S, const float* qkv, const float* mask, const float* emb_k, const float* emb_v, void* P, float* R,
                     void* vw, int32_t B, int32_t H, int32_t n_heads, int32_t T, int32_t Tk, int32_t window, void* stream);

/*
 * Scores + softmax of the attention for T <= 1024, head dim 96, as one cluster kernel (MultiHeadAttention.attention, synthesizers.py:221-250): a cluster
 * of ceil(T/256) CTAs per 128 queries of a (batch item, head), one 256-key tile each; S = q.k^T and the banded relative-key term on tensor cores from the
 * packed q / k planes the qkv projection wrote (planes = 2: split precision, the lo plane lo_chunks chunks behind the hi plane), softmax on the accumulator
 * with the row maxima / sums exchanged through distributed shared memory.  Replaces the S launch of rvc_conv1d_tc + rvc_attn_softmax (+ the R buffer):
 *   P  : packed fp16 [b][h][Tk/8 + 4][rvc_packed_rows(T)][8], Tk = 256 * ceil(T/256): chunks < Tk/8 hold softmax(s) over the keys, the last four chunks the
 *        2w+1 band probabilities p[i, i + r - w] as columns r (zero-initialised by the caller once: entries whose key falls outside [0, Tk) are never written);
 *   vw : the weights of the following P.v GEMM, fp16 [b][h][Tk/8 + 4][96][8]: v^T over the keys (written here) and, in the last four chunks, emb_v^T over
 *        the relative positions (written by the caller once) — the relative-value term rides along as four more k-steps of that GEMM (Cin = Tk + 32).
 * qk: packed planes [b][...chunks][rows][8] with q of head h in chunks q_chunk0 + 12 h .., k in k_chunk0 + 12 h ..; ek: emb_rel_k packed as planes
 * [planes][12][32][8] (row = relative position, zero past 2w+1); qkv: the fp32 [B, 3H, T] projection (v rows); window <= 11.
 */
typedef struct {
    const void* qk; int64_t qk_bs; int32_t q_chunk0, k_chunk0, lo_chunks, planes;
    const void* ek; const float* qkv; const float* mask; void* P; void* vw;
    int32_t B, H, n_heads, T, window, reserved;
    /* att != NULL: the P.v GEMM runs in the same kernel (P and vw are then unused and may be NULL) — every CTA multiplies its 256-key slice of P
     * (which never leaves shared memory) with v^T of its keys, the partial sums are reduce-scattered through distributed shared memory and the
     * cluster writes att, the packed fp16 operand of conv_o: [b][chunks][rvc_packed_rows(T)][8] with head h, dim d in chunk (96 h + d) / 8, batch
     * stride att_bs bytes, the lo plane fp16(O - hi) att_lo_chunks chunks behind the hi plane (0: no lo plane); evT = emb_rel_v^T packed
     * [4][96][8] fp16 (relative position major, zero past 2w+1). */
    void* att; int64_t att_bs; const void* evT; int32_t att_lo_chunks, reserved2;
} rvc_attn_scores_desc;
int rvc_attn_scores_cluster_tc(const rvc_attn_scores_desc* d, void* stream);

/*
 * The position-local tail of an encoder layer as one cluster kernel (Encoder.forward, synthesizers.py:344-347; LayerNorm :179-181; FFN :302-315), for
 * hidden 192 / filter 768 / kernel 3:   x1 = LN1(x + conv_o(att));   y = LN2(x1 + conv_2(relu(conv_1(x1 * mask)) * mask) * mask).
 * att: the packed [hi | lo] attention output ([b][2 * 24][rows][8], what the P.v GEMM wrote); x / y: the fp32 residual stream [B, 192, T] (may alias);
 * ya: y * mask as packed [hi | lo] planes for the next layer's qkv projection / proj; wo_hi / wo_lo: conv_o as two packed planes, w1 / w2 packed fp16
 * (rvc_pack_weight), fp32 biases and LayerNorm parameters; out_mask: y itself is masked too (the last layer, :348).  Replaces five launches
 * (rvc_conv1d_tc x 3, rvc_layernorm_f32 x 2) — see csrc/encoder_tail_cluster_tc.cu.
 */
typedef struct {
    const void* att; int64_t att_bs;
    const float* x; float* y; void* ya; int64_t ya_bs; const float* mask;
    const void* wo_hi; const void* wo_lo; const float* bo;
    const void* w1; const float* b1; const void* w2; const float* b2;
    const float* gamma1; const float* beta1; const float* gamma2; const float* beta2;
    int32_t B, T, hidden, filter, kernel_size, out_mask;
    float eps; int32_t reserved;
} rvc_encoder_tail_desc;
int rvc_encoder_tail_tc(const rvc_encoder_tail_desc* d, void* stream);

/*
 * Caller side of the path: the feature preparation VC.voice_conversion performs between the embedder and
 * Synthesizer.infer (main/inference/convert.py:361-378).
 *   out[b,t,:] = feats[b,t/2,:]                                    (feats0 == NULL: F.interpolate(scale_factor=2), :361)
 *   out[b,t,:] = feats[b,t/2,:]*pf + feats0[b,t/2,:]*(1-pf),  pf = protect where pitchf[b,t] < 1, else 1   (:372-378)
 * feats / feats0 are [B,Th,D] fp32 at 50 frames/s, pitchf is [B,>=T] fp32 (Hz) with row stride pitchf_bs, out is [B,T,D] with
 * T <= 2*Th (the p_len clamp of :365-370).  Each product and the sum are rounded separately: bit-identical to the reference's
 * elementwise ops.
 */
int rvc_feature_prep_f32(const float* feats, const float* feats0, const float* pitchf, int64_t pitchf_bs, float protect, float* out,
                         int32_t B, int32_t Th, int32_t T, int32_t D, void* stream);

/* result[b] = max(result[b], max_i |x[b*x_bs + i]|), i < n — the peak the pipeline normalises by (`np.abs(audio_opt).max()`,
 * convert.py:451), accumulated chunk by chunk on the device; the caller zeroes `result` first. */
int rvc_absmax_f32(const float* x, int64_t x_bs, int64_t n, float* result, int32_t B, void* stream);

/*
 * The f0 tail between the RMVPE network and Synthesizer.infer (round 2, third session; SURVEY 8f-2, post-network part).  The reference moves
 * the salience map to the host and does this in numpy (float64; float32 for the salience sums); here it stays on the device.
 *
 * rvc_rmvpe_decode_f32 — RMVPE.decode + to_local_average_cents (main/library/predictors/RMVPE.py:217-221, :237-252):
 *   salience [T][360] fp32 (row stride in elements) -> f0 [T] float64 (Hz, 0 = unvoiced).  Per frame: first arg-max, the 9 bins around it
 *   (zero padding outside 0..359), cents = sum(s*c)/sum(s) with the reference's operation order and widths (products float64, the 9-term
 *   sums in numpy's pairwise order, denominator summed in float32), 0 where max <= thred, f0 = 10*2^(cents/1200), f0 == 10 -> 0.
 *   gate_hi > 0: f0 outside [gate_lo, gate_hi] -> 0 (`infer_from_audio_with_pitch`, :229-233, the "rmvpe-legacy" method).
 * rvc_f0_coarse_f64 — the tail of VC.get_f0 (main/inference/convert.py:308-323) on a float64 f0 track of n frames:
 *   notes != NULL: Autotune.autotune_f0 (:173-179) towards the nearest of n_notes device-resident note frequencies (:200), by `strength`;
 *   f0 *= shift (shift = pow(2, semitones/12) computed by the caller in double, :311); coarse = rint(clip(mel map to 1..255)) (:318-323) with
 *   mel_min / mel_max = 1127*ln(1 + {50,1100}/700) (:198-199).  Writes pitch (int64, `.long()` of :439), pitchf (float32 of the shifted f0,
 *   `.float()` of :439) and, when f0_out != NULL, the shifted float64 f0.  f0 / f0_out may alias.
 * rvc_rmvpe_pitch_f32 — both in one launch (salience -> pitch, pitchf [, f0]).
 * Integer results are bit-exact; exp2 / log are the CUDA double routines (<= 1 ulp).
 */
int rvc_rmvpe_decode_f32(const float* salience, int64_t row_stride, int32_t T, float thred, double gate_lo, double gate_hi, double* f0,
                         void* stream);
int rvc_f0_coarse_f64(const double* f0, int64_t n, const double* notes, int32_t n_notes, double strength, double shift, double mel_min,
                      double mel_max, int64_t* pitch, float* pitchf, double* f0_out, void* stream);
int rvc_rmvpe_pitch_f32(const float* salience, int64_t row_stride, int32_t T, float thred, double gate_lo, double gate_hi,
                        const double* notes, int32_t n_notes, double strength, double shift, double mel_min, double mel_max,
                        int64_t* pitch, float* pitchf, double* f0_out, void* stream);

/*
 * The RMVPE mel front end (pre-network part), one launch: MelSpectrogram.forward for keyshift = 0, speed = 1, center = True
 * (main/library/predictors/RMVPE.py:161-181: torch.stft(n_fft = win = 1024, hop = 160, reflect padding), magnitude, mel_basis matmul,
 * log(clamp(., clamp))) and the reflect padding of the frame axis that mel2hidden applies before the network (:212-213).
 *   audio [n] fp32 (16 kHz), window [1024] (torch.hann_window(1024)), mel_basis [128][513] (the module's registered buffer, :146-148)
 *   out [128][ld] fp32: columns 0..T-1 = log-mel frames (T = 1 + n / 160), columns T..Tp-1 = columns T-2, T-3, ... (Tp = T rounded up to 32 for
 *   the network; Tp = T for the bare spectrogram).  Needs n > 512 and Tp - T < T, as the reference's two reflect paddings do.
 * fp32 like the reference (is_half = False); the summation order differs from cuFFT / cuBLAS, so parity is by tolerance (tests/test_gpu_mel_front.py).
 */
int rvc_mel_spectrogram_f32(const float* audio, int64_t n, const float* window, const float* mel_basis, const int32_t* ranges, float* out,
                            int64_t ld, int32_t T, int32_t Tp, float clamp, void* stream);
/* ranges [128][2] = [lo, hi) of the non-zero bins of every mel_basis row (triangular filters: 2 - 45 of 513 bins), computed once per filterbank;
 * passed to rvc_mel_spectrogram_f32 it restricts the filterbank product to them (NULL: dense product). */
int rvc_mel_basis_ranges(const float* mel_basis, int32_t* ranges, void* stream);

/*
 * The split-point search at the head of VC.pipeline's chunk loop (main/inference/convert.py:404-412): for every candidate t = t_center, 2 t_center, ... < n
 * the first sample in [t - t_query, min(t + t_query, n)) at which |audio_sum| is smallest, audio_sum[p] = sum_{i < window} audio_pad[p + i] over the
 * signal reflect-padded by window/2.  Each sum is formed like numpy forms it (sequentially from 0, one rounding per addition, in the signal's dtype:
 * is_f64 = 1 for the float64 signal behind the reference's filtfilt, 0 for float32), so opt_ts is bit-identical to the reference's.
 *   audio [n] (device), opt_ts [ (n - 1) / t_center ] int64 (device), ws: rvc_split_points_workspace(n, t_center, t_query) bytes (device, 16-byte aligned).
 * The caller applies the reference's guard (`if audio_pad.shape[0] > t_max`, :406) itself.
 */
int64_t rvc_split_points_workspace(int64_t n, int64_t t_center, int64_t t_query);
int rvc_split_points(const void* audio, int32_t is_f64, int64_t n, int32_t window, int64_t t_center, int64_t t_query, int64_t* opt_ts, void* ws,
                     void* stream);

/*
 * The zero-phase high-pass at the head of VC.pipeline, `audio = signal.filtfilt(bh, ah, audio)` (main/inference/convert.py:402; scipy defaults: odd
 * extension by padlen = 3 * ncoef, steady-state initial conditions), in float64 on the device: the direct-form-II-transposed recurrence of scipy's
 * lfilter with scipy's operation order inside blocks of rvc_filtfilt_block(n) samples (4096), the blocks' start states chained through ML = M^block (the
 * state-transition matrix of the recurrence raised to the block length, row-major [order][order]; rvc_maker_b200.pipeline obtains it by running the
 * recurrence).  Not bit-identical to scipy's sample-by-sample evaluation: ~2e-8 absolute on a signal of amplitude 0.5, the rounding-noise level of the
 * filter itself.
 *   x [n] float32 / float64 (device), b, a [ncoef], zi [ncoef - 1] (scipy.signal.lfilter_zi), ML: HOST pointers; y [n] float64 (device);
 *   ws: rvc_filtfilt_workspace(n, ncoef) bytes (device).  Needs a[0] == 1 and n > 3 * ncoef.
 */
int32_t rvc_filtfilt_block(int64_t n);
int64_t rvc_filtfilt_workspace(int64_t n, int32_t ncoef);
int rvc_filtfilt(const void* x, int32_t is_f64, int64_t n, const double* b, const double* a, int32_t ncoef, const double* zi, const double* ML,
                 double* y, void* ws, void* stream);

/*
 * The retrieval blend of VC.voice_conversion behind the FAISS search (main/inference/convert.py:354-359; the search itself, :353, stays the caller's):
 *   weight = square(1 / score); weight /= weight.sum(1); npy = sum_k big_npy[ix[:, k]] * weight[:, k]; out = npy * index_rate + (1 - index_rate) * feats
 * feats, out [T][D] fp32, big_npy [N][D] fp32 (index.reconstruct_n(0, ntotal), :395, resident on the device), score [T][8] fp32, ix [T][8] int64.
 * one_minus_rate = (float)(1 - index_rate) computed by the caller in double, as Python does.  numpy's orders and roundings: bit-identical.
 */
int rvc_retrieval_blend_f32(const float* feats, const float* big_npy, const float* score, const int64_t* ix, float index_rate, float one_minus_rate,
                            float* out, int32_t T, int32_t D, int32_t K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RVC_B200_H */
