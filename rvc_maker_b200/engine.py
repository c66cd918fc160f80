"""Host-side execution engine: folds/packs the weights once, builds a launch plan per (B, T) and replays it.

A *plan* is a flat list of C-ABI calls with pre-built descriptors pointing into persistent device workspaces, so a
replay costs one ctypes call per kernel (and is CUDA-graph capturable).  PyTorch is used only for device memory,
streams, the RNG draws (same generator/order/shapes as the reference: synthesizers.py:449, :87, :98) and copies.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, DT_BF16, DT_FP16, AttnScoresDesc, Conv1dDesc, Conv1dTcDesc, ConvTr1dDesc, EncoderTailDesc, FlowDesc, ResblockPairDesc

LRELU_SLOPE = 0.1
WINDOW = 10


class V:
    """A channel-slice view of a contiguous fp32 [B, Ctot, L] device tensor."""

    __slots__ = ("t", "ptr", "bs", "cs")

    def __init__(self, t, ctot, L, coff=0, cs=None):
        self.t = t
        self.ptr = t.data_ptr() + coff * L * 4
        self.bs = ctot * L
        self.cs = L if cs is None else cs   # cs = 8 * L: channel-blocked [b][c/8][l][8] (stride between 8-channel blocks)


class Plan:
    def __init__(self):
        self.ops = []
        self.meta = []
        self.keep = []
        self.launches = 0
        self.last_stream = None   # torch stream of the last use (plan build included): a use on another stream waits for it first
        self.side = None          # torch stream of the plan's side branch (ops added with side=True run there between fork and join)
        self.side_ops = set()

    def add(self, name, fn, *args, launches=1, flops=0.0, bytes_=0.0, kind="other", side=False):
        if side:
            self.side_ops.add(len(self.ops))
        self.ops.append((name, fn, args))
        self.meta.append(dict(name=name, kind=kind, flops=float(flops), bytes=float(bytes_), launches=launches))
        self.launches += launches

    def hoist_side(self, fork_at, join_at, side, dev):
        """Move the ops added with side=True to position `fork_at` of the launch list, on the stream `side`: a fork (side waits for what the main
        stream has issued so far) in front of them, a join (main waits for side) at the original position `join_at` of their first consumer.  Inside a
        CUDA-graph capture the two waits become the fork / join edges of a parallel branch."""
        if not self.side_ops or join_at is None or join_at <= fork_at:
            self.side_ops = set()
            return
        items = list(zip(self.ops, self.meta))
        moved = [items[i] for i in sorted(self.side_ops)]
        fork = ((None, lambda: side.wait_stream(torch.cuda.current_stream(dev)), ()), None)
        join = ((None, lambda: torch.cuda.current_stream(dev).wait_stream(side), ()), None)
        out, flags = [], set()
        for i in range(len(items) + 1):
            if i == fork_at:
                out.append(fork)
                for it in moved:
                    flags.add(len(out))
                    out.append(it)
            if i == join_at:
                out.append(join)
            if i < len(items) and i not in self.side_ops:
                out.append(items[i])
        self.ops, self.meta = [o for o, _ in out], [m for _, m in out]
        self.side_ops, self.side = flags, side

    def add_py(self, fn):
        self.ops.append((None, fn, ()))
        self.meta.append(None)

    def release(self):
        """Drop every buffer and launch record now (eviction from the engine's plan cache), without waiting for the garbage collector."""
        self.ops, self.meta, self.keep, self.side_ops = [], [], [], set()
        for k in [k for k, v in self.__dict__.items() if torch.is_tensor(v)]:
            delattr(self, k)

    def run(self, stream, lo=0, hi=None):
        """Issue ops [lo, hi) (default: all) on `stream`."""
        side = self.side.cuda_stream if self.side is not None else stream
        for n, (name, fn, args) in enumerate(self.ops[:hi]):
            if n < lo:
                continue
            if name is None:
                fn()
            else:
                rc = fn(*args, side if n in self.side_ops else stream)
                if rc != 0:
                    _lib.check(rc, name)

    def run_timed(self, stream):
        """Replay with a CUDA-event pair around every C-ABI call (profiling aid; not used for headline numbers)."""
        evs = []
        for (name, fn, args), meta in zip(self.ops, self.meta):
            if name is None:
                fn()
                continue
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args, stream)
            b.record()
            if rc != 0:
                _lib.check(rc, name)
            evs.append((meta, a, b))
        torch.cuda.synchronize()
        return [dict(m, ms=a.elapsed_time(b)) for m, a, b in evs]


def _fold_wn(mod):
    """w = g * v / ||v|| (norm over all dims but 0) in fp32 — torch.nn.utils.parametrizations.weight_norm semantics."""
    g = mod.parametrizations.weight.original0.detach().float()
    v = mod.parametrizations.weight.original1.detach().float()
    n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / n)


def _f(t):
    return t.detach().float().contiguous()


def _pack_conv(w):
    """[Cout, Cin, K] -> [Cin][K][Cout]"""
    return w.permute(1, 2, 0).contiguous()


class Engine:
    def __init__(self, m, device):
        self.lib = _lib.load()
        self.dev = device
        info = (C.c_int * 4)()
        _lib.check(self.lib.rvc_device_info(device.index if device.index is not None else torch.cuda.current_device(), info), "device_info")
        self.sm_count, cc_major = info[0], info[1]
        if cc_major != 10:
            raise RuntimeError(f"librvc_b200 targets sm_100a (B200); found compute capability {info[1]}.{info[2]}")
        self.cfg = dict(H=m.hidden_channels, I=m.inter_channels, Fc=m.filter_channels, nh=m.n_heads, nl=m.n_layers, ks=m.kernel_size,
                        U=m.upsample_initial_channel, G=m.gin_channels, D=m.text_enc_hidden_dim, rates=m.upsample_rates,
                        uks=m.upsample_kernel_sizes, rks=m.resblock_kernel_sizes, rds=m.resblock_dilation_sizes, sr=float(m.sr),
                        upp=m.upp, f0=m.use_f0)
        self.precision = getattr(m, "precision", "tc")
        if self.precision not in ("tc", "tc-bf16", "fp32"):
            raise ValueError(f"precision must be 'tc', 'tc-bf16' or 'fp32', got {self.precision!r}")
        self.tc = self.precision != "fp32"
        # stride-2 ConvTranspose stages with their two phases as column blocks of one plain conv (rvc_conv1d_tc_desc.up_cols)
        self.up_cols = os.environ.get("RVC_B200_UP_COLS", "1") != "0"
        self._fold(m)
        self.plans = {}
        self.max_plans = 16   # encoder + decoder plans of the 8 most recently used (B, T) shapes (a bucketed stream needs 3-6)
        self.graphs = {}          # (B, T, head) -> (CUDAGraph, plan keys it replays into), most recently used last
        self._warm = set()
        self._lock = threading.RLock()   # infer() is serialised per engine: plans, workspaces and graphs are shared per (B, T)
        self.use_graph = bool(getattr(m, "use_cuda_graph", True))
        self.poison = {"1": "nan", "nan": "nan", "big": "big"}.get(os.environ.get("RVC_B200_POISON", ""), "")
        self.pair_fusion = os.environ.get("RVC_B200_PAIR_FUSION", "1") != "0"   # fused ResBlock steps in the C = 32 / 64 stages
        # ... with the residual stream of those stages carried as two packed fp16 planes (hi = the conv operand, lo = the rounding
        # remainder: 22 bits in 4 bytes) instead of fp32 + a packed operand copy: 8 instead of 12 bytes per element and step
        self.pair_split = os.environ.get("RVC_B200_PAIR_SPLIT", "1") != "0"
        # ... and in the wide stages (C >= 128, separate conv1 / conv2 launches): conv2 reads its residual as res_hi / res_lo and writes y_act / y_lo
        # OFF by default: measured on configs[1] (profiles/r3_history.md) conv2 gets 6 us faster per launch (117 vs 123 us at C = 128) but the ups launch that
        # has to write the lo plane gets 6 us slower and the stage-average launch, which decodes three split residuals per item, 14 us slower
        self.tc_split = os.environ.get("RVC_B200_TC_SPLIT", "0") != "0"
        # dependent short-sequence launches (the flow) as one persistent kernel with grid barriers between the layers (rvc_conv1d_tc_chain_*)
        self.use_chain = os.environ.get("RVC_B200_CHAIN", "1") != "0"
        # ... or, for the standard flow shape, the whole reverse flow as one cluster kernel that keeps every intermediate on the chip (rvc_flow_reverse_tc)
        self.use_flow_cluster = os.environ.get("RVC_B200_FLOW_CLUSTER", "1") != "0"
        # attention scores + softmax of a layer as one cluster kernel for T <= 1024 (rvc_attn_scores_cluster_tc) instead of S launch + softmax launch
        self.use_attn_cluster = os.environ.get("RVC_B200_ATTN_CLUSTER", "1") != "0"
        # ... with the P.v GEMM inside it (P stays in shared memory, the key tiles' partial sums are reduce-scattered through DSMEM): one launch less per layer
        self.attn_pv = os.environ.get("RVC_B200_ATTN_PV", "1") != "0"
        # conv_o -> LayerNorm -> FFN -> LayerNorm of an encoder layer as one cluster kernel (rvc_encoder_tail_tc) instead of five launches
        self.use_enc_tail = os.environ.get("RVC_B200_ENC_TAIL", "1") != "0"
        # the NSF source (SineGen + the long noise convs: 70 us of CUDA-core work that depends on f0 only) on a side branch of the graph, next to the
        # flow kernel, which occupies 78 of the 148 SMs for 130 us.  OFF by default: measured 2.365 ms per step against 2.345 without it (round 2v; the
        # same experiment next to the per-launch flow in round 1: 2.58 vs 2.57) — the branch's CTAs delay the placement of the flow's clusters and the
        # fork / join edges take the programmatic-dependent-launch overlap away from the kernels around them.
        self.use_side_branch = os.environ.get("RVC_B200_SIDE_BRANCH", "0") != "0"
        self.side_stream = torch.cuda.Stream(self.dev)
        self._chain = None

    # ------------------------------------------------------------------ weight folding / packing
    def _fold(self, m):
        c = self.cfg
        H, I = c["H"], c["I"]
        half = I // 2
        W = {}
        e = m.enc_p
        W["emb_wT"] = _f(e.emb_phone.weight).t().contiguous()
        W["emb_b"] = _f(e.emb_phone.bias)
        W["emb_pitch"] = _f(e.emb_pitch.weight) if c["f0"] else None
        for i in range(c["nl"]):
            a = e.encoder.attn_layers[i]
            W[f"qkv_w{i}"] = _pack_conv(torch.cat([_f(a.conv_q.weight), _f(a.conv_k.weight), _f(a.conv_v.weight)], 0))
            W[f"qkv_b{i}"] = torch.cat([_f(a.conv_q.bias), _f(a.conv_k.bias), _f(a.conv_v.bias)], 0).contiguous()
            W[f"o_w{i}"], W[f"o_b{i}"] = _pack_conv(_f(a.conv_o.weight)), _f(a.conv_o.bias)
            W[f"ek{i}"], W[f"ev{i}"] = _f(a.emb_rel_k)[0].contiguous(), _f(a.emb_rel_v)[0].contiguous()
            for j, lst in ((1, e.encoder.norm_layers_1), (2, e.encoder.norm_layers_2)):
                W[f"ln{j}_g{i}"], W[f"ln{j}_b{i}"] = _f(lst[i].gamma), _f(lst[i].beta)
            f = e.encoder.ffn_layers[i]
            W[f"f1_w{i}"], W[f"f1_b{i}"] = _pack_conv(_f(f.conv_1.weight)), _f(f.conv_1.bias)
            W[f"f2_w{i}"], W[f"f2_b{i}"] = _pack_conv(_f(f.conv_2.weight)), _f(f.conv_2.bias)
        W["proj_w"], W["proj_b"] = _pack_conv(_f(e.proj.weight)), _f(e.proj.bias)

        # flow: the four Flips are folded into channel-reversed pre/post weights (residuals.py:53-58, :91-93)
        cond_w, cond_b = [], []
        for f in range(4):
            cp = m.flow.flows[2 * f]
            flipped = f in (3, 1)  # an odd number of Flips precedes couplings 3 and 1 in the reverse pass
            pre_w, post_w, post_b = _f(cp.pre.weight), _f(cp.post.weight), _f(cp.post.bias)
            if flipped:
                pre_w, post_w, post_b = pre_w.flip(1), post_w.flip(0), post_b.flip(0)
            W[f"fl{f}_pre_w"], W[f"fl{f}_pre_b"] = _pack_conv(pre_w), _f(cp.pre.bias)
            W[f"fl{f}_post_w"], W[f"fl{f}_post_b"] = _pack_conv(post_w), post_b.contiguous()
            for l in range(3):
                W[f"fl{f}_in_w{l}"] = _pack_conv(_fold_wn(cp.enc.in_layers[l]))
                W[f"fl{f}_in_b{l}"] = _f(cp.enc.in_layers[l].bias)
                rs_w, rs_b = _fold_wn(cp.enc.res_skip_layers[l]), _f(cp.enc.res_skip_layers[l].bias)
                if l < 2:
                    W[f"fl{f}_res_w{l}"], W[f"fl{f}_res_b{l}"] = _pack_conv(rs_w[:H]), rs_b[:H].contiguous()
                    W[f"fl{f}_skip_w{l}"], W[f"fl{f}_skip_b{l}"] = _pack_conv(rs_w[H:]), rs_b[H:].contiguous()
                else:
                    W[f"fl{f}_skip_w{l}"], W[f"fl{f}_skip_b{l}"] = _pack_conv(rs_w), rs_b
            cond_w.append(_fold_wn(cp.enc.cond_layer))
            cond_b.append(_f(cp.enc.cond_layer.bias))
        # all speaker-conditioning 1x1 convs (4 x cond_layer + dec.cond) as ONE conv on g (L = 1)
        cond_w.append(_f(m.dec.cond.weight))
        cond_b.append(_f(m.dec.cond.bias))
        W["cond_w"] = _pack_conv(torch.cat(cond_w, 0))
        W["cond_b"] = torch.cat(cond_b, 0).contiguous()
        self.cond_off = [f * 6 * H for f in range(4)] + [4 * 6 * H]
        self.cond_total = 4 * 6 * H + c["U"]

        d = m.dec
        W["pre_w"], W["pre_b"] = _pack_conv(_f(d.conv_pre.weight)), _f(d.conv_pre.bias)
        nk = len(c["rks"])
        # ResBlock stage operand format on the tensor-core paths: fp16 (10-bit mantissa, fp32 accumulate, fp32 residual stream)
        # by default; "tc-bf16" uses bf16 for the C = 256 stage — bf16 on the narrower stages or on the ups breaks the 1e-3
        # waveform bar (SURVEY 7 hard part 2; measured here: ups in bf16 1.37e-3, C >= 128 in bf16 1.0e-3).  fp32 CUDA cores otherwise.
        self.stage_dt = []
        tc_w = {}
        for i in range(len(c["rates"])):
            ch = c["U"] // (2 ** (i + 1))
            dt = None
            if self.tc and nk == 3 and len({len(dl) for dl in c["rds"]}) == 1 and all(k % 2 == 1 and (k - 1) // 2 * max(dl) <= 64 for k, dl in zip(c["rks"], c["rds"])):
                if ch in (128, 256):
                    # bf16 only where the 1e-3 budget allows it: C = 256 (6.8e-5 per SURVEY 7.2); bf16 at C = 128 measured 1.0e-3
                    dt = DT_BF16 if (self.precision == "tc-bf16" and ch == 256) else DT_FP16
                elif ch in (16, 32, 64):
                    dt = DT_FP16
            self.stage_dt.append(dt)
        self.up_dt = []
        for i in range(len(c["rates"])):
            up = d.ups[i] if c["f0"] else d.ups_and_resblocks[i * (nk + 1)]
            wt = _fold_wn(up)                                           # [Cin][Cout][k]
            W[f"up_w{i}"] = wt.permute(0, 2, 1).contiguous()            # [Cin][k][Cout]
            W[f"up_b{i}"] = _f(up.bias)
            cin_i, cout_i, k_i, u_i = wt.shape[0], wt.shape[1], wt.shape[2], c["rates"][i]
            self.up_dt.append(None)
            if self.tc and cin_i % 16 == 0 and cout_i % 16 == 0 and (k_i - u_i) % 2 == 0:
                # phase-major conv weight [Cout][Cin][u*M]: block r, tap t = w_T[:, :, r + (M-1-t)*u] (include/rvc_b200.h)
                M = -(-k_i // u_i)
                wk = torch.zeros(cout_i, cin_i, u_i * M)
                for r in range(u_i):
                    for t in range(M):
                        kk = r + (M - 1 - t) * u_i
                        if kk < k_i:
                            wk[:, :, r * M + t] = wt[:, :, kk].t()
                self.up_dt[i] = DT_FP16
                tc_w[f"t_up_w{i}"] = (wk, self.up_dt[i])
                # phase-column form for the stride-2 stages (rvc_conv1d_tc_desc.up_cols): the u phases side by side as u * Cout weight columns of
                # one plain conv with M taps — half the tiles, MMAs of twice the width
                if self.up_cols and u_i * cout_i <= 128 and cout_i >= 32 and cout_i & (cout_i - 1) == 0:
                    tc_w[f"t_up_wm{i}"] = (wk.view(cout_i, cin_i, u_i, M).permute(2, 0, 1, 3).reshape(u_i * cout_i, cin_i, M).contiguous(), self.up_dt[i])
                    W[f"up_bm{i}"] = W[f"up_b{i}"].repeat(u_i)
            if c["f0"]:
                W[f"nz_w{i}"] = _f(d.noise_convs[i].weight)[:, 0, :].contiguous()
                W[f"nz_b{i}"] = _f(d.noise_convs[i].bias)
            for j in range(nk):
                rb = d.resblocks[i * nk + j] if c["f0"] else d.ups_and_resblocks[i * (nk + 1) + 1 + j]
                for mm in range(len(c["rds"][j])):
                    w1, w2 = _fold_wn(rb.convs1[mm]), _fold_wn(rb.convs2[mm])
                    W[f"rb{i}_{j}_c1b{mm}"], W[f"rb{i}_{j}_c2b{mm}"] = _f(rb.convs1[mm].bias), _f(rb.convs2[mm].bias)
                    if self.stage_dt[i] is None:
                        W[f"rb{i}_{j}_c1w{mm}"], W[f"rb{i}_{j}_c2w{mm}"] = _pack_conv(w1), _pack_conv(w2)
                    else:
                        tc_w[f"rb{i}_{j}_c1w{mm}"], tc_w[f"rb{i}_{j}_c2w{mm}"] = (w1, self.stage_dt[i]), (w2, self.stage_dt[i])
        W["post_w"] = _f(d.conv_post.weight)[0].contiguous()
        if c["f0"]:
            W["lin_w"], W["lin_b"] = _f(d.m_source.l_linear.weight).view(1), _f(d.m_source.l_linear.bias).view(1)
        W["emb_g"] = _f(m.emb_g.weight)
        # ---- TextEncoder / flow / conv_pre on tensor cores (fp16 operands, fp32 accumulate) when every shape tiles
        half = I // 2
        dims_ok = all(v % 16 == 0 for v in (H, c["Fc"], half, I)) and all(v % 32 == 0 for v in (H, 3 * H, c["Fc"], 2 * I, half, c["U"])) \
            and H % 64 == 0 and c["ks"] % 2 == 1
        # (RVC_B200_ENC_TC=0 / RVC_B200_FLOW_TC=0: error-budget experiments — that part runs on the fp32 CUDA-core kernels instead)
        self.enc_split = os.environ.get("RVC_B200_ENC_SPLIT", "1") != "0"     # split-precision encoder GEMMs (see _enc_plan)
        self.enc_tc = self.tc and dims_ok and os.environ.get("RVC_B200_ENC_TC", "1") != "0"
        self.flow_tc = self.tc and dims_ok and os.environ.get("RVC_B200_FLOW_TC", "1") != "0"   # flow + conv_pre
        if self.enc_tc:
            dt = DT_FP16
            e = m.enc_p
            for i in range(c["nl"]):
                a = e.encoder.attn_layers[i]
                tc_w[f"t_qkv_w{i}"] = (torch.cat([_f(a.conv_q.weight), _f(a.conv_k.weight), _f(a.conv_v.weight)], 0), dt)
                tc_w[f"t_o_w{i}"] = (_f(a.conv_o.weight), dt)
                f = e.encoder.ffn_layers[i]
                tc_w[f"t_f1_w{i}"], tc_w[f"t_f2_w{i}"] = (_f(f.conv_1.weight), dt), (_f(f.conv_2.weight), dt)
                if c["Fc"] % 48 == 0:   # three input-channel slices of conv_2 for the split-K launch of short sequences
                    fs = c["Fc"] // 3
                    for sp in range(3):
                        tc_w[f"t_f2_w{i}_k{sp}"] = (_f(f.conv_2.weight)[:, sp * fs:(sp + 1) * fs].contiguous(), dt)
            tc_w["t_proj_w"] = (_f(e.proj.weight), dt)
            if c["D"] % 16 == 0:
                tc_w["t_emb_w"] = (_f(e.emb_phone.weight).unsqueeze(-1).contiguous(), dt)   # Linear as a 1x1 conv [H, D, 1]
            # attention cluster kernel: emb_rel_k as packed fp16 planes [hi | lo][dk/8][32 rows][8] (row = relative position), emb_rel_v^T as the
            # four extra weight chunks of the P.v GEMM [4][dk][8] (relative positions as "keys")
            if c["nh"] * 96 == H and 2 * WINDOW + 1 <= 24:
                for i in range(c["nl"]):
                    a = e.encoder.attn_layers[i]
                    ek = torch.zeros(32, 96)
                    ek[:2 * WINDOW + 1] = _f(a.emb_rel_k)[0]
                    hi = ek.half()
                    planes = torch.stack([hi, (ek - hi.float()).half()])                        # [2][32][96]
                    W[f"t_ek_p{i}"] = planes.view(2, 32, 12, 8).permute(0, 2, 1, 3).contiguous().view(torch.uint8).reshape(-1)
                    ev = torch.zeros(32, 96)
                    ev[:2 * WINDOW + 1] = _f(a.emb_rel_v)[0]
                    W[f"t_evT{i}"] = ev.view(4, 8, 96).permute(0, 2, 1).contiguous().half().view(torch.uint8).reshape(-1)   # [4][96][8]
            if self.enc_split:   # low weight planes of the split-precision GEMMs: fp16(W - fp16(W))
                for k in [k for k in tc_w if k.startswith(("t_qkv_w", "t_o_w", "t_proj_w", "t_emb_w"))]:
                    w_ = tc_w[k][0]
                    tc_w[k + "_lo"] = ((w_ - w_.half().float()).contiguous(), dt)
        if self.flow_tc:
            dt = DT_FP16
            # gated in_layers: weight columns ordered per 64-wide N tile as [32 tanh | 32 sigmoid] (see include/rvc_b200.h)
            perm = []
            for t in range(2 * H // 64):
                perm += list(range(t * 32, t * 32 + 32)) + list(range(H + t * 32, H + t * 32 + 32))
            perm = torch.tensor(perm)
            cond_w2, cond_b2 = [], []
            for f in range(4):
                cp = m.flow.flows[2 * f]
                flipped = f in (3, 1)
                pre_w, post_w = _f(cp.pre.weight), _f(cp.post.weight)
                if flipped:
                    pre_w, post_w = pre_w.flip(1), post_w.flip(0)
                # (post: Cout = I/2 = 96 rows, padded with zero rows to a multiple of the chain kernel's 64-column tile)
                self.post_rows = -(-post_w.shape[0] // 64) * 64
                post_w = torch.cat([post_w, post_w.new_zeros(self.post_rows - post_w.shape[0], *post_w.shape[1:])], 0)
                tc_w[f"t_fl{f}_pre_w"], tc_w[f"t_fl{f}_post_w"] = (pre_w.contiguous(), dt), (post_w.contiguous(), dt)
                cw, cb = _fold_wn(cp.enc.cond_layer), _f(cp.enc.cond_layer.bias)
                for l in range(3):
                    tc_w[f"t_fl{f}_in_w{l}"] = (_fold_wn(cp.enc.in_layers[l])[perm].contiguous(), dt)
                    W[f"t_fl{f}_in_b{l}"] = _f(cp.enc.in_layers[l].bias)[perm].contiguous()
                    rs_w = _fold_wn(cp.enc.res_skip_layers[l])
                    if l < 2:
                        tc_w[f"t_fl{f}_res_w{l}"], tc_w[f"t_fl{f}_skip_w{l}"] = (rs_w[:H].contiguous(), dt), (rs_w[H:].contiguous(), dt)
                    else:
                        tc_w[f"t_fl{f}_skip_w{l}"] = (rs_w.contiguous(), dt)
                    cond_w2.append(cw[l * 2 * H:(l + 1) * 2 * H][perm])
                    cond_b2.append(cb[l * 2 * H:(l + 1) * 2 * H][perm])
            # rvc_flow_reverse_tc accumulates the hidden stream / skip sum in TMEM and adds the biases when it reads them
            for f in range(4):
                acc = W[f"fl{f}_pre_b"].clone()
                for l in range(3):
                    W[f"t_fl{f}_hb{l}"] = acc.clone()
                    if l < 2:
                        acc = acc + W[f"fl{f}_res_b{l}"]
                W[f"t_fl{f}_sb"] = (W[f"fl{f}_skip_b0"] + W[f"fl{f}_skip_b1"] + W[f"fl{f}_skip_b2"]).contiguous()
            cond_w2.append(_f(m.dec.cond.weight))
            cond_b2.append(_f(m.dec.cond.bias))
            W["cond_w"] = _pack_conv(torch.cat(cond_w2, 0))      # same layout as before, gate-permuted rows
            W["cond_b"] = torch.cat(cond_b2, 0).contiguous()
            tc_w["t_pre_w"] = (_f(m.dec.conv_pre.weight), dt)
        self.W = {k: (v.to(self.dev) if v is not None else None) for k, v in W.items()}
        # tensor-core weights: [Cout][Cin][K] fp32 -> packed [K][Cin/CK][Cout][CK] in bf16 / tf32 (device-side pack kernel)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        with torch.cuda.device(self.dev):
            for k, (w, dt) in tc_w.items():
                wd = w.to(self.dev).contiguous()
                out = torch.empty(wd.numel() * 2, dtype=torch.uint8, device=self.dev)
                _lib.check(self.lib.rvc_pack_weight(wd.data_ptr(), out.data_ptr(), wd.shape[0], wd.shape[1], wd.shape[2], dt, st), "pack_weight")
                self.W[k] = out
            # speaker conditioning is a function of sid alone ("precompute per sid"): one row of cond_layer(g) x 4 | dec.cond(g)
            # per speaker, computed once by the same GEMV kernel the plans would run (bit-identical), gathered by sid at infer time
            self.cond_table = None
            nspk = self.W["emb_g"].shape[0]
            if nspk * self.cond_total * 4 <= (256 << 20) and os.environ.get("RVC_B200_COND_TABLE", "1") != "0":
                tab = torch.empty(nspk, self.cond_total, dtype=torch.float32, device=self.dev)
                for s0 in range(0, nspk, 4096):
                    n = min(4096, nspk - s0)
                    _lib.check(self.lib.rvc_cond_gemv_f32(self.W["emb_g"][s0:s0 + n].contiguous().data_ptr(), self.W["cond_w"].data_ptr(), self.W["cond_b"].data_ptr(),
                                                          tab[s0:s0 + n].data_ptr(), n, c["G"], self.cond_total, st), "cond table")
                self.cond_table = tab
            torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------ descriptor helpers
    def _conv(self, plan, name, x, w, y, *, B, Cin, Cout, L, K=1, dil=1, pad=0, bias=None, bias_b=None, bias_b_bs=0, mask=None,
              in_mask=False, out_mask=False, res=None, pre_slope=None, act=ACT_NONE, gated=False, alpha=1.0, out_scale=1.0, beta=0.0):
        d = Conv1dDesc()
        d.x, d.x_bs, d.x_cs = x.ptr, x.bs, x.cs
        d.w = w.data_ptr()
        d.bias = bias.data_ptr() if bias is not None else None
        d.bias_b = bias_b if bias_b is not None else None
        d.bias_b_bs = bias_b_bs
        d.mask = mask.data_ptr() if mask is not None else None
        if res is not None:
            d.res, d.res_bs, d.res_cs = res.ptr, res.bs, res.cs
        d.y, d.y_bs, d.y_cs = y.ptr, y.bs, y.cs
        d.B, d.Cin, d.Cout, d.L, d.K, d.dil, d.pad = B, Cin, Cout, L, K, dil, pad
        d.pre_lrelu, d.pre_slope = (1, pre_slope) if pre_slope is not None else (0, 0.0)
        d.in_mask, d.out_mask = int(in_mask), int(out_mask)
        d.act, d.gated = act, int(gated)
        d.alpha, d.out_scale, d.beta = alpha, out_scale, beta
        plan.keep.append(d)
        g2 = 2 if gated else 1
        flops = 2.0 * B * Cin * Cout * g2 * K * L
        bytes_ = 4.0 * (B * Cin * L + B * Cout * L * (1 + (res is not None) + (beta != 0.0)) + Cin * K * Cout * g2)
        plan.add(name, self.lib.rvc_conv1d_f32, C.byref(d), flops=flops, bytes_=bytes_, kind=f"conv1d_f32[C{Cout}k{K}]")

    def _tc(self, plan, name, jobs, *, B, Cin, Cout, L, dtype, nt=0, mt=0, tag="", up=1, pshift=0, flops=None, blocked=0, sum_jobs=0, cout_valid=None,
            up_cols=0):
        """One persistent tcgen05 launch over up to six jobs (dicts of rvc_tc_job fields; V views for res / y)."""
        d = Conv1dTcDesc()
        d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype, d.nt, d.mt = len(jobs), B, Cin, Cout, L, dtype, nt, mt
        d.up, d.pshift, d.f32_blocked, d.sum_jobs, d.up_cols = up, pshift, blocked, sum_jobs, up_cols
        fl = by = 0.0
        for slot, j in enumerate(jobs):
            jb = d.jobs[slot]
            jb.a = j["a"] if isinstance(j["a"], int) else j["a"].data_ptr()
            jb.w = j["w"] if isinstance(j["w"], int) else j["w"].data_ptr()
            jb.bias = j["bias"].data_ptr() if j.get("bias") is not None else None
            if j.get("bias_b") is not None:
                jb.bias_b, jb.bias_b_bs = j["bias_b"], j["bias_b_bs"]
            if j.get("mask") is not None:
                jb.mask = j["mask"].data_ptr()
            if j.get("res") is not None:
                jb.res, jb.res_bs, jb.res_cs = j["res"].ptr, j["res"].bs, j["res"].cs
            if j.get("y") is not None:
                jb.y_f32, jb.y_bs, jb.y_cs = j["y"].ptr, j["y"].bs, j["y"].cs
            if j.get("y_act") is not None:
                jb.y_act = j["y_act"] if isinstance(j["y_act"], int) else j["y_act"].data_ptr()
            if j.get("y_lo") is not None:
                jb.y_lo = j["y_lo"] if isinstance(j["y_lo"], int) else j["y_lo"].data_ptr()
            jb.K, jb.dil = j.get("K", 1), j.get("dil", 1)
            jb.pad = j.get("pad", (jb.K - 1) // 2 * jb.dil)
            jb.gated, jb.relu = int(j.get("gated", False)), int(j.get("relu", 0))
            jb.alpha, jb.act_slope = j.get("alpha", 1.0), j.get("act_slope", 1.0)
            jb.a_bs, jb.w_bs, jb.y_act_bs, jb.w_cs = j.get("a_bs", 0), j.get("w_bs", 0), j.get("y_act_bs", 0), j.get("w_cs", 0)
            if j.get("nz_har"):
                jb.nz_har, jb.nz_w, jb.nz_b, jb.nz_har_bs = j["nz_har"], j["nz_w"], j["nz_b"], j["nz_har_bs"]
                jb.nz_kn, jb.nz_sn, jb.nz_pn = j["nz_kn"], j["nz_sn"], j["nz_pn"]
            ce = Cout // 2 if jb.gated else Cout      # (phase columns: Cout = up_cols * Cr output values per input position)
            fl += 2.0 * B * Cin * Cout * jb.K * L
            by += B * L * (2.0 * Cin + up * ce * (4.0 * (j.get("res") is not None) + 4.0 * (j.get("y") is not None) + 2.0 * (j.get("y_act") is not None) + 2.0 * (j.get("y_lo") is not None)))
        plan.keep.append(d)
        plan.keep.append(jobs)
        if self._chain is not None:     # inside a chain scope: collected, launched as one kernel by _chain_end
            self._chain.append((d, cout_valid or 0, fl if flops is None else flops, by))
            return
        plan.add(name, self.lib.rvc_conv1d_tc, C.byref(d), flops=fl if flops is None else flops, bytes_=by,
                 kind=f"conv1d_tc_{'bf16' if dtype == DT_BF16 else 'fp16'}[{tag or 'C%d' % Cout}]")

    def _chain_begin(self):
        self._chain = []

    def _chain_end(self, plan, name, tag):
        """Turn the rvc_conv1d_tc descriptors collected since _chain_begin into ONE persistent launch (rvc_conv1d_tc_chain_run)."""
        items, self._chain = self._chain, None
        n = len(items)
        descs = (Conv1dTcDesc * n)(*[it[0] for it in items])
        cv = (C.c_int32 * n)(*[it[1] for it in items])
        nbytes = int(self.lib.rvc_conv1d_tc_chain_blob_bytes(n))
        blob = (C.c_uint8 * nbytes)()
        _lib.check(self.lib.rvc_conv1d_tc_chain_build(descs, cv, n, C.addressof(blob), nbytes), "rvc_conv1d_tc_chain_build")
        dev_blob = torch.frombuffer(blob, dtype=torch.uint8).to(self.dev)
        sync = torch.zeros(4, dtype=torch.int32, device=self.dev)
        plan.keep += [descs, cv, blob, dev_blob, sync]
        plan.chain_flags = getattr(plan, "chain_flags", []) + [sync]
        plan.add(name, self.lib.rvc_conv1d_tc_chain_run, C.addressof(blob), dev_blob.data_ptr(), sync.data_ptr(), flops=sum(it[2] for it in items),
                 bytes_=sum(it[3] for it in items), kind=f"conv1d_tc_chain[{tag}]")

    def _packed(self, plan, B, Cch, L):
        return self._zbuf(plan, B * Cch * int(self.lib.rvc_packed_rows(L)) * 2)

    GUARD = 1 << 16   # poison mode: bytes of 0x7e (fp16 / bf16 / fp32 NaN-or-huge patterns) behind every workspace

    def _alloc(self, plan, nbytes):
        """Raw workspace bytes.  With RVC_B200_POISON set every workspace is followed by GUARD bytes of 0x7e, so a read past the end
        of a buffer (or of a packed chunk's last rows) meets a NaN / 1e38 pattern instead of whatever the allocator placed there."""
        if not self.poison:
            t = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.dev)
            plan.keep.append(t)
            return t
        full = torch.empty(nbytes + self.GUARD, dtype=torch.uint8, device=self.dev)
        full[nbytes:].fill_(0x7E)
        plan.keep.append(full)
        return full[:nbytes]

    def _zbuf(self, plan, nbytes):
        t = self._alloc(plan, nbytes)
        t.zero_()
        return t

    def _buf(self, plan, *shape, dtype=torch.float32):
        n = int(math.prod(shape)) if shape else 1
        t = self._alloc(plan, n * torch.empty((), dtype=dtype).element_size())[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)
        if self.poison:
            # debugging aid (RVC_B200_POISON=nan | big): every workspace starts as NaN (or 1e30) — integer / byte workspaces as 0x7e bytes
            # (fp16 NaN) — so an element that a kernel reads without any kernel having written it shows up in the output instead of
            # depending on what the allocator's block held before
            if t.is_floating_point():
                t.fill_(float("nan") if self.poison == "nan" else 1e30)
            else:
                t.view(torch.uint8).fill_(0x7E)
        return t

    # ------------------------------------------------------------------ plans
    def _enc_plan(self, B, T):
        c, W, lib = self.cfg, self.W, self.lib
        H, I, Fc, D, ks = c["H"], c["I"], c["Fc"], c["D"], c["ks"]
        p = Plan()
        p.phone = self._buf(p, B, T, D)
        p.pitch = self._buf(p, B, T, dtype=torch.int64) if c["f0"] else None
        p.lengths = self._buf(p, B, dtype=torch.int64)
        p.eps_z = self._buf(p, B, I, T)
        p.mask = self._buf(p, B, T)
        x, qkv, att, y = self._buf(p, B, H, T), self._buf(p, B, 3 * H, T), self._buf(p, B, H, T), self._buf(p, B, H, T)
        h = self._buf(p, B, Fc, T)
        p.stats = self._buf(p, B, 2 * I, T)
        p.zp = self._buf(p, B, I, T)
        vx, vqkv, vatt, vy, vh, vst = V(x, H, T), V(qkv, 3 * H, T), V(att, H, T), V(y, H, T), V(h, Fc, T), V(p.stats, 2 * I, T)
        padl = (ks - 1) // 2
        tc_embed = self.enc_tc and "t_emb_w" in W
        if not tc_embed:
            p.add("embed", lib.rvc_embed_f32, p.phone.data_ptr(), p.pitch.data_ptr() if c["f0"] else None, p.lengths.data_ptr(),
                  W["emb_wT"].data_ptr(), W["emb_b"].data_ptr(), W["emb_pitch"].data_ptr() if c["f0"] else None, x.data_ptr(),
                  p.mask.data_ptr(), B, T, D, H, flops=2.0 * B * T * D * H, bytes_=4.0 * B * T * (D + H), kind="embed_f32")
        if self.enc_tc:
            dt = DT_FP16
            nh = c["nh"]
            dk = H // nh
            tc_attn = nh <= 3 and dk % 32 == 0
            Lp = int(lib.rvc_packed_rows(T))
            # Split-precision GEMMs (self.enc_split): the error of the tensor-core path's waveform is dominated by the fp16 operand
            # rounding of the encoder's 1x1 GEMMs (every site feeds the softmax logits, tests/encoder_rounding_study.py), and those
            # launches are latency-bound — so they run as a.W ~= a_hi.W_hi + a_lo.W_hi + a_hi.W_lo: three jobs of one accumulator
            # (rvc_conv1d_tc_desc.sum_jobs = 3) over [hi | lo] activation planes and two weight planes.  FFN and P.v stay plain fp16.
            split = self.enc_split and tc_embed and tc_attn
            npl = 2 if split else 1
            adt = _lib.DT_FP16_SPLIT if split else dt
            pk = lambda Cch: self._zbuf(p, B * npl * Cch * Lp * 2)        # packed activations, [hi | lo] planes when split
            xa, atta, ha = pk(H), pk(H), self._packed(p, B, Fc, T)
            y2, y3 = self._buf(p, B, H, T), self._buf(p, B, H, T)      # split-K partial sums of the FFN's second conv
            cb = Lp * 16                                                   # bytes per 8-channel chunk of a packed buffer

            def gemm(name, a, cin, wkey, cout, lead, tag="enc", wlo=None):
                """1x1 conv over the packed buffer `a` (cin channels per plane); `lead` = epilogue options of the (leading) job."""
                a_bs = npl * (cin // 8) * cb
                if not split:
                    self._tc(p, name, [dict(a=a, w=W[wkey], a_bs=a_bs, **lead)], B=B, Cin=cin, Cout=cout, L=T, dtype=dt, tag=tag)
                    return
                base = a.data_ptr()
                jobs = [dict(a=base, w=W[wkey], a_bs=a_bs, **lead), dict(a=base + (cin // 8) * cb, w=W[wkey], a_bs=a_bs),
                        dict(a=base, w=W[wlo or wkey + "_lo"], a_bs=a_bs)]
                self._tc(p, name, jobs, B=B, Cin=cin, Cout=cout, L=T, dtype=dt, tag=tag, sum_jobs=3, flops=2.0 * B * cin * cout * T)

            def outs(buf, cout):
                """packed output planes of a launch writing `cout` channels per plane into `buf`"""
                d = dict(y_act=buf, y_act_bs=npl * (cout // 8) * cb)
                if split:
                    d["y_lo"] = buf.data_ptr() + (cout // 8) * cb
                return d

            if tc_embed:
                # x = lrelu((phone.W^T + b + E[pitch]) * sqrt(H), 0.1) * mask  (synthesizers.py:367): exact gather + tensor-core Linear
                pha = pk(D)
                pbias = self._buf(p, B, H, T) if c["f0"] else None
                sq = math.sqrt(H)
                # plain: v = E[pitch]*sqrt(H) + sqrt(H)*(acc + b); split (sum mode): v = sqrt(H)*(E[pitch] + acc + b), the reference's own order
                p.add("embed_prep", lib.rvc_embed_prep, p.pitch.data_ptr() if c["f0"] else None, p.lengths.data_ptr(),
                      W["emb_pitch"].data_ptr() if c["f0"] else None, 1.0 if split else sq, pbias.data_ptr() if c["f0"] else None, p.mask.data_ptr(), B, T, H,
                      bytes_=8.0 * B * H * T, kind="embed_prep")
                p.add("pack_phone", lib.rvc_pack_rows, p.phone.data_ptr(), pha.data_ptr(), B, T, D, adt, bytes_=(4.0 + 2.0 * npl) * B * T * D, kind="pack_act")
                gemm("embed", pha, D, "t_emb_w", H, dict(bias=W["emb_b"], res=(V(pbias, H, T) if c["f0"] else None), alpha=sq,
                                                          relu=2, act_slope=0.1, mask=p.mask, y=vx, **outs(xa, H)))
            else:
                p.add("pack_x", lib.rvc_pack_act, x.data_ptr(), None, None, 1.0, H * T, T, None, xa.data_ptr(), B, H, T, dt, 1.0, kind="pack_act",
                      bytes_=6.0 * B * H * T)
            attn_cl = tc_attn and self.use_attn_cluster and dk == 96 and T <= 1024 and "t_ek_p0" in W
            enc_tail = split and self.use_enc_tail and (H, Fc, ks) == (192, 768, 3)
            xcur, xnxt = x, y        # (with the tail kernel the residual stream alternates between the two fp32 buffers)
            if tc_attn:
                Tk = (T + 127) // 128 * 128
                qkva = pk(3 * H)
                qkv_bs = npl * (3 * H // 8) * cb
            if tc_attn and not attn_cl:
                Sb = self._buf(p, B, nh, Tk, T)
                Pb = self._zbuf(p, B * nh * (Tk // 8) * Lp * 16)
                Rb = self._buf(p, B, H, T)
                vwb = self._buf(p, B * nh * (Tk // 8) * dk * 4, dtype=torch.int32)
            attn_pv = attn_cl and self.attn_pv
            if attn_cl and not attn_pv:
                Tk = (T + 255) // 256 * 256                        # one 256-key tile per CTA of the cluster
                nct = Tk // 8 + 4                                  # key chunks + the four "relative position" chunks of P / of the P.v weights
                Pb = self._zbuf(p, B * nh * nct * Lp * 16)         # (band entries whose key falls outside the sequence stay zero)
                vws = []
                for i in range(c["nl"]):                           # per layer: v^T is rewritten every call, the emb_rel_v^T tail is constant
                    vw_i = self._zbuf(p, B * nh * nct * dk * 16)
                    vw_i.view(B * nh, nct, dk * 16)[:, Tk // 8:, :] = W[f"t_evT{i}"].view(1, 4, dk * 16)
                    vws.append(vw_i)
            elif attn_cl:
                Tk = (T + 255) // 256 * 256
            for i in range(c["nl"]):
                gemm(f"qkv{i}", xa, H, f"t_qkv_w{i}", 3 * H, dict(bias=W[f"qkv_b{i}"], y=vqkv, **(outs(qkva, 3 * H) if tc_attn else {})))
                if attn_cl:
                    # scores + softmax in one cluster kernel; P carries the band probabilities as 32 extra columns, the P.v weights carry emb_rel_v^T
                    ad = AttnScoresDesc()
                    ad.qk, ad.qk_bs, ad.q_chunk0, ad.k_chunk0, ad.lo_chunks, ad.planes = qkva.data_ptr(), qkv_bs, 0, H // 8, 3 * H // 8, npl
                    ad.ek, ad.qkv, ad.mask = W[f"t_ek_p{i}"].data_ptr(), qkv.data_ptr(), p.mask.data_ptr()
                    ad.B, ad.H, ad.n_heads, ad.T, ad.window = B, H, nh, T, WINDOW
                    p.keep.append(ad)
                    if attn_pv:
                        # ... and P.v (+ the relative-value term) in the same kernel: it writes the packed operand of conv_o directly
                        ad.att, ad.att_bs, ad.att_lo_chunks, ad.evT = atta.data_ptr(), npl * (H // 8) * cb, (H // 8 if split else 0), W[f"t_evT{i}"].data_ptr()
                        p.add(f"attn_sc{i}", lib.rvc_attn_scores_cluster_tc, C.byref(ad), flops=4.0 * B * H * T * T, bytes_=B * (2.0 * 4 * H * T + 4.0 * H * T),
                              kind="attn_scores_cluster_tc")
                    else:
                        ad.P, ad.vw = Pb.data_ptr(), vws[i].data_ptr()
                        p.add(f"attn_sc{i}", lib.rvc_attn_scores_cluster_tc, C.byref(ad), flops=2.0 * B * H * T * T, bytes_=B * nh * (Tk * T * 2.0 + 2.0 * 4 * H // nh * T),
                              kind="attn_scores_cluster_tc")
                    if not attn_pv:
                        jobs = []
                        for h in range(nh):
                            ya = atta.data_ptr() + (h * dk // 8) * cb
                            jobs.append(dict(a=Pb.data_ptr() + h * nct * cb, a_bs=nh * nct * cb, w=vws[i].data_ptr() + h * nct * dk * 16, w_bs=nh * nct * dk * 16,
                                             y_act=ya, y_act_bs=npl * (H // 8) * cb, y_lo=(ya + (H // 8) * cb if split else None)))
                        self._tc(p, f"attn_pv{i}", jobs, B=B, Cin=Tk + 32, Cout=dk, L=T, dtype=dt, tag="attn", flops=2.0 * B * H * T * T)
                elif tc_attn:
                    # S[b][h][key/8][query][8] = q . k^T (channel-blocked over the keys): the packed k activations of each batch item are the GEMM's weights
                    jobs = []
                    for h in range(nh):
                        q_hi, k_hi = qkva.data_ptr() + (h * dk // 8) * cb, qkva.data_ptr() + (((H + h * dk) // 8) * Lp + 64) * 16
                        lo = (3 * H // 8) * cb
                        terms = [(q_hi, k_hi), (q_hi + lo, k_hi), (q_hi, k_hi + lo)] if split else [(q_hi, k_hi)]
                        for t_i, (qa, kw) in enumerate(terms):
                            jobs.append(dict(a=qa, a_bs=qkv_bs, w=kw, w_bs=qkv_bs, w_cs=Lp, y=(V(Sb, nh * Tk, T, h * Tk, cs=8 * T) if t_i == 0 else None)))
                    self._tc(p, f"attn_qk{i}", jobs, B=B, Cin=dk, Cout=Tk, L=T, dtype=dt, tag="attn", flops=2.0 * B * H * T * T, blocked=2,
                             sum_jobs=3 if split else 0)
                    p.add(f"attn_sm{i}", lib.rvc_attn_softmax, Sb.data_ptr(), qkv.data_ptr(), p.mask.data_ptr(), W[f"ek{i}"].data_ptr(),
                          W[f"ev{i}"].data_ptr(), Pb.data_ptr(), Rb.data_ptr(), vwb.data_ptr(), B, H, nh, T, Tk, WINDOW,
                          bytes_=B * nh * Tk * T * 10.0, kind="attn_softmax")
                    # O = P . v + R, written straight into the packed input of conv_o
                    jobs = []
                    for h in range(nh):
                        ya = atta.data_ptr() + (h * dk // 8) * cb
                        jobs.append(dict(a=Pb.data_ptr() + h * (Tk // 8) * cb, a_bs=nh * (Tk // 8) * cb,
                                         w=vwb.data_ptr() + h * (Tk // 8) * dk * 16, w_bs=nh * (Tk // 8) * dk * 16,
                                         res=V(Rb, H, T, h * dk), y_act=ya, y_act_bs=npl * (H // 8) * cb, y_lo=(ya + (H // 8) * cb if split else None)))
                    self._tc(p, f"attn_pv{i}", jobs, B=B, Cin=Tk, Cout=dk, L=T, dtype=dt, tag="attn", flops=2.0 * B * H * T * T)
                else:
                    p.add(f"attn{i}", lib.rvc_rel_attention_f32, qkv.data_ptr(), p.mask.data_ptr(), W[f"ek{i}"].data_ptr(), W[f"ev{i}"].data_ptr(),
                          att.data_ptr(), B, H, nh, T, WINDOW, flops=4.0 * B * H * T * T, bytes_=16.0 * B * H * T, kind="rel_attention_f32")
                    p.add(f"pack_att{i}", lib.rvc_pack_act, att.data_ptr(), None, None, 1.0, H * T, T, None, atta.data_ptr(), B, H, T, dt, 1.0,
                          kind="pack_act", bytes_=6.0 * B * H * T)
                if enc_tail:
                    # the rest of the layer in one cluster kernel: x (fp32, ping-pong between two buffers: neighbouring tiles read each other's halo rows) -> x'
                    td = EncoderTailDesc()
                    td.att, td.att_bs, td.x, td.y = atta.data_ptr(), npl * (H // 8) * cb, xcur.data_ptr(), xnxt.data_ptr()
                    td.ya, td.ya_bs, td.mask = xa.data_ptr(), npl * (H // 8) * cb, p.mask.data_ptr()
                    td.wo_hi, td.wo_lo, td.bo = W[f"t_o_w{i}"].data_ptr(), W[f"t_o_w{i}_lo"].data_ptr(), W[f"o_b{i}"].data_ptr()
                    td.w1, td.b1, td.w2, td.b2 = W[f"t_f1_w{i}"].data_ptr(), W[f"f1_b{i}"].data_ptr(), W[f"t_f2_w{i}"].data_ptr(), W[f"f2_b{i}"].data_ptr()
                    td.gamma1, td.beta1, td.gamma2, td.beta2 = (W[f"ln1_g{i}"].data_ptr(), W[f"ln1_b{i}"].data_ptr(), W[f"ln2_g{i}"].data_ptr(),
                                                                W[f"ln2_b{i}"].data_ptr())
                    td.B, td.T, td.hidden, td.filter, td.kernel_size, td.out_mask, td.eps = B, T, H, Fc, ks, int(i == c["nl"] - 1), 1e-5
                    p.keep.append(td)
                    p.add(f"enc_tail{i}", lib.rvc_encoder_tail_tc, C.byref(td), flops=2.0 * B * T * (H * H + 2 * ks * H * Fc), bytes_=B * H * T * (4.0 + 4.0 + 4.0 + 4.0),
                          kind="encoder_tail_tc")
                    xcur, xnxt = xnxt, xcur
                    continue
                gemm(f"attn_o{i}", atta, H, f"t_o_w{i}", H, dict(bias=W[f"o_b{i}"], y=vy))
                p.add(f"ln1_{i}", lib.rvc_layernorm_f32, x.data_ptr(), y.data_ptr(), None, None, W[f"ln1_g{i}"].data_ptr(), W[f"ln1_b{i}"].data_ptr(),
                      p.mask.data_ptr(), x.data_ptr(), B, H, T, 1e-5, 0, xa.data_ptr(), adt, kind="layernorm_f32", bytes_=(12.0 + 2.0 * npl) * B * H * T)
                self._tc(p, f"ffn1_{i}", [dict(a=xa, a_bs=npl * (H // 8) * cb, w=W[f"t_f1_w{i}"], bias=W[f"f1_b{i}"], relu=True, mask=p.mask, y_act=ha, K=ks)],
                         B=B, Cin=H, Cout=Fc, L=T, dtype=dt, tag="enc")
                # conv_2 (Fc -> H, Fc = 4H input channels): with few position tiles the launch is a handful of CTAs streaming long K
                # loops (14.5 us at T = 1000).  Three input-channel slices as three jobs fill the GPU (7.2 us); LayerNorm adds the
                # three partial sums.
                ksplit = f"t_f2_w{i}_k0" in W and B * (-(-T // 128)) * (H // 32) < 100
                if ksplit:
                    fs = Fc // 3
                    jobs = [dict(a=ha.data_ptr() + sp * (fs // 8) * cb, a_bs=(Fc // 8) * cb, w=W[f"t_f2_w{i}_k{sp}"],
                                 bias=(W[f"f2_b{i}"] if sp == 0 else None), mask=p.mask, y=(vy, V(y2, H, T), V(y3, H, T))[sp], K=ks) for sp in range(3)]
                    self._tc(p, f"ffn2_{i}", jobs, B=B, Cin=fs, Cout=H, L=T, dtype=dt, tag="enc")
                else:
                    self._tc(p, f"ffn2_{i}", [dict(a=ha, w=W[f"t_f2_w{i}"], bias=W[f"f2_b{i}"], mask=p.mask, y=vy, K=ks)],
                             B=B, Cin=Fc, Cout=H, L=T, dtype=dt, tag="enc")
                p.add(f"ln2_{i}", lib.rvc_layernorm_f32, x.data_ptr(), y.data_ptr(), y2.data_ptr() if ksplit else None, y3.data_ptr() if ksplit else None,
                      W[f"ln2_g{i}"].data_ptr(), W[f"ln2_b{i}"].data_ptr(),
                      p.mask.data_ptr(), x.data_ptr(), B, H, T, 1e-5, 1 if i == c["nl"] - 1 else 0, xa.data_ptr(), adt, kind="layernorm_f32",
                      bytes_=(12.0 + 2.0 * npl) * B * H * T)
            gemm("proj", xa, H, "t_proj_w", 2 * I, dict(bias=W["proj_b"], mask=p.mask, y=vst))
            p.add("sample_zp", lib.rvc_sample_zp_f32, p.stats.data_ptr(), p.eps_z.data_ptr(), p.mask.data_ptr(), p.zp.data_ptr(), B, I, T)
            return p
        for i in range(c["nl"]):
            self._conv(p, f"qkv{i}", vx, W[f"qkv_w{i}"], vqkv, B=B, Cin=H, Cout=3 * H, L=T, bias=W[f"qkv_b{i}"])
            p.add(f"attn{i}", lib.rvc_rel_attention_f32, qkv.data_ptr(), p.mask.data_ptr(), W[f"ek{i}"].data_ptr(), W[f"ev{i}"].data_ptr(),
                  att.data_ptr(), B, H, c["nh"], T, WINDOW, flops=4.0 * B * H * T * T, bytes_=16.0 * B * H * T, kind="rel_attention_f32")
            self._conv(p, f"attn_o{i}", vatt, W[f"o_w{i}"], vy, B=B, Cin=H, Cout=H, L=T, bias=W[f"o_b{i}"])
            p.add(f"ln1_{i}", lib.rvc_layernorm_f32, x.data_ptr(), y.data_ptr(), None, None, W[f"ln1_g{i}"].data_ptr(), W[f"ln1_b{i}"].data_ptr(),
                  p.mask.data_ptr(), x.data_ptr(), B, H, T, 1e-5, 0, None, 0)
            self._conv(p, f"ffn1_{i}", vx, W[f"f1_w{i}"], vh, B=B, Cin=H, Cout=Fc, L=T, K=ks, pad=padl, bias=W[f"f1_b{i}"], mask=p.mask,
                       in_mask=True, act=ACT_RELU)
            self._conv(p, f"ffn2_{i}", vh, W[f"f2_w{i}"], vy, B=B, Cin=Fc, Cout=H, L=T, K=ks, pad=padl, bias=W[f"f2_b{i}"], mask=p.mask,
                       in_mask=True, out_mask=True)
            p.add(f"ln2_{i}", lib.rvc_layernorm_f32, x.data_ptr(), y.data_ptr(), None, None, W[f"ln2_g{i}"].data_ptr(), W[f"ln2_b{i}"].data_ptr(),
                  p.mask.data_ptr(), x.data_ptr(), B, H, T, 1e-5, 1 if i == c["nl"] - 1 else 0, None, 0)
        self._conv(p, "proj", vx, W["proj_w"], vst, B=B, Cin=H, Cout=2 * I, L=T, bias=W["proj_b"], mask=p.mask, out_mask=True)
        p.add("sample_zp", lib.rvc_sample_zp_f32, p.stats.data_ptr(), p.eps_z.data_ptr(), p.mask.data_ptr(), p.zp.data_ptr(), B, I, T)
        return p

    def _dec_plan(self, B, T):
        c, W, lib = self.cfg, self.W, self.lib
        H, I, U, G = c["H"], c["I"], c["U"], c["G"]
        half = I // 2
        upp = c["upp"]
        p = Plan()
        # the four results the decoder plan hands back (o, mask, z, z_p) share one arena, so infer() clones them with a single copy
        Lfin_ = T * c["upp"]
        seg = lambda n: -(-n // 64) * 64                       # segments start 256-byte aligned
        n_o, n_m, n_z = seg(B * Lfin_), seg(B * T), seg(B * I * T)
        p.out_arena = self._buf(p, n_o + n_m + 2 * n_z)
        p.o = p.out_arena[:B * Lfin_].view(B, 1, Lfin_)
        p.mask = p.out_arena[n_o:n_o + B * T].view(B, T)
        p.z = p.out_arena[n_o + n_m:n_o + n_m + B * I * T].view(B, I, T)
        p.zp = p.out_arena[n_o + n_m + n_z:n_o + n_m + n_z + B * I * T].view(B, I, T)
        p.out_offsets = (n_o, n_m, n_z)
        p.g = self._buf(p, B, G)
        gc = self._buf(p, B, self.cond_total)
        p.gc = gc
        hb, acts, outb = self._buf(p, B, H, T), self._buf(p, B, H, T), self._buf(p, B, H, T)
        vh, va, vo = V(hb, H, T), V(acts, H, T), V(outb, H, T)
        # speaker conditioning: cond_layer(g) x4 + dec.cond(g) in one launch (modules.py:39, synthesizers.py:147)
        if self.cond_table is None:   # (otherwise infer() gathers the rows of the per-speaker table straight into gc)
            p.add("cond", lib.rvc_cond_gemv_f32, p.g.data_ptr(), W["cond_w"].data_ptr(), W["cond_b"].data_ptr(), gc.data_ptr(), B, G, self.cond_total,
                  flops=2.0 * B * G * self.cond_total, bytes_=4.0 * G * self.cond_total, kind="cond_gemv_f32")
        fork_at, join_at = len(p.ops), None      # side branch (NSF source): hoisted to here at the end of the build (Plan.hoist_side)
        side = self.use_side_branch
        cluster = self.flow_tc and self.use_flow_cluster and (H, half) == (192, 96)     # (the flow's WaveNet is always 3 layers of k = 5, x 4 couplings)
        za = self._packed(p, B, I, T) if self.flow_tc else None       # masked packed z: conv_pre's operand
        if cluster:
            # ---- the whole reverse flow in one cluster kernel: reads z_p, writes z and its packed copy
            fd = FlowDesc()
            for f in range(4):
                cp = fd.flows[f]
                cp.pre_w, cp.post_w = W[f"t_fl{f}_pre_w"].data_ptr(), W[f"t_fl{f}_post_w"].data_ptr()
                cp.post_b, cp.sb, cp.gc_off = W[f"fl{f}_post_b"].data_ptr(), W[f"t_fl{f}_sb"].data_ptr(), self.cond_off[f]
                for l in range(3):
                    cp.in_w[l], cp.in_b[l], cp.hb[l] = W[f"t_fl{f}_in_w{l}"].data_ptr(), W[f"t_fl{f}_in_b{l}"].data_ptr(), W[f"t_fl{f}_hb{l}"].data_ptr()
                    cp.skip_w[l] = W[f"t_fl{f}_skip_w{l}"].data_ptr()
                    if l < 2:
                        cp.res_w[l] = W[f"t_fl{f}_res_w{l}"].data_ptr()
            fd.zp, fd.z, fd.za, fd.mask, fd.gc, fd.gc_bs = p.zp.data_ptr(), p.z.data_ptr(), za.data_ptr(), p.mask.data_ptr(), gc.data_ptr(), self.cond_total
            fd.B, fd.T, fd.hidden, fd.half, fd.n_flows, fd.n_layers, fd.kernel_size, fd.post_rows = B, T, H, half, 4, 3, 5, self.post_rows
            p.keep.append(fd)
            p.add("flow_cluster", lib.rvc_flow_reverse_tc, C.byref(fd), flops=4 * (2.0 * B * T * (half * H + 3 * H * 2 * H * 5 + 2 * H * 2 * H + H * H + H * half)),
                  bytes_=4.0 * B * I * T * 2 + 2.0 * B * I * T, kind="flow_cluster_tc")
        else:
            z_t, zp_t = p.z, p.zp   # (capture the tensors, not the plan: a closure over `p` would be a reference cycle that keeps an evicted plan's workspace alive)
            p.add_py(lambda: z_t.copy_(zp_t))
        chain = self.flow_tc and self.use_chain and H % 64 == 0 and half % 32 == 0
        for f in ((3, 2, 1, 0) if (self.flow_tc and not cluster) else ()):
            dt = DT_FP16
            if f == 3:
                x0a, hpa, acta, oa = self._packed(p, B, half, T), self._packed(p, B, H, T), self._packed(p, B, H, T), self._packed(p, B, H, T)
                # x0 of the first coupling = (flipped) second half of z_p, not masked (residuals.py:128,137)
                p.add("pack_x0", lib.rvc_pack_act, p.z.data_ptr() + half * T * 4, None, None, 1.0, I * T, T, None, x0a.data_ptr(), B, half, T, dt, 1.0,
                      kind="pack_act", bytes_=6.0 * B * half * T)
                if chain:
                    self._chain_begin()
            flipped = f in (3, 1)
            dst = V(p.z, I, T, 0 if flipped else half)
            self._tc(p, f"fl{f}_pre", [dict(a=x0a, w=W[f"t_fl{f}_pre_w"], bias=W[f"fl{f}_pre_b"], mask=p.mask, y=vh, y_act=hpa)],
                     B=B, Cin=half, Cout=H, L=T, dtype=dt, tag="flow")
            for l in range(3):
                self._tc(p, f"fl{f}_in{l}", [dict(a=hpa, w=W[f"t_fl{f}_in_w{l}"], bias=W[f"t_fl{f}_in_b{l}"], K=5, gated=True, y_act=acta,
                                                  bias_b=gc.data_ptr() + (self.cond_off[f] + l * 2 * H) * 4, bias_b_bs=self.cond_total)],
                         B=B, Cin=H, Cout=2 * H, L=T, dtype=dt, nt=64, tag="flow")
                if l < 2:
                    jobs = [dict(a=acta, w=W[f"t_fl{f}_res_w{l}"], bias=W[f"fl{f}_res_b{l}"], res=vh, mask=p.mask, y=vh, y_act=hpa),
                            dict(a=acta, w=W[f"t_fl{f}_skip_w{l}"], bias=W[f"fl{f}_skip_b{l}"], res=(vo if l > 0 else None), y=vo)]
                else:  # last layer: all channels are skip; output*mask feeds `post` directly in packed form (modules.py:49-51)
                    jobs = [dict(a=acta, w=W[f"t_fl{f}_skip_w{l}"], bias=W[f"fl{f}_skip_b{l}"], res=vo, mask=p.mask, y_act=oa)]
                self._tc(p, f"fl{f}_rs{l}", jobs, B=B, Cin=H, Cout=H, L=T, dtype=dt, tag="flow")
            # x1 <- (x1 - post(...)*mask)*mask; its packed copy is the next coupling's x0
            post = dict(a=oa, w=W[f"t_fl{f}_post_w"], bias=W[f"fl{f}_post_b"], res=dst, alpha=-1.0, mask=p.mask, y=dst, y_act=(x0a if f > 0 else None))
            if chain:
                self._tc(p, f"fl{f}_post", [post], B=B, Cin=H, Cout=self.post_rows, cout_valid=half, L=T, dtype=dt, tag="flow")
            else:
                self._tc(p, f"fl{f}_post", [dict(post, w_cs=self.post_rows)], B=B, Cin=H, Cout=half, L=T, dtype=dt, tag="flow")
            if chain and f == 0:
                self._chain_end(p, "flow_chain", "flow")
        for f in (() if self.flow_tc else (3, 2, 1, 0)):
            flipped = f in (3, 1)
            src = V(p.z, I, T, half if flipped else 0)
            dst = V(p.z, I, T, 0 if flipped else half)
            self._conv(p, f"fl{f}_pre", src, W[f"fl{f}_pre_w"], vh, B=B, Cin=half, Cout=H, L=T, bias=W[f"fl{f}_pre_b"], mask=p.mask, out_mask=True)
            for l in range(3):
                self._conv(p, f"fl{f}_in{l}", vh, W[f"fl{f}_in_w{l}"], va, B=B, Cin=H, Cout=H, L=T, K=5, pad=2, bias=W[f"fl{f}_in_b{l}"],
                           bias_b=gc.data_ptr() + (self.cond_off[f] + l * 2 * H) * 4, bias_b_bs=self.cond_total, gated=True)
                if l < 2:
                    self._conv(p, f"fl{f}_res{l}", va, W[f"fl{f}_res_w{l}"], vh, B=B, Cin=H, Cout=H, L=T, bias=W[f"fl{f}_res_b{l}"], res=vh,
                               mask=p.mask, out_mask=True)
                self._conv(p, f"fl{f}_skip{l}", va, W[f"fl{f}_skip_w{l}"], vo, B=B, Cin=H, Cout=H, L=T, bias=W[f"fl{f}_skip_b{l}"],
                           beta=0.0 if l == 0 else 1.0)
            self._conv(p, f"fl{f}_post", vo, W[f"fl{f}_post_w"], dst, B=B, Cin=H, Cout=half, L=T, bias=W[f"fl{f}_post_b"], mask=p.mask,
                       in_mask=True, out_mask=True, res=dst, alpha=-1.0)

        # ---- decoder
        Lfin = T * upp
        if c["f0"]:
            p.f0 = self._buf(p, B, T)
            p.eps_src = self._buf(p, B, Lfin)
            har = self._buf(p, B, Lfin)
            scratch = self._buf(p, B, T)
            p.add("sine_source", lib.rvc_sine_source_f32, p.f0.data_ptr(), p.eps_src.data_ptr(), W["lin_w"].data_ptr(), W["lin_b"].data_ptr(),
                  scratch.data_ptr(), har.data_ptr(), B, T, upp, c["sr"], launches=2, bytes_=8.0 * B * Lfin, kind="sine_source_f32", side=side)
        xpre = self._buf(p, B, U, T)
        xpre_a = None
        if self.flow_tc:
            if not cluster:
                p.add("pack_z", lib.rvc_pack_act, p.z.data_ptr(), None, None, 1.0, I * T, T, p.mask.data_ptr(), za.data_ptr(), B, I, T, DT_FP16, 1.0, kind="pack_act",
                      bytes_=6.0 * B * I * T)
            xpre_a = self._packed(p, B, U, T) if self.up_dt[0] == DT_FP16 else None
            self._tc(p, "conv_pre", [dict(a=za, w=W["t_pre_w"], bias=W["pre_b"], K=7, y=V(xpre, U, T), y_act=xpre_a, act_slope=LRELU_SLOPE,
                                          bias_b=gc.data_ptr() + self.cond_off[4] * 4, bias_b_bs=self.cond_total)],
                     B=B, Cin=I, Cout=U, L=T, dtype=DT_FP16, tag="conv_pre")
        else:
            self._conv(p, "conv_pre", V(p.z, I, T), W["pre_w"], V(xpre, U, T), B=B, Cin=I, Cout=U, L=T, K=7, pad=3, bias=W["pre_b"],
                       bias_b=gc.data_ptr() + self.cond_off[4] * 4, bias_b_bs=self.cond_total, mask=p.mask, in_mask=True)
        rates, uks, rks, rds = c["rates"], c["uks"], c["rks"], c["rds"]
        nk = len(rks)
        chans = [U // (2 ** (i + 1)) for i in range(len(rates))]
        lens, L = [], T
        for u in rates:
            L *= u
            lens.append(L)
        big = max(ch * ln for ch, ln in zip(chans, lens))
        S = [self._buf(p, B * big) for _ in range(4)]
        cur, curC, curL = (xpre, None, None), U, T   # cur = (x, x2, x3): the stage input is (x + x2 + x3) / nk when x2 is set
        prev_ua = None
        for i, (u, k) in enumerate(zip(rates, uks)):
            Ci, Li = chans[i], lens[i]
            dt = self.stage_dt[i]
            vs0 = V(S[0], Ci, Li)
            sn = int(math.prod(rates[i + 1:])) if i + 1 < len(rates) else 1
            kn = 1 if sn == 1 else sn * 2 - sn % 2
            pn = 0 if sn == 1 else (kn - sn) // 2
            xa = None   # packed lrelu(x) of this stage's input to the ResBlocks, produced by the ups epilogue when possible
            # a stage that runs entirely on the tensor-core kernel keeps its fp32 residual streams (S[0], P[j]) channel-blocked,
            # [b][c/8][l][8]: the epilogues then move whole 32-byte sectors per thread (rvc_conv1d_tc_desc.f32_blocked)
            blk = self.up_dt[i] is not None and dt is not None and dt == self.up_dt[i] and Ci % 8 == 0
            if blk:
                vs0 = V(S[0], Ci, Li, cs=8 * Li)
            split, x_lo = False, None
            if self.up_dt[i] is not None:
                # ---- ConvTranspose1d on tensor cores: u phase GEMMs; the noise conv result is the epilogue's residual input
                udt = self.up_dt[i]
                if i == 0 and xpre_a is not None:
                    ua = xpre_a
                elif prev_ua is not None:
                    ua = prev_ua          # written by the previous stage's fused last conv2
                else:
                    ua = self._packed(p, B, curC, curL)
                    x2 = cur[1].data_ptr() if cur[1] is not None else None
                    x3 = cur[2].data_ptr() if cur[2] is not None else None
                    p.add(f"pack_up{i}", lib.rvc_pack_act, cur[0].data_ptr(), x2, x3, float(nk), curC * curL, curL, None, ua.data_ptr(), B, curC,
                          curL, udt, LRELU_SLOPE, bytes_=B * curC * curL * (2.0 + (12.0 if x2 else 4.0)), kind="pack_act")
                # short noise convs (kernel <= 4: the last two stages) are computed inside the ups epilogue instead of a launch of
                # their own followed by an fp32 round trip (rvc_tc_job.nz_*); the long ones keep their kernel (tried in round 2: the 8-tap
                # conv of stage 1 inside the issue-bound ups epilogue cost 55 us against 29 us saved)
                nz_fused = bool(c["f0"]) and blk and kn <= 4
                nzv = vs0                       # where the (unfused) noise conv leaves its result = the residual input of this stage's ups
                if c["f0"] and not nz_fused:
                    if side:
                        # on the side branch the conv runs long before its stage: the first one may use S[0] (nothing touches it before ups0), later ones
                        # need a buffer of their own
                        if join_at is None:
                            join_at = len(p.ops)
                        else:
                            nzv = V(self._buf(p, B * Ci * Li), Ci, Li, cs=(8 * Li if blk else None))
                    p.add(f"noise{i}", lib.rvc_noise_conv_f32, har.data_ptr(), Lfin, W[f"nz_w{i}"].data_ptr(), W[f"nz_b{i}"].data_ptr(),
                          nzv.ptr, Ci * Li, B, Ci, Li, kn, sn, pn, int(blk), flops=2.0 * B * Ci * Li * kn, bytes_=4.0 * B * (Ci * Li + Li * sn),
                          kind="noise_conv_f32", side=side)
                if dt is not None and dt == udt:
                    xa = self._packed(p, B, Ci, Li)
                # narrow fused stages: the stage input leaves the ups as the split-fp16 stream (xa = hi, x_lo), no fp32 copy
                split = blk and dt == DT_FP16 and xa is not None and (
                    (Ci in (32, 64) and self.pair_fusion and self.pair_split) or (not (Ci in (32, 64) and self.pair_fusion) and self.tc_split))
                x_lo = self._buf(p, xa.numel(), dtype=torch.uint8) if split else None
                M = -(-k // u)
                nz = dict(nz_har=har.data_ptr(), nz_w=W[f"nz_w{i}"].data_ptr(), nz_b=W[f"nz_b{i}"].data_ptr(), nz_har_bs=Lfin, nz_kn=kn, nz_sn=sn,
                          nz_pn=pn) if nz_fused else {}
                merged = f"t_up_wm{i}" in W and (nz_fused or not c["f0"])      # (phase columns take no residual input)
                if merged:
                    self._tc(p, f"ups{i}", [dict(a=ua, w=W[f"t_up_wm{i}"], bias=W[f"up_bm{i}"], y=(None if split else vs0), y_act=xa, y_lo=x_lo,
                                                 act_slope=LRELU_SLOPE, K=M, pad=M - 1, **nz)],
                             B=B, Cin=curC, Cout=u * Ci, L=curL, dtype=udt, up_cols=u, pshift=(k - u) // 2, tag=f"ups",
                             flops=2.0 * B * curC * Ci * k * curL, blocked=3 if blk else 0)
                else:
                    self._tc(p, f"ups{i}", [dict(a=ua, w=W[f"t_up_w{i}"], bias=W[f"up_b{i}"], res=(nzv if c["f0"] and not nz_fused else None),
                                                 y=(None if split else vs0), y_act=xa, y_lo=x_lo, act_slope=LRELU_SLOPE, K=M, pad=M - 1, **nz)],
                             B=B, Cin=curC, Cout=Ci, L=curL, dtype=udt, up=u, pshift=(k - u) // 2, tag=f"ups",
                             flops=2.0 * B * curC * Ci * k * curL, blocked=3 if blk else 0)
            else:
                d = ConvTr1dDesc()
                vx = V(cur[0], curC, curL)
                d.x, d.x_bs, d.x_cs = vx.ptr, vx.bs, vx.cs
                if cur[1] is not None:
                    d.x2, d.x3, d.in_div = cur[1].data_ptr(), cur[2].data_ptr(), float(nk)
                d.w, d.bias = W[f"up_w{i}"].data_ptr(), W[f"up_b{i}"].data_ptr()
                if c["f0"]:
                    d.har, d.har_bs = har.data_ptr(), Lfin
                    d.nw, d.nbias = W[f"nz_w{i}"].data_ptr(), W[f"nz_b{i}"].data_ptr()
                    d.kn, d.sn, d.pn = kn, sn, pn
                d.y, d.y_bs, d.y_cs = vs0.ptr, vs0.bs, vs0.cs
                d.B, d.Cin, d.Cout, d.Lin, d.k, d.u = B, curC, Ci, curL, k, u
                d.pre_slope = LRELU_SLOPE
                p.keep.append(d)
                p.add(f"ups{i}", lib.rvc_convtr1d_f32, C.byref(d), flops=2.0 * B * curC * Ci * k * curL, bytes_=4.0 * B * (curC * curL + Ci * Li),
                      kind="convtr1d_f32")
            if dt is None:
                # ---- fp32 CUDA-core ResBlocks, sequential, accumulated into S[3]
                vs1, vs2, vs3 = V(S[1], Ci, Li), V(S[2], Ci, Li), V(S[3], Ci, Li)
                for j, (kk, dils) in enumerate(zip(rks, rds)):
                    src = vs0
                    for mm, dl in enumerate(dils):
                        last = mm == len(dils) - 1
                        self._conv(p, f"rb{i}_{j}_c1_{mm}", src, W[f"rb{i}_{j}_c1w{mm}"], vs1, B=B, Cin=Ci, Cout=Ci, L=Li, K=kk, dil=dl,
                                   pad=(kk * dl - dl) // 2, bias=W[f"rb{i}_{j}_c1b{mm}"], pre_slope=LRELU_SLOPE)
                        if not last:
                            self._conv(p, f"rb{i}_{j}_c2_{mm}", vs1, W[f"rb{i}_{j}_c2w{mm}"], vs2, B=B, Cin=Ci, Cout=Ci, L=Li, K=kk,
                                       pad=(kk - 1) // 2, bias=W[f"rb{i}_{j}_c2b{mm}"], pre_slope=LRELU_SLOPE, res=src)
                            src = vs2
                        else:
                            self._conv(p, f"rb{i}_{j}_c2_{mm}", vs1, W[f"rb{i}_{j}_c2w{mm}"], vs3, B=B, Cin=Ci, Cout=Ci, L=Li, K=kk,
                                       pad=(kk - 1) // 2, bias=W[f"rb{i}_{j}_c2b{mm}"], pre_slope=LRELU_SLOPE, res=src,
                                       out_scale=1.0 / nk, beta=0.0 if j == 0 else 1.0)
                cur, curC, curL = (S[3], None, None), Ci, Li
                prev_ua = None
            else:
                # ---- tensor-core ResBlocks: the nk blocks advance in lock-step, one persistent launch per conv step
                esz = 2
                rows = int(lib.rvc_packed_rows(Li))
                pk = lambda: self._zbuf(p, B * Ci * rows * esz)          # packed activations (pad rows stay zero)
                ta, pa = [pk() for _ in range(nk)], [pk() for _ in range(nk)]
                P = [S[1 + j] for j in range(nk)]                        # fp32 residual streams / block outputs
                if xa is None:
                    xa = pk()
                    p.add(f"pack{i}", lib.rvc_pack_act, S[0].data_ptr(), None, None, 1.0, Ci * Li, Li, None, xa.data_ptr(), B, Ci, Li, dt,
                          LRELU_SLOPE, bytes_=B * Ci * Li * (4.0 + esz), kind="pack_act")
                order = sorted(range(nk), key=lambda j: -rks[j])          # heavy kernels first (static round-robin tiles)
                nxt = i + 1
                next_tc_ups = nxt < len(rates) and self.up_dt[nxt] is not None and self.up_dt[nxt] == dt
                next_ua = self._packed(p, B, Ci, Li) if next_tc_ups else None   # packed lrelu(stage average) for the next ups
                need_f32 = not next_tc_ups                                         # conv_post / fp32 ups read the fp32 average
                # blocked residuals -> plain [B, C, L] average cannot be written in place over P[0]
                avg = (self._buf(p, B * Ci * Li) if blk else S[1]) if need_f32 else None
                nsteps = len(rds[0])
                assert all(len(dl) == nsteps for dl in rds)
                # narrow, fully blocked stages: one fused launch per ResBlock step (conv1 -> lrelu -> conv2 + x, the intermediate
                # stays in shared memory, rvc_resblock_pair_tc) for every step but the last, whose conv2 is the fused stage average
                use_pair = blk and Ci in (32, 64) and self.pair_fusion
                src_a = {j: xa for j in range(nk)}     # packed lrelu(x_j) read by the next conv1
                # split-fp16 stream: lo plane of each block's stream (step 0 reads the shared x_lo; later steps update lo[j] in place)
                lo_j = [self._buf(p, xa.numel(), dtype=torch.uint8) for _ in range(nk)] if split else None
                t_of, rhi_of = {}, {}
                for mm in range(nsteps):
                    last = mm == nsteps - 1
                    if use_pair:
                        # the last step runs in sum mode: the three conv2 accumulate into one tile and the stage average is written
                        # once — plain fp32 for conv_post / an fp32 ups and (or) packed for the next tensor-core ups
                        pd = ResblockPairDesc()
                        pd.njobs, pd.B, pd.C, pd.L, pd.dtype, pd.slope = nk, B, Ci, Li, dt, LRELU_SLOPE
                        pd.sum_jobs, pd.y_ncl, pd.alpha = int(last), int(last), 1.0 / nk
                        fl = 0.0
                        for slot, j in enumerate(order):
                            pj = pd.jobs[slot]
                            kk = rks[j]
                            pj.a = src_a[j].data_ptr()
                            pj.w1, pj.b1 = W[f"rb{i}_{j}_c1w{mm}"].data_ptr(), W[f"rb{i}_{j}_c1b{mm}"].data_ptr()
                            pj.w2, pj.b2 = W[f"rb{i}_{j}_c2w{mm}"].data_ptr(), W[f"rb{i}_{j}_c2b{mm}"].data_ptr()
                            if split:
                                pj.a_lo = (x_lo if mm == 0 else lo_j[j]).data_ptr()
                            else:
                                pj.res = (S[0] if mm == 0 else P[j]).data_ptr()
                            pj.K1, pj.dil1, pj.K2 = kk, rds[j][mm], kk
                            if not last:
                                dst_a = pa[j] if src_a[j] is not pa[j] else ta[j]     # ping-pong: tiles read `a` with halos while others write y_act
                                pj.y_act = dst_a.data_ptr()
                                if split:
                                    pj.y_lo = lo_j[j].data_ptr()
                                else:
                                    pj.y = P[j].data_ptr()
                                src_a[j] = dst_a
                            elif slot == 0:
                                pj.y = avg.data_ptr() if need_f32 else None
                                pj.y_act = next_ua.data_ptr() if next_ua is not None else None
                            fl += 4.0 * B * Ci * Ci * kk * Li
                        p.keep.append(pd)
                        rbytes = 2.0 * esz if split else 4.0      # residual stream: hi (the operand itself, re-read through L2) + lo | fp32
                        by = nk * B * Ci * Li * (esz + 2 * rbytes - (esz if split else 0.0) + (0.0 if split else esz)) if not last else \
                            B * Ci * Li * (nk * (esz + rbytes - (esz if split else 0.0)) + (4.0 if need_f32 else 0.0) + (esz if next_ua is not None else 0.0))
                        p.add(f"rbpair{i}_s{mm}", lib.rvc_resblock_pair_tc, C.byref(pd), flops=fl, bytes_=by,
                              kind=f"resblock_pair_tc_{'bf16' if dt == DT_BF16 else 'fp16'}[C{Ci}]")
                        continue
                    for half in (1, 2):
                        d = Conv1dTcDesc()
                        d.njobs, d.B, d.Cin, d.Cout, d.L, d.dtype = nk, B, Ci, Ci, Li, dt
                        # last conv2 of the stage: all three blocks accumulate into one tile and the stage average
                        # (x0' + x1' + x2') / nk is written once — fp32 for conv_post / an fp32 ups, and (or) directly as
                        # the packed leaky_relu input of the next tensor-core ups (synthesizers.py:150,157-159)
                        fused = last and half == 2
                        d.sum_jobs = int(fused)
                        rcs = 8 * Li if blk else Li     # channel(-block) stride of the fp32 streams
                        # the fused stage average leaves the blocked domain: conv_post / an fp32 ups read plain [B, C, L]
                        d.f32_blocked = (1 if fused else 3) if (blk and half == 2) else 0
                        fl = 0.0
                        for slot, j in enumerate(order):
                            jb = d.jobs[slot]
                            kk = rks[j]
                            if half == 1:   # t = conv1_d(lrelu(x));  store lrelu(t) packed
                                tdst = ta[j] if src_a[j] is not ta[j] else pa[j]
                                jb.a = src_a[j].data_ptr()
                                jb.w, jb.bias = W[f"rb{i}_{j}_c1w{mm}"].data_ptr(), W[f"rb{i}_{j}_c1b{mm}"].data_ptr()
                                jb.y_act = tdst.data_ptr()
                                t_of[j], rhi_of[j] = tdst, src_a[j]      # (the operand of conv1 is the hi plane of this step's residual)
                                jb.K, jb.dil, jb.pad = kk, rds[j][mm], (kk - 1) // 2 * rds[j][mm]
                            else:           # x = conv2(lrelu(t)) + x;  store x (fp32) and lrelu(x) packed for the next pair
                                jb.a = t_of[j].data_ptr()
                                jb.w, jb.bias = W[f"rb{i}_{j}_c2w{mm}"].data_ptr(), W[f"rb{i}_{j}_c2b{mm}"].data_ptr()
                                if split:     # x = inv_lrelu(hi + lo); x' leaves the same way (the lo plane is updated in place)
                                    jb.res_hi, jb.res_lo = rhi_of[j].data_ptr(), (x_lo if mm == 0 else lo_j[j]).data_ptr()
                                else:
                                    jb.res, jb.res_bs, jb.res_cs = (S[0] if mm == 0 else P[j]).data_ptr(), Ci * Li, rcs
                                if not fused:
                                    if split:
                                        jb.y_lo = lo_j[j].data_ptr()
                                    else:
                                        jb.y_f32, jb.y_bs, jb.y_cs = P[j].data_ptr(), Ci * Li, rcs
                                    ndst = pa[j] if t_of[j] is not pa[j] else ta[j]    # (conv1 of this step has consumed src_a[j])
                                    jb.y_act = ndst.data_ptr()
                                    src_a[j] = ndst
                                elif slot == 0:
                                    if need_f32:   # NCL: into P[0]'s buffer (every thread reads its residuals before it writes)
                                        jb.y_f32, jb.y_bs, jb.y_cs = avg.data_ptr(), Ci * Li, Li
                                    if next_ua is not None:
                                        jb.y_act = next_ua.data_ptr()
                                jb.K, jb.dil, jb.pad = kk, 1, (kk - 1) // 2
                            jb.act_slope, jb.alpha = LRELU_SLOPE, (1.0 / nk if fused else 1.0)
                            fl += 2.0 * B * Ci * Ci * kk * Li
                        p.keep.append(d)
                        rb = 2.0 * esz if split else 4.0       # residual bytes per element: hi + lo | fp32
                        by = nk * B * Ci * Li * (2.0 * esz if half == 1 else (esz + 2 * rb + (0.0 if split else esz))) if not fused else \
                            B * Ci * Li * (nk * (esz + rb) + (4.0 if need_f32 else 0.0) + (esz if next_ua is not None else 0.0))
                        p.add(f"rbtc{i}_s{mm}_c{half}", lib.rvc_conv1d_tc, C.byref(d), flops=fl, bytes_=by,
                              kind=f"conv1d_tc_{'bf16' if dt == DT_BF16 else 'fp16'}[C{Ci}]")
                cur, curC, curL = (avg, None, None), Ci, Li       # stage average (fp32, plain layout), None when the next ups is fused
                prev_ua = next_ua
        x2 = cur[1].data_ptr() if cur[1] is not None else None
        x3 = cur[2].data_ptr() if cur[2] is not None else None
        p.add("conv_post", lib.rvc_conv_post_f32, cur[0].data_ptr(), x2, x3, float(nk), curC * curL, curL, W["post_w"].data_ptr(), p.o.data_ptr(),
              Lfin, B, curC, curL, 0.01, flops=2.0 * B * curC * 7 * curL, bytes_=4.0 * B * (curC * (3 if x2 else 1) + 1) * curL,
              kind="conv_post_f32")
        p.hoist_side(fork_at, join_at, self.side_stream, self.dev)
        return p

    def plan(self, kind, B, T, slot=0):
        key = (kind, B, T) if slot == 0 else (kind, B, T, slot)     # (slot > 0: a second set of workspaces of the same shape, for infer_lanes)
        pl = self.plans.pop(key, None)
        cur = torch.cuda.current_stream(self.dev) if self.dev.type == "cuda" else None
        if pl is None:
            if len(self.plans) >= self.max_plans:  # bound the workspace arena: drop the least recently used shape
                old = next(iter(self.plans))
                # ... and the graphs that replay into its buffers (before the buffers go); graphs of other shapes stay
                for gk in [gk for gk, (_, pkeys) in self.graphs.items() if old in pkeys]:
                    del self.graphs[gk]
                    self._warm.discard(gk)
                self.plans.pop(old).release()
            pl = self._enc_plan(B, T) if kind == "enc" else self._dec_plan(B, T)
        elif pl.last_stream is not None and pl.last_stream != cur:
            cur.wait_stream(pl.last_stream)   # workspaces are shared per shape: order this use behind the previous one
        pl.last_stream = cur
        self.plans[key] = pl   # most recently used last
        return pl

    # ------------------------------------------------------------------ two lanes: encoder of chunk i+1 next to the generator of chunk i
    def _lane_graph(self, gkey, pkeys, body):
        """Replay (or, on the second use of a shape, capture) one lane's kernel sequence as its own CUDA graph on the current stream."""
        ent = self.graphs.pop(gkey, None) if self.use_graph else None
        gr = ent[0] if ent is not None else None
        if self.use_graph and gr is None and gkey in self._warm:
            if len(self.graphs) >= 2 * self.max_plans:
                self.graphs.pop(next(iter(self.graphs)))
            gr = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.dev)
            with torch.cuda.graph(gr):
                body(torch.cuda.current_stream(self.dev).cuda_stream)
            ent = (gr, pkeys)
        if ent is not None:
            self.graphs[gkey] = ent
        if gr is not None:
            gr.replay()
        else:
            body(torch.cuda.current_stream(self.dev).cuda_stream)
            self._warm.add(gkey)

    @staticmethod
    def _front_ops(pd):
        """How many leading ops of a generator plan belong to the FRONT lane: the reverse flow (everything ahead of the NSF source / conv_pre)."""
        if pd.side is not None:         # (a plan with a side branch keeps its fork / join ops together: everything stays in the back lane)
            return 0
        for i, op in enumerate(pd.ops):
            if op[0] in ("sine_source", "conv_pre"):
                return i
        return 0

    GROUP_MAX = 8

    def group_size(self, B, T):
        """How many consecutive items of shape (B, T) ``infer_lanes`` runs as one batch: as many as fit RVC_B200_GROUP_FRAMES frames (default 3000,
        i.e. three 10 s chunks; 0 = every item on its own), at most GROUP_MAX.  Measured on the 10 s clip of configs[1] (bench.py, ms per chunk):
        1 chunk per pass 2.06, 2: 1.88, 3: 1.83 - 1.89, 4: 1.85 - 1.91, 6: 1.92, 8: 1.93 (larger batches run into the power cap)."""
        return max(1, min(self.GROUP_MAX, int(os.environ.get("RVC_B200_GROUP_FRAMES", "3000")) // max(1, B * T)))

    def infer_lanes(self, items, out_dtype):
        """Pipelined ``infer`` over an iterable of ``(phone, phone_lengths, pitch, nsff0, sid[, rate[, noise]])``: yields the result tuples in order.

        The chunks of a conversion are independent (main/inference/convert.py:442-447 is a serial loop over slices), but inside one chunk the
        TextEncoder and the reverse flow are chains of short launches that fill a third of the GPU at best, and only then does the generator
        start.  Here the two halves run in two lanes — streams — of their own: while the generator of chunk i runs in the back lane, the encoder
        and flow of chunk i+1 run next to it in the front lane; the hardware places the front lane's CTAs as the generator's persistent kernels
        hand SMs over at their launch boundaries (measured: 2.25 -> 2.06 ms per 10 s chunk; capping the generator's grids to reserve SMs for
        the other lane, giving either lane priority, or dealing the generator's tiles dynamically was no faster: tools/overlap_probe.py).
        The generator plan exists twice (slots i % 2: chunk i+1's flow writes z while chunk i's generator still reads its own); the front lane
        of chunk i+2 waits for the back lane of chunk i.  Consecutive short items of one shape are batched (``group_size``: three 10 s chunks per pass —
        the encoder's chains of launches take as long for two chunks as for one); every item keeps its own inputs, speaker and noise draws.  Same
        kernels and RNG draw order as ``infer`` item by item; a group's results equal ``infer`` on the stacked batch bit for bit, and ``infer`` item by
        item up to the summation order of the tensor-core generator, whose tile shapes follow the batch (rvc_conv1d_tc; ~2e-4 on the waveform; the
        encoder and flow outputs are bit-identical).  With RVC_B200_GROUP_FRAMES=0 nothing is batched and everything is bit-identical
        (tests/test_gpu_parity.py).  The results of a group are handed out once the NEXT group has been enqueued, or at the end.
        """
        c = self.cfg
        dev = self.dev
        if getattr(self, "_lanes", None) is None:
            self._lanes = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        F0, D = self._lanes
        # with many frames per call (batched chunks, long chunks) the encoder is a small share and fills the GPU on its own: a second lane then only
        # costs interference (measured: B = 8 x T = 1000 -3 %, B = 2 x T = 4000 +-0, B = 1 x T = 1000 +9 %); such items run both halves in one lane
        lane_max = int(os.environ.get("RVC_B200_LANE_MAX_FRAMES", "4000"))
        caller = torch.cuda.current_stream(dev)
        done = [None, None]   # per slot: event "the back lane has finished the item that used this slot's generator plan"
        used = []
        pending = None
        n_item = 0

        def finish(h):
            """the result tuples of the items of one group: views of the group's (fresh) result arena, one slice of the batch per item"""
            res, stats, pd_shape, ev, offsets, G = h
            caller.wait_event(ev)
            res.record_stream(caller)
            stats.record_stream(caller)
            GB, T2 = pd_shape
            B = GB // G
            I = c["I"]
            n_o, n_m, n_z = offsets
            Lfin = T2 * c["upp"]
            o = res[:GB * Lfin].view(GB, 1, Lfin)
            x_mask = res[n_o:n_o + GB * T2].view(GB, 1, T2)
            z = res[n_o + n_m:n_o + n_m + GB * I * T2].view(GB, I, T2)
            z_p = res[n_o + n_m + n_z:n_o + n_m + n_z + GB * I * T2].view(GB, I, T2)
            for i in range(G):
                sl = slice(i * B, (i + 1) * B)
                yield o[sl], x_mask[sl], (z[sl], z_p[sl], stats[sl, :I], stats[sl, I:])

        def signature(item):
            phone, lengths, pitch, nsff0, sid = item[:5]
            if phone.dim() != 3 or phone.shape[2] != c["D"]:
                raise ValueError(f"phone must be [B, T, {c['D']}], got {tuple(phone.shape)}")
            if c["f0"] and (pitch is None or nsff0 is None):
                raise ValueError("use_f0 model: pitch and nsff0 are required")
            if sid is None:
                raise ValueError("sid is required")
            rate = item[5] if len(item) > 5 else None
            noise = item[6] if len(item) > 6 else None
            return (int(phone.shape[0]), int(phone.shape[1]), rate, noise is None)

        def grouped(it):
            """consecutive items of one shape, as many as fit ``group_frames`` frames: they run as ONE batch (a chunk's encoder and flow are chains of
            launches that take as long for two short chunks as for one), every item with its own inputs, speaker and noise draws"""
            grp, sig = [], None
            for item in it:
                item = tuple(item)
                s_ = signature(item)
                if grp and (s_ != sig or len(grp) >= gmax):
                    yield grp
                    grp = []
                if not grp:
                    sig = s_
                    gmax = self.group_size(s_[0], s_[1])
                grp.append(item)
            if grp:
                yield grp

        try:
            for grp in grouped(items):
                G = len(grp)
                Bi, T = int(grp[0][0].shape[0]), int(grp[0][0].shape[1])
                B = G * Bi
                rate = grp[0][5] if len(grp[0]) > 5 else None
                head = int(T * (1.0 - rate)) if rate is not None else 0
                T2 = T - head
                slot = n_item % 2
                n_item += 1
                F = F0 if B * T < lane_max else D
                dkey = ("dec", B, T2) if slot == 0 else ("dec", B, T2, slot)
                rows = [slice(i * Bi, (i + 1) * Bi) for i in range(G)]
                with self._lock, torch.cuda.device(dev):
                    if len(self.plans) >= self.max_plans - 1 and (("enc", B, T) not in self.plans or dkey not in self.plans):
                        torch.cuda.synchronize(dev)      # a plan is about to be evicted: nothing of either lane may still be using its buffers
                    F.wait_stream(caller)                # the items' tensors may have been produced on the caller's stream
                    D.wait_stream(caller)
                    with torch.cuda.stream(F):
                        pe = self.plan("enc", B, T)
                        pd = self.plan("dec", B, T2, slot)
                        used[:] = [pl for pl in used if pl is not pe and pl is not pd] + [pe, pd]
                        if done[slot] is not None:
                            F.wait_event(done[slot])     # this slot's generator plan is free again
                    # inputs and noise item by item, in the order ``infer`` called item by item would draw: randn_like(m_p) (front lane), then
                    # rand_ini and randn_like(sine_waves) (back lane: the back lane is in order, so its buffers are free when the draws run)
                    for item, r in zip(grp, rows):
                        phone, phone_lengths, pitch, nsff0, sid = item[:5]
                        noise = item[6] if len(item) > 6 else None
                        with torch.cuda.stream(F):
                            pe.phone[r].copy_(phone, non_blocking=True)
                            pe.lengths[r].copy_(phone_lengths, non_blocking=True)
                            if c["f0"]:
                                pe.pitch[r].copy_(pitch, non_blocking=True)
                            if noise is not None:
                                pe.eps_z[r].copy_(noise[0], non_blocking=True)
                            else:
                                pe.eps_z[r].normal_()        # randn_like(m_p), synthesizers.py:449
                            # (the flow's WaveNets are conditioned on the speaker: cond_layer(emb_g(sid)) belongs to the front lane)
                            if self.cond_table is not None:
                                pd.gc[r].copy_(self.cond_table[sid.to(dev, non_blocking=True)])
                            else:
                                pd.g[r].copy_(self.W["emb_g"][sid.to(dev, non_blocking=True)])
                            for t in (phone, phone_lengths, pitch, sid):
                                if t is not None and t.is_cuda:
                                    t.record_stream(F)
                        if c["f0"]:
                            with torch.cuda.stream(D):
                                pd.f0[r].copy_(nsff0[:, head:], non_blocking=True)
                                if noise is not None:
                                    pd.eps_src[r].copy_(noise[1].reshape(noise[1].shape[0], -1)[:, : T2 * c["upp"]], non_blocking=True)
                                else:
                                    torch.rand(1, 1, 1, device=dev)   # rand_ini (:87): drawn to keep the generator stream aligned, value unused (:88)
                                    pd.eps_src[r].normal_()           # randn_like(sine_waves), :98
                                if nsff0.is_cuda:
                                    nsff0.record_stream(D)
                    with torch.cuda.stream(F):
                        kf = self._front_ops(pd)

                        def front(st, pe=pe, pd=pd, head=head, kf=kf):
                            pe.run(st)
                            pd.zp.copy_(pe.zp[:, :, head:])
                            pd.mask.copy_(pe.mask[:, head:])
                            pd.run(st, 0, kf)

                        self._lane_graph(("lane-front", B, T, head, slot), (("enc", B, T), dkey), front)
                        stats = pe.stats.to(out_dtype, copy=True)
                        ev_front = torch.cuda.Event()
                        ev_front.record(F)
                    with torch.cuda.stream(D):
                        D.wait_event(ev_front)
                        self._lane_graph(("lane-back", B, T2, slot), (dkey,), lambda st, pd=pd, kf=kf: pd.run(st, kf, None))
                        res = pd.out_arena.to(out_dtype, copy=True)
                        ev_done = torch.cuda.Event()
                        ev_done.record(D)
                        done[slot] = ev_done
                    cur = (res, stats, (B, T2), ev_done, pd.out_offsets, G)
                if pending is not None:
                    yield from finish(pending)
                pending = cur
            if pending is not None:
                last, pending = pending, None
                yield from finish(last)
        finally:
            # whatever happens to this generator, later users of these plans (on any stream) queue behind both lanes
            caller.wait_stream(F0)
            caller.wait_stream(D)
            for pl in used:
                pl.last_stream = caller

    def profile(self, B, T, repeats=3):
        """Per-kernel CUDA-event timings of the (B, T) plans, replayed on the workspaces left by the last infer()."""
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        best = None
        with torch.cuda.device(self.dev):
            for _ in range(repeats):
                rows = self.plan("enc", B, T).run_timed(stream) + self.plan("dec", B, T).run_timed(stream)
                if best is None:
                    best = rows
                else:
                    for a, b in zip(best, rows):
                        a["ms"] = min(a["ms"], b["ms"])
        return best

    def launches_per_infer(self, B, T):
        return self.plan("enc", B, T).launches + self.plan("dec", B, T).launches

    # ------------------------------------------------------------------ infer
    def infer(self, phone, phone_lengths, pitch, nsff0, sid, rate, out_dtype, noise=None):
        c = self.cfg
        if phone.dim() != 3 or phone.shape[2] != c["D"]:
            raise ValueError(f"phone must be [B, T, {c['D']}], got {tuple(phone.shape)}")
        B, T = int(phone.shape[0]), int(phone.shape[1])
        if c["f0"] and (pitch is None or nsff0 is None):
            raise ValueError("use_f0 model: pitch and nsff0 are required")
        if sid is None:
            raise ValueError("sid is required")
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        # one infer at a time per engine (the reference's infer is stateless; here inputs, workspaces and graphs are persistent
        # per shape), and a shape used on a new stream waits for its previous use (Engine.plan)
        with self._lock, torch.cuda.device(self.dev):
            pe = self.plan("enc", B, T)
            pe.phone.copy_(phone, non_blocking=True)
            pe.lengths.copy_(phone_lengths, non_blocking=True)
            if c["f0"]:
                pe.pitch.copy_(pitch, non_blocking=True)
            if noise is not None:
                pe.eps_z.copy_(noise[0], non_blocking=True)
            else:
                pe.eps_z.normal_()  # randn_like(m_p), synthesizers.py:449
            head = int(T * (1.0 - rate)) if rate is not None else 0
            T2 = T - head
            pd = self.plan("dec", B, T2)
            if self.cond_table is not None:   # cond_layer(emb_g(sid)) x 4 + dec.cond(emb_g(sid)), precomputed per speaker at load time
                pd.gc.copy_(self.cond_table[sid.to(self.dev, non_blocking=True)])
            else:
                pd.g.copy_(self.W["emb_g"][sid.to(self.dev, non_blocking=True)])  # emb_g(sid), :447 (no host sync for a pinned sid)
            if c["f0"]:
                pd.f0.copy_(nsff0[:, head:], non_blocking=True)
                if noise is not None:
                    pd.eps_src.copy_(noise[1].reshape(noise[1].shape[0], -1)[:, : T2 * c["upp"]], non_blocking=True)
                else:
                    torch.rand(1, 1, 1, device=self.dev)  # rand_ini (:87) — drawn to keep the generator stream aligned, value unused (:88)
                    pd.eps_src.normal_()                  # randn_like(sine_waves), :98

            def body(st):
                pe.run(st)
                pd.zp.copy_(pe.zp[:, :, head:])
                pd.mask.copy_(pe.mask[:, head:])
                pd.run(st)

            # the kernel sequence of a (B, T, head) shape is static: capture it once into a CUDA graph and replay it
            # (inputs / noise were copied into the plans' persistent buffers above; RNG draws stay outside the graph)
            key = (B, T, head)
            ent = self.graphs.pop(key, None) if self.use_graph else None
            gr = ent[0] if ent is not None else None
            if self.use_graph and gr is None and key in self._warm:
                if len(self.graphs) >= self.max_plans // 2:
                    self.graphs.pop(next(iter(self.graphs)))   # least recently replayed
                gr = torch.cuda.CUDAGraph()
                torch.cuda.synchronize(self.dev)
                with torch.cuda.graph(gr):
                    body(torch.cuda.current_stream(self.dev).cuda_stream)
                ent = (gr, (("enc", B, T), ("dec", B, T2)))
            if ent is not None:
                self.graphs[key] = ent   # most recently used last
            if gr is not None:
                gr.replay()
            else:
                body(stream)
                self._warm.add(key)

            # fresh tensors for the caller: one copy of the decoder plan's result arena, one of the encoder statistics
            I = c["I"]
            res = pd.out_arena.to(out_dtype, copy=True)
            n_o, n_m, n_z = pd.out_offsets
            Lfin = T2 * c["upp"]
            o = res[:B * Lfin].view(B, 1, Lfin)
            x_mask = res[n_o:n_o + B * T2].view(B, 1, T2)
            z = res[n_o + n_m:n_o + n_m + B * I * T2].view(B, I, T2)
            z_p = res[n_o + n_m + n_z:n_o + n_m + n_z + B * I * T2].view(B, I, T2)
            stats = pe.stats.to(out_dtype, copy=True)
            m_p, logs_p = stats[:, :I], stats[:, I:]
        return o, x_mask, (z, z_p, m_p, logs_p)
