"""Drop-in replacement for the reference ``Synthesizer`` on the inference path, backed by librvc_b200.so.

Boundary mirrored from /root/reference/main/library/algorithm/synthesizers.py:
  * constructor signature  — ``Synthesizer.__init__`` :397
  * ``state_dict`` keys    — the module tree built at :397-432 (and residuals.py / modules.py); the on-disk aliases
                             ``weight_g``/``weight_v`` written by main/inference/train.py:742 load through torch's own
                             weight-norm compatibility hook, exactly as they do in the reference
  * ``infer(phone, phone_lengths, pitch, nsff0, sid, rate)`` -> ``(o, x_mask, (z, z_p, m_p, logs_p))`` — :446-465
  * caller protocol        — ``del net_g.enc_q``; ``load_state_dict(strict=False)``; ``.eval().to(device)``;
                             ``.half()`` / ``.float()`` (main/inference/convert.py:564-569)

The ``torch.nn`` sub-modules below are PARAMETER CONTAINERS ONLY (so that keys, shapes, dtypes and the weight-norm
parametrisation match); none of their ``forward`` methods is ever called.  All arithmetic of ``infer`` runs in the
sm_100a kernels behind the C ABI (include/rvc_b200.h).  There is no CPU / PyTorch fallback: calling ``infer`` on a
module whose parameters are not on a CUDA device, or without the built library, raises ``RuntimeError``.
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn
from torch.nn.utils.parametrizations import weight_norm

LRELU_SLOPE = 0.1
WINDOW = 10  # Encoder(window_size=10), synthesizers.py:318


class _Box(nn.Module):
    """Parameter container; never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("rvc_maker_b200 sub-modules are parameter containers; call Synthesizer.infer")


def _wavenet_box(hidden, kernel_size, n_layers, gin):
    # modules.py:10-33
    b = _Box()
    b.in_layers = nn.ModuleList()
    b.res_skip_layers = nn.ModuleList()
    if gin != 0:
        b.cond_layer = weight_norm(nn.Conv1d(gin, 2 * hidden * n_layers, 1), name="weight")
    for i in range(n_layers):
        b.in_layers.append(weight_norm(nn.Conv1d(hidden, 2 * hidden, kernel_size, padding=(kernel_size - 1) // 2), name="weight"))
        b.res_skip_layers.append(weight_norm(nn.Conv1d(hidden, hidden if i == n_layers - 1 else 2 * hidden, 1), name="weight"))
    return b


def _resblock_box(ch, k, dilations):
    # residuals.py:23-30
    b = _Box()
    b.convs1 = nn.ModuleList([weight_norm(nn.Conv1d(ch, ch, k, 1, dilation=d, padding=(k * d - d) // 2)) for d in dilations])
    b.convs2 = nn.ModuleList([weight_norm(nn.Conv1d(ch, ch, k, 1, dilation=1, padding=(k - 1) // 2)) for _ in dilations])
    return b


class Synthesizer(nn.Module):
    def __init__(self, spec_channels, segment_size, inter_channels, hidden_channels, filter_channels, n_heads, n_layers,
                 kernel_size, p_dropout, resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                 upsample_initial_channel, upsample_kernel_sizes, spk_embed_dim, gin_channels, sr, use_f0,
                 text_enc_hidden_dim=768, vocoder="Default", checkpointing=False, onnx=False, **kwargs):
        super().__init__()
        if vocoder not in ("Default", None):
            raise NotImplementedError(f"vocoder={vocoder!r}: only the Default NSF-HiFiGAN / HiFi-GAN generators are on the B200 path")
        if gin_channels == 0:
            raise NotImplementedError("gin_channels == 0 is not used by any RVC config")
        if hidden_channels % n_heads or (hidden_channels // n_heads) % 4:
            raise ValueError("hidden_channels / n_heads must be a multiple of 4")
        self.spec_channels = spec_channels
        self.inter_channels = inter_channels
        self.hidden_channels = hidden_channels
        self.filter_channels = filter_channels
        self.n_heads = n_heads
        self.n_layers = n_layers
        self.kernel_size = kernel_size
        self.p_dropout = float(p_dropout)
        self.resblock_kernel_sizes = list(resblock_kernel_sizes)
        self.resblock_dilation_sizes = [list(d) for d in resblock_dilation_sizes]
        self.upsample_rates = list(upsample_rates)
        self.upsample_initial_channel = upsample_initial_channel
        self.upsample_kernel_sizes = list(upsample_kernel_sizes)
        self.segment_size = segment_size
        self.gin_channels = gin_channels
        self.spk_embed_dim = spk_embed_dim
        self.use_f0 = bool(use_f0)
        self.sr = sr
        self.text_enc_hidden_dim = text_enc_hidden_dim
        self.upp = int(math.prod(self.upsample_rates))
        H, I, Fc, U, G = hidden_channels, inter_channels, filter_channels, upsample_initial_channel, gin_channels

        # ---- enc_p (TextEncoder :350-364, Encoder :317-338)
        e = _Box()
        e.emb_phone = nn.Linear(text_enc_hidden_dim, H)
        if self.use_f0:
            e.emb_pitch = nn.Embedding(256, H)
        enc = _Box()
        enc.attn_layers, enc.norm_layers_1, enc.ffn_layers, enc.norm_layers_2 = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        dk = H // n_heads
        for _ in range(n_layers):
            a = _Box()
            a.conv_q, a.conv_k, a.conv_v, a.conv_o = (nn.Conv1d(H, H, 1) for _ in range(4))
            a.emb_rel_k = nn.Parameter(torch.randn(1, 2 * WINDOW + 1, dk) * dk ** -0.5)
            a.emb_rel_v = nn.Parameter(torch.randn(1, 2 * WINDOW + 1, dk) * dk ** -0.5)
            enc.attn_layers.append(a)
            for lst in (enc.norm_layers_1, enc.norm_layers_2):
                n = _Box()
                n.gamma = nn.Parameter(torch.ones(H))
                n.beta = nn.Parameter(torch.zeros(H))
                lst.append(n)
            f = _Box()
            f.conv_1 = nn.Conv1d(H, Fc, kernel_size)
            f.conv_2 = nn.Conv1d(Fc, H, kernel_size)
            enc.ffn_layers.append(f)
        e.encoder = enc
        e.proj = nn.Conv1d(H, I * 2, 1)
        self.enc_p = e

        # ---- dec (GeneratorNSF :114-142, Generator :21-37)
        d = _Box()
        n_up = len(self.upsample_rates)
        chans = [U // (2 ** (i + 1)) for i in range(n_up)]
        d.conv_pre = nn.Conv1d(I, U, 7, 1, padding=3)
        if self.use_f0:
            ms = _Box()
            ms.l_linear = nn.Linear(1, 1)
            d.m_source = ms
            d.ups, d.noise_convs = nn.ModuleList(), nn.ModuleList()
            for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
                if u % 2 or (k - u) % 2:
                    raise NotImplementedError("odd upsample rates are not used by any RVC config")
                d.ups.append(weight_norm(nn.ConvTranspose1d(U // (2 ** i), chans[i], k, u, padding=(k - u) // 2)))
                s = int(math.prod(self.upsample_rates[i + 1:])) if i + 1 < n_up else 1
                kn = 1 if s == 1 else s * 2 - s % 2
                d.noise_convs.append(nn.Conv1d(1, chans[i], kernel_size=kn, stride=s, padding=0 if s == 1 else (kn - s) // 2))
            d.resblocks = nn.ModuleList([_resblock_box(chans[i], k, dl) for i in range(n_up)
                                         for k, dl in zip(self.resblock_kernel_sizes, self.resblock_dilation_sizes)])
        else:
            d.ups_and_resblocks = nn.ModuleList()
            for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
                if (k - u) % 2:
                    raise NotImplementedError("(kernel - stride) must be even")
                d.ups_and_resblocks.append(weight_norm(nn.ConvTranspose1d(U // (2 ** i), chans[i], k, u, padding=(k - u) // 2)))
                for kk, dl in zip(self.resblock_kernel_sizes, self.resblock_dilation_sizes):
                    d.ups_and_resblocks.append(_resblock_box(chans[i], kk, dl))
        d.conv_post = nn.Conv1d(chans[-1], 1, 7, 1, padding=3, bias=False)
        d.cond = nn.Conv1d(G, U, 1)
        self.dec = d

        # ---- enc_q (PosteriorEncoder :373-386) — train-only, kept so `del net_g.enc_q` and full checkpoints work
        q = _Box()
        q.pre = nn.Conv1d(spec_channels, H, 1)
        q.enc = _wavenet_box(H, 5, 16, G)
        q.proj = nn.Conv1d(H, I * 2, 1)
        self.enc_q = q

        # ---- flow (ResidualCouplingBlock residuals.py:71-85)
        fl = _Box()
        fl.flows = nn.ModuleList()
        for _ in range(4):
            c = _Box()
            c.pre = nn.Conv1d(I // 2, H, 1)
            c.enc = _wavenet_box(H, 5, 3, G)
            c.post = nn.Conv1d(H, I // 2, 1)
            c.post.weight.data.zero_()
            c.post.bias.data.zero_()
            fl.flows.append(c)
            fl.flows.append(_Box())  # Flip: no parameters
        self.flow = fl
        self.emb_g = nn.Embedding(spk_embed_dim, G)

        # "tc": every conv on tcgen05 tensor cores with fp16 operands / fp32 accumulation; "tc-bf16": same with bf16 operands in the
        # C = 256 ResBlocks; "fp32": CUDA-core fp32 everywhere
        self.precision = kwargs.get("precision", os.environ.get("RVC_B200_PRECISION", "tc"))
        # replay the static kernel sequence of a (B, T) shape as a CUDA graph from its second call on
        self.use_cuda_graph = bool(kwargs.get("use_cuda_graph", True))
        self._engine = None
        self._noise_override = None  # tests: (eps_z [B,I,T], eps_src [B,T*upp,1] or None)
        self._host_out = None        # pinned read-back buffer of infer_to_host()

    # ------------------------------------------------------------------ engine lifetime
    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, *a, **k):
        self._engine = None
        return super().load_state_dict(state_dict, strict, *a, **k)

    def refold(self):
        """Drop the folded/packed weight cache (call after mutating parameters in place)."""
        self._engine = None

    def forward(self, *a, **k):
        raise NotImplementedError("training forward (synthesizers.py:434-443) is outside the B200 inference path")

    def remove_weight_norm(self):
        """No-op: weight-norm is folded into packed weights at load time (the reference's own remover raises)."""
        return None

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def infer(self, phone, phone_lengths, pitch=None, nsff0=None, sid=None, rate=None):
        dev = self.emb_g.weight.device
        if dev.type != "cuda":
            raise RuntimeError("rvc_maker_b200.Synthesizer.infer needs its parameters on a CUDA device (no CPU fallback)")
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self, dev)
        if rate is not None:
            assert isinstance(rate, torch.Tensor)
            rate = float(rate.item())
        noise = self._noise_override
        return self._engine.infer(phone, phone_lengths, pitch, nsff0, sid, rate, self.emb_g.weight.dtype, noise)

    @torch.no_grad()
    def infer_to_host(self, phone, phone_lengths, pitch=None, nsff0=None, sid=None, rate=None):
        """``infer(...)[0]`` read back into a pinned host buffer: float32 ``[B, 1, L]``, complete when the call returns.

        Stands in for the caller's ``net_g.infer(...)[0][0, 0].data.cpu().float()`` (main/inference/convert.py:381): a pageable
        ``.cpu()`` costs 0.16 ms per 10 s chunk, the pinned asynchronous copy 0.04 ms.  The returned tensor is a view of a
        buffer owned by the module and is overwritten by the next call.
        """
        o = self.infer(phone, phone_lengths, pitch, nsff0, sid, rate)[0]
        n = o.numel()
        if self._host_out is None or self._host_out.numel() < n:
            self._host_out = torch.empty(n, dtype=torch.float32).pin_memory()
        out = self._host_out[:n].view(o.shape)
        out.copy_(o, non_blocking=True)
        torch.cuda.current_stream(o.device).synchronize()
        return out

    def _ensure_engine(self):
        dev = self.emb_g.weight.device
        if dev.type != "cuda":
            raise RuntimeError("rvc_maker_b200.Synthesizer needs its parameters on a CUDA device (no CPU fallback)")
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self, dev)
        return self._engine

    @torch.no_grad()
    def infer_many(self, items):
        """``infer`` over an iterable of ``(phone, phone_lengths, pitch, nsff0, sid[, rate[, noise]])`` tuples (device or pinned host tensors), pipelined:
        yields ``(o, x_mask, (z, z_p, m_p, logs_p))`` per item, in order, with the RNG draw order of ``infer`` called item by item.

        The reference's chunk loop (main/inference/convert.py:442-447) synthesises independent chunks one after the other; here the TextEncoder of
        chunk i+1 runs on a second stream next to the generator of chunk i (``Engine.infer_lanes``), and consecutive short items of one shape (three
        10 s chunks, ``Engine.group_size``) run as one batch, each with its own inputs, speaker and noise draws: a group's results are bit-identical to
        ``infer`` on the stacked batch with the same noise.  Against ``infer`` item by item the encoder and flow outputs are bit-identical too, while the
        tensor-core generator picks its tile shapes by the batch and so sums in a different order (waveform differences of ~2e-4, the size of either
        result's distance to the fp32 reference); RVC_B200_GROUP_FRAMES=0 switches the batching off and restores item-by-item bit-identity.  A result is
        handed out once the next group has been enqueued (or at the end), ordered on the caller's current stream; the tensors of one item are views of
        its group's fresh result buffer.
        """
        eng = self._ensure_engine()
        dt = self.emb_g.weight.dtype

        def norm(it):
            for item in it:
                item = tuple(item)
                rate = item[5] if len(item) > 5 else None
                if rate is not None:
                    assert isinstance(rate, torch.Tensor)
                    rate = float(rate.item())
                yield item[:5] + (rate, item[6] if len(item) > 6 else None)

        yield from eng.infer_lanes(norm(items), dt)

    @torch.no_grad()
    def infer_stream(self, items, depth=2):
        """Pipelined ``infer_to_host`` over an iterable of ``(phone, phone_lengths, pitch, nsff0, sid[, noise])`` tuples of (pinned) HOST tensors: yields
        ``o`` as float32 ``[B, 1, L]`` views of pinned host buffers, in order.

        The caller's chunk loop (main/inference/convert.py:442-447) is strictly serial: copy in, synthesise, copy out.  Here the host-to-device
        copy of item n+1 runs on a copy stream while item n computes, the read-back of item n runs on a second copy stream while item n+1
        computes, and the compute itself runs in two lanes (``infer_many``: encoder of item n+1 next to the generator of item n).  Every item's
        copies still happen, they are just off the critical path.  A yielded view stays valid until ``depth`` further items have been yielded.
        Same RNG order as ``infer`` called item by item; numerics as ``infer_many`` (consecutive short items of one shape run as one batch).
        """
        dev = self.emb_g.weight.device
        if dev.type != "cuda":
            raise RuntimeError("rvc_maker_b200.Synthesizer.infer_stream needs its parameters on a CUDA device (no CPU fallback)")
        comp = torch.cuda.current_stream(dev)
        # Two pools that grow on demand and are reused first-in first-out (allocated during the first calls, then steady): device staging sets
        # (taken when an item is copied in, returned when its result has been enqueued for read-back: one being filled ahead + two groups in the
        # lanes) and pinned host result buffers (taken at read-back, returned once `depth` further results have been handed out).
        if getattr(self, "_pipe", None) is None or self._pipe["dev"] != dev:
            self._pipe = dict(dev=dev, h2d=torch.cuda.Stream(dev), d2h=torch.cuda.Stream(dev), stage=[], free_stage=[], host=[], free_host=[])
        P = self._pipe
        h2d, d2h = P["h2d"], P["d2h"]

        def stage(item):
            """host -> device into a free staging set on the copy stream (after the lanes have finished the item that used the set before)"""
            item = item[:5]
            if P["free_stage"]:
                k = P["free_stage"].pop(0)
            else:
                k = len(P["stage"])
                P["stage"].append(dict(bufs=None, free=None))
            S = P["stage"][k]
            bufs = S["bufs"]
            if bufs is None or any((a is None) != (b is None) or (a is not None and (a.shape != b.shape or a.dtype != b.dtype)) for a, b in zip(bufs, item)):
                bufs = S["bufs"] = [None if t is None else torch.empty(t.shape, dtype=t.dtype, device=dev) for t in item]
            if S["free"] is not None:
                h2d.wait_event(S["free"])
            with torch.cuda.stream(h2d):
                for d, t in zip(bufs, item):
                    if t is not None:
                        d.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(h2d)
            return k, ev

        slots = []                        # staging set of every item handed to the lanes, in order

        def staged():
            it = iter(items)
            nxt = next(it, None)
            k_next, ev_next = stage(nxt) if nxt is not None else (None, None)
            while nxt is not None:
                k, cur_ev, cur_noise = k_next, ev_next, (nxt[5] if len(nxt) > 5 else None)
                nxt = next(it, None)
                if nxt is not None:
                    k_next, ev_next = stage(nxt)            # overlaps the compute of the items before it
                comp.wait_event(cur_ev)
                slots.append(k)
                yield tuple(P["stage"][k]["bufs"]) + (None, cur_noise)

        def read_back(k, o):
            S = P["stage"][k]
            S["free"] = torch.cuda.Event()
            S["free"].record(comp)                         # (ordered behind the item's result on the caller's stream: both lanes are done with the set)
            P["free_stage"].append(k)
            n = o.numel()
            if P["free_host"]:
                h = P["free_host"].pop(0)
            else:
                h = len(P["host"])
                P["host"].append(None)
            if P["host"][h] is None or P["host"][h].numel() < n:
                P["host"][h] = torch.empty(n, dtype=torch.float32).pin_memory()
            out = P["host"][h][:n].view(o.shape)
            d2h.wait_event(S["free"])
            with torch.cuda.stream(d2h):
                out.copy_(o, non_blocking=True)
                o.record_stream(d2h)
                done = torch.cuda.Event()
                done.record(d2h)
            return h, done, out

        pending, handed = [], []          # read-backs in flight; host buffers whose views the caller may still be using
        try:
            for i, res in enumerate(self.infer_many(staged())):
                pending.append(read_back(slots[i], res[0]))
                while len(pending) >= depth:
                    h, done, out = pending.pop(0)
                    done.synchronize()
                    handed.append(h)
                    if len(handed) > depth:
                        P["free_host"].append(handed.pop(0))
                    yield out
            for h, done, out in pending:
                done.synchronize()
                handed.append(h)
                yield out
        finally:
            P["free_host"].extend(handed)                  # (the views of this call expire with it)
            for k in range(len(P["stage"])):               # (only after an exception: sets whose item never reached its read-back)
                if k not in P["free_stage"]:
                    P["stage"][k]["free"] = torch.cuda.Event()
                    P["stage"][k]["free"].record(comp)     # infer_many has made the caller's stream wait for both lanes by now
                    P["free_stage"].append(k)
