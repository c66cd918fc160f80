"""Chunk pipeline around ``Synthesizer.infer`` — the caller on both sides of the hot path.

Mirrors ``VC.voice_conversion`` from the embedder output on (main/inference/convert.py:361-381) and the chunk loop of
``VC.pipeline`` (:403-452).  The reference converts one chunk at a time: features -> ``F.interpolate(x2)`` -> protect blend ->
``infer`` -> ``.cpu().numpy()`` -> ``empty_cache()``, then ``np.concatenate`` and a peak normalisation.  Here

* chunks with the same shape are batched into one ``infer`` (the 38 s chunks of a long stream all have T = 4000);
* ``bucket=N`` (optional): the reference cuts real audio at energy minima (:403-420), so every chunk has its own length and each
  would need its own launch plan / CUDA graph.  With a bucket, chunks are zero-padded to the next multiple of N frames and run
  with their true ``phone_lengths`` — exactly what the reference computes for the shorter items of a ragged batch — so a few
  shapes serve the whole stream and unequal chunks batch together.  Encoder and flow are masked at every layer (identical
  values); the generator's tail sees the padding only within its receptive field (~10 frames), far inside the ``x_pad`` = 1 s of
  context that is trimmed from every piece;
* up-sampling + protect blend are one kernel (``rvc_feature_prep_f32``), bit-identical to the reference's elementwise ops;
* inputs of batch i+1 are staged host->device on a copy stream while batch i computes, finished pieces are trimmed on the
  device and copied into ONE pinned host buffer at their final offsets on a second copy stream (no per-chunk sync, no concatenate);
* the peak (``np.abs(audio_opt).max()``, :451) is accumulated on the device (``rvc_absmax_f32``) and applied once on the host;
* with several ranks every rank synthesises the chunks ``shard_chunks`` gives it; the pieces are gathered once over NCCL
  (the only collective, after the data path) and rank 0 assembles the stream.

* the split-point search at the head of the loop (:404-412; 160 host passes over the whole signal in the reference) is one launch over the
  search windows, bit-identical indices (``split_points`` below; ``chunks.split_points`` is the host planner the CPU tests use).

* the high-pass in front of it (``signal.filtfilt``, :402) has a device version (``filtfilt`` below: scipy's recurrence in blocks, within the filter's
  rounding noise of scipy, not bit-identical — read its docstring before building plans on it).

* the retrieval blend behind the FAISS search (:354-359) runs on a device-resident ``big_npy``, bit-identical (``retrieval_blend`` below).

Out of scope (hooks instead): the embedder and the FAISS search that produce ``feats`` / ``feats0`` (:336-353), the f0 network
(:431, 8f-2; its two ends are in ``rvc_maker_b200.pitch``) and ``change_rms`` (:450, only when volume_envelope != 1).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .chunks import WINDOW, ChunkConfig, chunk_spans, shard_chunks
from .chunks import split_points as ch_split_points   # the host planner (numpy), next to the device search below


def split_points(audio, cfg=ChunkConfig()):
    """``chunks.split_points`` on the device — the search at the head of ``VC.pipeline``'s chunk loop (convert.py:404-412): ``opt_ts`` as a list of ints.

    audio: 1-D CUDA tensor of the (unpadded) 16 kHz signal, float64 (what the reference's ``signal.filtfilt`` hands on, :402) or float32.  The 160
    passes over the whole signal that the reference makes on the host (1.5 s for ten minutes of audio) become one launch over the 2 x ``x_query``
    seconds around every candidate; the sums are formed in numpy's order and dtype, so the indices are the reference's, bit for bit."""
    if not isinstance(audio, torch.Tensor) or audio.device.type != "cuda":
        raise RuntimeError("rvc_maker_b200.pipeline.split_points needs a CUDA tensor (no CPU fallback; chunks.split_points is the host planner)")
    if audio.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"audio must be float32 or float64, got {audio.dtype}")
    audio = audio.reshape(-1).contiguous()
    n = audio.numel()
    if n + WINDOW <= cfg.t_max:                        # :406 (audio_pad.shape[0] > t_max)
        return []
    n_t = (n - 1) // cfg.t_center
    if n_t == 0:
        return []
    lib = _lib.load()
    ws = torch.empty(lib.rvc_split_points_workspace(n, cfg.t_center, cfg.t_query), dtype=torch.uint8, device=audio.device)
    out = torch.empty(n_t, dtype=torch.int64, device=audio.device)
    st = torch.cuda.current_stream(audio.device).cuda_stream
    _lib.check(lib.rvc_split_points(audio.data_ptr(), int(audio.dtype == torch.float64), n, WINDOW, cfg.t_center, cfg.t_query, out.data_ptr(),
                                    ws.data_ptr(), st), "rvc_split_points")
    return out.tolist()


_filt_cache = {}


def _filter_constants(b, a, block):
    """(b, a, zi, M^block) as contiguous float64 arrays.  zi as scipy.signal.lfilter_zi computes it (steady state of the step response); M^block by running
    the zero-input recurrence ``block`` steps from every unit state — the companion form is far from normal, repeated squaring of M loses everything."""
    key = (tuple(b), tuple(a), block)
    if key not in _filt_cache:
        b64, a64 = np.asarray(b, np.float64), np.asarray(a, np.float64)
        if b64.shape != a64.shape or b64.ndim != 1:
            raise ValueError("b and a must be 1-D and of equal length (scipy.signal.butter gives them so)")
        if a64[0] != 1.0:
            b64, a64 = b64 / a64[0], a64 / a64[0]
        y_inf = np.sum(b64) / np.sum(a64)
        zi = np.flip(np.cumsum(np.flip(b64 - y_inf * a64)))[1:].copy()
        ns = len(a64) - 1
        Z = np.eye(ns)                                     # column j = the state reached from unit state j
        for _ in range(block):
            y = Z[0].copy()
            Z[:-1] = Z[1:] - a64[1:-1, None] * y[None, :]
            Z[-1] = -a64[-1] * y
        _filt_cache[key] = (np.ascontiguousarray(b64), np.ascontiguousarray(a64), zi, np.ascontiguousarray(Z))
    return _filt_cache[key]


def filtfilt(audio, b, a):
    """``scipy.signal.filtfilt(b, a, audio)`` (scipy's defaults, as convert.py:402 calls it) on the device: a float64 CUDA tensor [n].

    audio: 1-D CUDA tensor, float32 or float64; b, a: the coefficients (``signal.butter(N=5, Wn=48, btype="high", fs=16000)``, convert.py:29).  The
    recurrence and its operation order are scipy's inside blocks of 4096 samples, the blocks run in parallel (137 - 259 ms on the host per 10 minutes of
    audio, 2 ms here); the result agrees with scipy's to ~2e-8 absolute at amplitude 0.5 — the rounding noise of the filter itself —
    not to the bit, so keep scipy's call where bit-identical split points matter (``split_points`` is exact for whatever signal it is given)."""
    if not isinstance(audio, torch.Tensor) or audio.device.type != "cuda":
        raise RuntimeError("rvc_maker_b200.pipeline.filtfilt needs a CUDA tensor (no CPU fallback)")
    if audio.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"audio must be float32 or float64, got {audio.dtype}")
    lib = _lib.load()
    audio = audio.reshape(-1).contiguous()
    n = audio.numel()
    b64, a64, zi, ML = _filter_constants(b, a, lib.rvc_filtfilt_block(n))
    y = torch.empty(n, dtype=torch.float64, device=audio.device)
    ws = torch.empty(max(lib.rvc_filtfilt_workspace(n, len(b64)), 8), dtype=torch.uint8, device=audio.device)
    st = torch.cuda.current_stream(audio.device).cuda_stream
    _lib.check(lib.rvc_filtfilt(audio.data_ptr(), int(audio.dtype == torch.float64), n, b64.ctypes.data, a64.ctypes.data, len(b64), zi.ctypes.data,
                                ML.ctypes.data, y.data_ptr(), ws.data_ptr(), st), "rvc_filtfilt")
    return y


def retrieval_blend(feats, big_npy, score, ix, index_rate):
    """The blend behind ``score, ix = index.search(npy, k=8)`` (convert.py:354-359) on the device: ``feats`` [T, D] (or [1, T, D]) fp32 CUDA, ``big_npy``
    [N, D] fp32 CUDA (``index.reconstruct_n(0, index.ntotal)``, :395, uploaded once), ``score`` [T, 8] fp32 and ``ix`` [T, 8] int64 as FAISS returns
    them (host or device).  Returns the blended features in ``feats``' shape, bit-identical to the reference's numpy + torch statements; the
    reference gathers the [T, 8, D] neighbours on the host and copies the result up again."""
    if not isinstance(feats, torch.Tensor) or feats.device.type != "cuda" or not isinstance(big_npy, torch.Tensor) or big_npy.device != feats.device:
        raise RuntimeError("rvc_maker_b200.pipeline.retrieval_blend needs feats and big_npy as CUDA tensors on one device (no CPU fallback)")
    shape = feats.shape
    f2 = feats.float().reshape(-1, shape[-1]).contiguous()
    T, D = f2.shape
    big_npy = big_npy.float().contiguous()
    score = torch.as_tensor(score, dtype=torch.float32).to(feats.device).contiguous()
    ix = torch.as_tensor(ix, dtype=torch.int64).to(feats.device).contiguous()
    if tuple(score.shape) != (T, 8) or tuple(ix.shape) != (T, 8) or big_npy.dim() != 2 or big_npy.shape[1] != D:
        raise ValueError(f"need score / ix [T={T}, 8] and big_npy [N, {D}], got {tuple(score.shape)}, {tuple(ix.shape)}, {tuple(big_npy.shape)}")
    if T and (int(ix.min()) < 0 or int(ix.max()) >= big_npy.shape[0]):
        raise ValueError("ix holds indices outside big_npy (FAISS returns -1 when a list has fewer than k vectors)")
    out = torch.empty_like(f2)
    st = torch.cuda.current_stream(feats.device).cuda_stream
    _lib.check(_lib.load().rvc_retrieval_blend_f32(f2.data_ptr(), big_npy.data_ptr(), score.data_ptr(), ix.data_ptr(), float(index_rate),
                                                   float(1 - index_rate), out.data_ptr(), T, D, 8, st), "rvc_retrieval_blend_f32")
    return out.reshape(shape)


def feature_prep(feats, feats0, pitchf, protect, T=None):
    """``F.interpolate(feats, scale_factor=2)`` + protect blend on the device (convert.py:361-378).

    feats, feats0 (or None): [B, Th, D] fp32 CUDA tensors; pitchf: [B, >=T] fp32 (Hz) or None; returns [B, T, D] (T defaults
    to 2*Th).  The blend runs only ``if protect < 0.5 and pitch_guidance`` (:362, :372), as in the reference.
    """
    if feats.device.type != "cuda":
        raise RuntimeError("rvc_maker_b200.pipeline.feature_prep needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    feats = feats.float().contiguous()
    B, Th, D = feats.shape
    T = 2 * Th if T is None else int(T)
    blend = feats0 is not None and pitchf is not None and protect < 0.5
    f0p = pfp = None
    pbs = 0
    if blend:
        feats0 = feats0.float().contiguous()
        pitchf = pitchf.float()
        if pitchf.stride(-1) != 1:
            pitchf = pitchf.contiguous()
        if pitchf.shape[1] < T:
            raise ValueError(f"pitchf has {pitchf.shape[1]} frames, features need {T}")
        f0p, pfp, pbs = feats0.data_ptr(), pitchf.data_ptr(), pitchf.stride(0)
    out = torch.empty(B, T, D, device=feats.device, dtype=torch.float32)
    st = torch.cuda.current_stream(feats.device).cuda_stream
    _lib.check(lib.rvc_feature_prep_f32(feats.data_ptr(), f0p, pfp, pbs, float(protect), out.data_ptr(), B, Th, T, D, st), "rvc_feature_prep_f32")
    return out


def conversion_inputs(feats, p_len, pitch, pitchf, sid, protect=0.5, feats0=None):
    """convert.py:361-380 for a batch of equal-shape chunks: the arguments of ``net_g.infer`` — ``(phone, lengths, pitch, pitchf, sid)``.

    feats / feats0: [B, Th, D] embedder output (after / before the retrieval blend), p_len = chunk samples // window (:363),
    pitch / pitchf: [B, frames] (None for a model without pitch guidance).
    """
    guided = pitch is not None and pitchf is not None
    T = 2 * feats.shape[1]
    if T < p_len:                       # :365-369
        p_len = T
        if guided:
            pitch, pitchf = pitch[:, :p_len], pitchf[:, :p_len]
    if guided and pitch.shape[1] != T:   # the reference fails inside TextEncoder with a broadcast error here
        raise ValueError(f"features have {T} frames after up-sampling but pitch has {pitch.shape[1]}")
    phone = feature_prep(feats, feats0 if guided else None, pitchf if guided else None, protect, T)
    lengths = torch.full((feats.shape[0],), p_len, dtype=torch.long, device=feats.device)   # :380
    return phone, lengths, (pitch if guided else None), (pitchf.float() if guided else None), sid


def voice_conversion(net_g, feats, p_len, pitch, pitchf, sid, protect=0.5, feats0=None):
    """convert.py:361-381 for a batch of equal-shape chunks: returns the waveforms ``infer(...)[0][:, 0]`` ([B, T*upp], on device)."""
    return net_g.infer(*conversion_inputs(feats, p_len, pitch, pitchf, sid, protect, feats0))[0][:, 0]


def conversion_inputs_padded(feats, lengths, pitch, pitchf, sid, protect=0.5, feats0=None):
    """``conversion_inputs`` for a bucketed batch: feats / feats0 [B, Tb/2, D], pitch / pitchf [B, Tb] are zero-padded (pitch index 1)
    behind every item's own frames and ``lengths`` [B] holds each item's p_len (settled per item by the caller, :363-369)."""
    guided = pitch is not None and pitchf is not None
    phone = feature_prep(feats, feats0 if guided else None, pitchf if guided else None, protect, 2 * feats.shape[1])
    return phone, lengths, (pitch if guided else None), (pitchf.float() if guided else None), sid


def voice_conversion_padded(net_g, feats, lengths, pitch, pitchf, sid, protect=0.5, feats0=None):
    """``voice_conversion`` for a bucketed batch (``conversion_inputs_padded``).  Returns [B, Tb*upp]; item b is valid up to its own ``T_b * upp`` samples."""
    return net_g.infer(*conversion_inputs_padded(feats, lengths, pitch, pitchf, sid, protect, feats0))[0][:, 0]


def batch_chunks(chunks, max_batch, bucket=0):
    """Group chunks of identical length (frames and samples) — or, with ``bucket``, of the same length rounded up to a multiple of
    ``bucket`` frames — into batches of at most ``max_batch``, longest first."""
    if bucket:
        key = lambda c: -(-c.frames // bucket)
    else:
        key = lambda c: (c.frames, c.sample_end - c.sample_start)
    groups, cur = [], []
    for c in sorted(chunks, key=lambda c: (-c.frames, -(c.sample_end - c.sample_start), c.index)):
        if cur and (len(cur) == max_batch or key(c) != key(cur[0])):
            groups.append(cur)
            cur = []
        cur.append(c)
    if cur:
        groups.append(cur)
    return groups


def piece_layout(spans, upp, tgt_sr, cfg=ChunkConfig()):
    """Offset and length (target-rate samples) of every chunk's trimmed piece in the stitched stream (convert.py:444-448)."""
    t_pad_tgt = tgt_sr * cfg.x_pad
    lens = [max(c.frames * upp - 2 * t_pad_tgt, 0) for c in spans]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return [int(o) for o in offs[:-1]], lens, int(offs[-1])


class StreamConverter:
    """``VC.pipeline`` from the chunk plan on (convert.py:403-452), batched and overlapped.  One instance per process / GPU."""

    def __init__(self, net_g, tgt_sr, upp, *, cfg=ChunkConfig(), max_batch=2, device=None, bucket=0):
        self.net, self.sr, self.upp, self.cfg, self.max_batch = net_g, int(tgt_sr), int(upp), cfg, int(max_batch)
        self.bucket = int(bucket)
        if self.bucket < 0 or self.bucket % 2:
            raise ValueError("bucket must be an even number of frames (0 = exact shapes)")
        self.dev = torch.device(device) if device is not None else next(net_g.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("rvc_maker_b200.pipeline.StreamConverter needs the model on a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.h2d = torch.cuda.Stream(self.dev)
        self.d2h = torch.cuda.Stream(self.dev)
        self._host = None       # pinned output stream, reused between calls
        self._peak = torch.zeros(1, device=self.dev)
        self.last_stats = {}

    # ------------------------------------------------------------------ staging
    def _stage(self, group, embed, pitch, pitchf, sid):
        """Issue the host->device copies of one batch on the copy stream; returns the device tensors and the event to wait on."""
        with torch.cuda.stream(self.h2d):
            fl, f0l = [], []
            for c in group:
                f, f0 = embed(c)
                fl.append(f)
                f0l.append(f0)
            if self.bucket:
                return self._stage_padded(group, fl, f0l, pitch, pitchf, sid)
            if any(f.shape != fl[0].shape for f in fl):
                raise ValueError("chunks of one batch must give features of one shape")
            feats = torch.stack([f.to(self.dev, non_blocking=True) for f in fl])
            if any(f is None for f in f0l):
                feats0 = None
            elif all(f0 is f for f, f0 in zip(fl, f0l)):
                feats0 = feats      # no retrieval blend happened (index_rate = 0): feats0 = feats.clone(), one copy is enough
            else:
                feats0 = torch.stack([f.to(self.dev, non_blocking=True) for f in f0l])
            pi = pf = None
            if pitch is not None:
                pi = torch.stack([pitch[c.frame_start:c.frame_end].to(self.dev, non_blocking=True) for c in group])
                pf = torch.stack([pitchf[c.frame_start:c.frame_end].to(self.dev, non_blocking=True) for c in group])
            sd = sid.to(self.dev, non_blocking=True).reshape(-1)[:1].expand(len(group)).contiguous()
            ev = torch.cuda.Event()
            ev.record(self.h2d)
        nbytes = sum(t.numel() * t.element_size() for t in (feats, None if feats0 is feats else feats0, pi, pf) if t is not None)
        return dict(group=group, feats=feats, feats0=feats0, pitch=pi, pitchf=pf, sid=sd, ev=ev, bytes=nbytes, lengths=None,
                    T_items=[2 * int(feats.shape[1])] * len(group))

    def _stage_padded(self, group, fl, f0l, pitch, pitchf, sid):
        """Bucketed batch (called on the copy stream): every item keeps its own frames, the rest of the bucket is zeros / pitch index 1."""
        B, D = len(group), int(fl[0].shape[1])
        guided = pitch is not None
        T_items, p_lens = [], []
        for c, f in zip(group, fl):
            T, p_len, n_pitch = 2 * int(f.shape[0]), (c.sample_end - c.sample_start) // WINDOW, c.frame_end - c.frame_start
            if T < p_len:                                         # convert.py:365-369
                p_len, n_pitch = T, min(n_pitch, T)
            if guided and n_pitch != T:
                raise ValueError(f"features have {T} frames after up-sampling but pitch has {n_pitch}")
            T_items.append(T)
            p_lens.append(p_len)
        Tb = -(-max(T_items) // self.bucket) * self.bucket
        feats = torch.zeros(B, Tb // 2, D, device=self.dev)
        same = all(f0 is f for f, f0 in zip(fl, f0l))
        feats0 = None if any(f0 is None for f0 in f0l) else (feats if same else torch.zeros(B, Tb // 2, D, device=self.dev))
        pi = torch.ones(B, Tb, dtype=torch.long, device=self.dev) if guided else None
        pf = torch.zeros(B, Tb, device=self.dev) if guided else None
        nbytes = 0
        for b, (c, f, f0) in enumerate(zip(group, fl, f0l)):
            feats[b, :f.shape[0]].copy_(f, non_blocking=True)
            nbytes += f.numel() * 4
            if feats0 is not None and feats0 is not feats:
                feats0[b, :f0.shape[0]].copy_(f0, non_blocking=True)
                nbytes += f0.numel() * 4
            if guided:
                pi[b, :T_items[b]].copy_(pitch[c.frame_start:c.frame_start + T_items[b]], non_blocking=True)
                pf[b, :T_items[b]].copy_(pitchf[c.frame_start:c.frame_start + T_items[b]], non_blocking=True)
                nbytes += T_items[b] * 12
        lengths = torch.tensor(p_lens, dtype=torch.long).to(self.dev)
        sd = sid.to(self.dev, non_blocking=True).reshape(-1)[:1].expand(B).contiguous()
        ev = torch.cuda.Event()
        ev.record(self.h2d)
        return dict(group=group, feats=feats, feats0=feats0, pitch=pi, pitchf=pf, sid=sd, ev=ev, bytes=nbytes, lengths=lengths, T_items=T_items)

    # ------------------------------------------------------------------ convert
    def convert(self, n_samples, opt_ts, embed, pitch, pitchf, sid, *, protect=0.5, rank=0, world=1, noise_fn=None, to_host=True):
        """Synthesise the stream of ``n_samples`` 16 kHz samples cut at ``opt_ts`` (chunks.split_points).

        embed(chunk) -> (feats [Th, D], feats0 [Th, D] or None): fp32 tensors (pinned host or device) for the chunk's audio slice
        ``audio_pad[chunk.sample_start:chunk.sample_end]``; pitch (int64) / pitchf (fp32): [frames of the padded stream] or None;
        sid: 0-d or 1-element int64 tensor.  noise_fn(chunk) -> (eps_z [1, I, T], eps_src [1, T*upp, 1]) pins the model's two
        noise draws (tests).  Returns the stitched float32 waveform as a numpy view of the pinned buffer on rank 0 (valid until the
        next call; None on other ranks), or the per-chunk device pieces when ``to_host`` is False.
        """
        if pitch is not None:
            pitch, pitchf = pitch.reshape(-1), pitchf.reshape(-1)
        spans = chunk_spans(n_samples, opt_ts, self.cfg)
        owner = shard_chunks(spans, world)
        mine = [c for c in spans if owner[c.index] == rank]
        offs, lens, total = piece_layout(spans, self.upp, self.sr, self.cfg)
        t_pad_tgt = self.sr * self.cfg.x_pad
        groups = batch_chunks(mine, self.max_batch, self.bucket)
        comp = torch.cuda.current_stream(self.dev)
        st = comp.cuda_stream
        if to_host and rank == 0 and (self._host is None or self._host.numel() < total):
            self._host = torch.empty(total, dtype=torch.float32).pin_memory()
        self._peak.zero_()
        self.h2d.wait_stream(comp)
        self.d2h.wait_stream(comp)
        staged = self._stage(groups[0], embed, pitch, pitchf, sid) if groups else None
        prev_noise = getattr(self.net, "_noise_override", None)
        try:
            return self._convert_groups(groups, staged, embed, pitch, pitchf, sid, protect, noise_fn, comp, st, t_pad_tgt, lens, offs, total,
                                        spans, owner, rank, world, to_host)
        finally:
            # the pinned draws belong to this call only: later infer()/convert() calls draw fresh noise again, as the reference does
            if noise_fn is not None:
                self.net._noise_override = prev_noise

    def _convert_groups(self, groups, staged, embed, pitch, pitchf, sid, protect, noise_fn, comp, st, t_pad_tgt, lens, offs, total, spans, owner,
                        rank, world, to_host):
        pieces, keep, short, h2d_bytes, shapes = {}, [], False, 0, set()
        ctx = []          # per batch, in order: what the post-processing of its waveform needs

        def items():
            """the arguments of ``infer`` for every batch, in order (staging of batch n+1 overlaps the compute of the batches before it)"""
            nonlocal staged, h2d_bytes
            for gi, group in enumerate(groups):
                cur = staged
                comp.wait_event(cur["ev"])
                staged = self._stage(groups[gi + 1], embed, pitch, pitchf, sid) if gi + 1 < len(groups) else None
                h2d_bytes += cur["bytes"]
                c0 = group[0]
                shapes.add((len(group), 2 * int(cur["feats"].shape[1])))
                noise = None
                if noise_fn is not None:
                    nz = [noise_fn(c) for c in group]
                    if cur["lengths"] is not None:   # bucketed: the pinned draws cover each item's own frames, zeros behind them
                        Tb = cur["pitch"].shape[1] if cur["pitch"] is not None else 2 * cur["feats"].shape[1]
                        nz = [(torch.nn.functional.pad(a, (0, Tb - a.shape[2])), torch.nn.functional.pad(b, (0, 0, 0, Tb * self.upp - b.shape[1])))
                              for a, b in nz]
                    noise = (torch.cat([a for a, _ in nz]), torch.cat([b for _, b in nz]))
                if cur["lengths"] is not None:
                    args = conversion_inputs_padded(cur["feats"], cur["lengths"], cur["pitch"], cur["pitchf"], cur["sid"], protect, cur["feats0"])
                else:
                    args = conversion_inputs(cur["feats"], (c0.sample_end - c0.sample_start) // WINDOW, cur["pitch"], cur["pitchf"], cur["sid"],
                                             protect, cur["feats0"])
                for t in (cur["feats"], cur["feats0"], cur["pitch"], cur["pitchf"], cur["sid"], cur["lengths"]):
                    if t is not None:
                        t.record_stream(comp)
                ctx.append((group, cur["T_items"]))
                yield args + (None, noise)

        def results():
            """waveforms [B, T*upp] per batch, in order: two lanes when the module offers them (Synthesizer.infer_many), else call by call"""
            if hasattr(self.net, "infer_many"):
                for r in self.net.infer_many(items()):
                    yield r[0][:, 0]
                return
            for args in items():
                if args[6] is not None:
                    self.net._noise_override = args[6]
                yield self.net.infer(*args[:5])[0][:, 0]

        for gi, wav in enumerate(results()):
            group, T_items = ctx[gi]
            keeps = [T * self.upp - 2 * t_pad_tgt for T in T_items]
            if min(keeps) <= 0:
                raise ValueError(f"chunk of {min(T_items)} frames is shorter than its 2 x {self.cfg.x_pad} s of context")
            bodies = [wav[b, t_pad_tgt:t_pad_tgt + keeps[b]] for b in range(len(group))]   # [t_pad_tgt : -t_pad_tgt], :444
            for b in range(len(group)):   # one running peak for the whole stream
                _lib.check(self.lib.rvc_absmax_f32(bodies[b].data_ptr(), 0, keeps[b], self._peak.data_ptr(), 1, st), "rvc_absmax_f32")
            ev = torch.cuda.Event()
            ev.record(comp)
            for b, c in enumerate(group):
                short |= keeps[b] != lens[c.index]
                pieces[c.index] = bodies[b]
                if to_host and world == 1:
                    self.d2h.wait_event(ev)
                    with torch.cuda.stream(self.d2h):
                        self._host[offs[c.index]: offs[c.index] + min(keeps[b], lens[c.index])].copy_(bodies[b][:lens[c.index]], non_blocking=True)
            wav.record_stream(self.d2h)
            keep.append(wav)
        self.last_stats = {"chunks": len(spans), "batches": len(groups), "h2d_bytes": h2d_bytes, "d2h_bytes": 4 * total if rank == 0 else 0,
                           "shapes": len(shapes)}
        if not to_host:
            return pieces
        if world > 1:
            return self._gather(pieces, spans, owner, offs, lens, total, rank, world, short)
        self.d2h.synchronize()
        comp.synchronize()
        return self._finish(pieces, spans, offs, lens, total, short, float(self._peak.item()))

    def _finish(self, pieces, spans, offs, lens, total, short, peak):
        out = self._host[:total].numpy()
        if short:   # an embedder that returned fewer frames than the pitch grid: pieces are shorter, stitch what exists (:448)
            out = np.concatenate([out[offs[c.index]: offs[c.index] + min(int(pieces[c.index].shape[0]), lens[c.index])] for c in spans])
        audio_max = np.float32(peak) / 0.99                                            # :451-452 (float32 peak, as np.abs(x).max())
        if audio_max > 1:
            out /= audio_max
        return out

    def _gather(self, pieces, spans, owner, offs, lens, total, rank, world, short):
        peak, total = gather_pieces(pieces, self._peak, spans, owner, offs, lens, rank, world, self._host, self.dev)
        torch.cuda.current_stream(self.dev).synchronize()
        if rank != 0:
            return None
        return self._finish(pieces, spans, offs, lens, total, False, peak)   # (already laid out back to back)


def gather_pieces(pieces, peak, spans, owner, offs, lens, rank, world, host, device):
    """Every rank packs the pieces it owns into one buffer (+ its running peak as the last element); ONE ``dist.gather`` to rank 0
    (NCCL over NVLink for CUDA tensors; the only collective of the stream workload, after the data path); rank 0 copies each piece
    to its offset in ``host`` and returns ``(global peak, stream length)`` (``(None, None)`` elsewhere).  Pieces may be shorter than
    planned (an embedder that returns fewer frames than the pitch grid, e.g. a chunk with an odd number of frames loses one,
    convert.py:365-369): the ranks first agree on the actual lengths (one all-reduce of n_chunks integers) and the pieces are laid out
    back to back, as the reference's ``np.concatenate`` does (:448).  Device-agnostic so that the layout logic is testable under gloo."""
    import torch.distributed as dist

    act = torch.tensor([min(int(pieces[c.index].shape[0]), lens[c.index]) if owner[c.index] == rank else 0 for c in spans],
                       dtype=torch.int64).to(device)
    dist.all_reduce(act)
    lens = [int(v) for v in act.tolist()]
    offs = [0] * len(spans)
    for i in range(1, len(spans)):
        offs[i] = offs[i - 1] + lens[i - 1]
    per_rank = [sum(lens[c.index] for c in spans if owner[c.index] == r) for r in range(world)]
    cap = max(per_rank) + 1
    buf = torch.zeros(cap, device=device)
    o = 0
    for c in spans:
        if owner[c.index] == rank:
            buf[o:o + lens[c.index]].copy_(pieces[c.index][:lens[c.index]])
            o += lens[c.index]
    buf[cap - 1:cap].copy_(peak.reshape(-1)[:1])
    dst = [torch.empty(cap, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, dst, dst=0)
    if rank != 0:
        return None, None
    cursor = [0] * world
    for c in spans:
        r = owner[c.index]
        host[offs[c.index]: offs[c.index] + lens[c.index]].copy_(dst[r][cursor[r]: cursor[r] + lens[c.index]], non_blocking=True)
        cursor[r] += lens[c.index]
    return float(torch.stack([d[cap - 1] for d in dst]).max().item()), sum(lens)
