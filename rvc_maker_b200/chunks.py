"""Host-side chunk planner and multi-GPU sharding around ``Synthesizer.infer``.

The reference converts long audio as a sequential loop over independent chunks (main/inference/convert.py:388-458): split
points every ``x_center`` seconds moved to the quietest sample within +-``x_query`` s (:406-412), every chunk extended by
``x_pad`` seconds of context on both sides (:416, :442-447) and the context trimmed from the synthesised waveform (:444).
Chunks are independent given that plan, so they are the data-parallel unit: one process per GPU takes a subset of chunks
(longest first), nothing is exchanged on the data path, and the waveforms are concatenated on the host (:448).

Constants: ``(x_pad, x_query, x_center, x_max) = (1, 6, 38, 41)`` is the reference's fp32 default (main/configs/config.py:83).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SAMPLE_RATE = 16000   # convert.py:186
WINDOW = 160          # convert.py:187 -> 100 feature frames per second


@dataclass(frozen=True)
class ChunkConfig:
    x_pad: int = 1
    x_query: int = 6
    x_center: int = 38
    x_max: int = 41

    @property
    def t_pad(self):
        return SAMPLE_RATE * self.x_pad

    @property
    def t_query(self):
        return SAMPLE_RATE * self.x_query

    @property
    def t_center(self):
        return SAMPLE_RATE * self.x_center

    @property
    def t_max(self):
        return SAMPLE_RATE * self.x_max


def split_points(audio, cfg=ChunkConfig()):
    """Split points ``opt_ts`` in samples of the (unpadded) 16 kHz audio — convert.py:403-412."""
    audio = np.asarray(audio)
    opt_ts = []
    audio_pad = np.pad(audio, (WINDOW // 2, WINDOW // 2), mode="reflect")
    if audio_pad.shape[0] > cfg.t_max:
        audio_sum = np.zeros_like(audio)
        for i in range(WINDOW):
            audio_sum += audio_pad[i: i - WINDOW]
        for t in range(cfg.t_center, audio.shape[0], cfg.t_center):
            seg = np.abs(audio_sum[t - cfg.t_query: t + cfg.t_query])
            opt_ts.append(t - cfg.t_query + int(np.where(seg == seg.min())[0][0]))
    return opt_ts


def uniform_split_points(n_samples, cfg=ChunkConfig()):
    """Split points of a stream with no energy minimum to look for (synthetic benchmark stream): exactly every x_center s."""
    if n_samples + WINDOW <= cfg.t_max:
        return []
    return list(range(cfg.t_center, n_samples, cfg.t_center))


@dataclass(frozen=True)
class Chunk:
    index: int
    frame_start: int   # first feature / pitch frame (100 fps) in the padded stream
    frame_end: int     # one past the last frame
    sample_start: int  # slice of the reflect-padded 16 kHz audio (for the feature extractor)
    sample_end: int

    @property
    def frames(self):
        return self.frame_end - self.frame_start


def chunk_spans(n_samples, opt_ts, cfg=ChunkConfig()):
    """Chunks over the reflect-padded stream (``t_pad`` samples each side) — the slicing of convert.py:441-447."""
    total = n_samples + 2 * cfg.t_pad
    t_pad2 = 2 * cfg.t_pad
    out, s, t = [], 0, None
    for t in opt_ts:
        t = t // WINDOW * WINDOW
        out.append(Chunk(len(out), s // WINDOW, (t + t_pad2) // WINDOW, s, min(t + t_pad2 + WINDOW, total)))
        s = t
    start = t if t is not None else 0
    out.append(Chunk(len(out), start // WINDOW, total // WINDOW, start, total))
    return out


def shard_chunks(chunks, world_size):
    """Longest-processing-time-first assignment of chunks to ranks; deterministic, so every rank computes the same map."""
    load = [0] * world_size
    owner = {}
    for c in sorted(chunks, key=lambda c: (-c.frames, c.index)):
        r = min(range(world_size), key=lambda i: (load[i], i))
        owner[c.index] = r
        load[r] += c.frames
    return owner


def trim(wave, tgt_sr, cfg=ChunkConfig()):
    """Drop the x_pad seconds of context from both ends of a synthesised chunk — ``[t_pad_tgt : -t_pad_tgt]`` (:444)."""
    t = tgt_sr * cfg.x_pad
    return wave[t:-t] if t > 0 else wave


def synthesize_chunks(infer_fn, chunks, owner, rank, tgt_sr, cfg=ChunkConfig()):
    """Run ``infer_fn(chunk) -> 1-D waveform`` for the chunks this rank owns; returns {chunk index: trimmed waveform}."""
    return {c.index: trim(infer_fn(c), tgt_sr, cfg) for c in chunks if owner[c.index] == rank}


def gather_stream(parts, n_chunks, rank=0, world_size=1, dst=0):
    """Host-side gather + concatenation (convert.py:448).  With world_size > 1 uses torch.distributed object gather
    (no device collective on the data path); returns the full waveform on ``dst`` and None elsewhere."""
    if world_size > 1:
        import torch.distributed as dist

        gathered = [None] * world_size if rank == dst else None
        dist.gather_object(parts, gathered, dst=dst)
        if rank != dst:
            return None
        parts = {}
        for g in gathered:
            parts.update(g)
    missing = [i for i in range(n_chunks) if i not in parts]
    if missing:
        raise RuntimeError(f"chunks {missing} were not synthesised by any rank")
    return np.concatenate([np.asarray(parts[i]) for i in range(n_chunks)])
