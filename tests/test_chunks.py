"""Chunk planner / sharding host logic (CPU): plan parity with the restated reference loop, and a world_size-2 gloo run."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.chunk_oracle import reference_chunk_loop
from rvc_maker_b200 import chunks as ch

TGT_SR = 48000


def fake_vc_from_frames(f0, f1, tgt_sr=TGT_SR):
    """Deterministic stand-in for voice_conversion: one output sample block per input frame."""
    n = (f1 - f0) * (tgt_sr // 100)
    return (np.arange(n, dtype=np.float64) * 1e-6 + f0).astype(np.float32)


def make_audio(seconds, seed=0):
    rng = np.random.default_rng(seed)
    n = int(seconds * 16000)
    env = 0.5 + 0.5 * np.sin(np.arange(n) / 16000.0 * 2 * np.pi / 7.3)
    return (rng.standard_normal(n) * env * 0.1).astype(np.float64)


@pytest.mark.parametrize("seconds", [3.0, 40.9, 41.5, 100.0, 200.3])
def test_plan_matches_reference_loop(seconds):
    audio = make_audio(seconds)
    ref_wave, ref_ts = reference_chunk_loop(audio, lambda a, f0, f1: fake_vc_from_frames(f0, f1), TGT_SR)
    cfg = ch.ChunkConfig()
    ts = ch.split_points(audio, cfg)
    assert ts == ref_ts
    spans = ch.chunk_spans(audio.shape[0], ts, cfg)
    owner = ch.shard_chunks(spans, 1)
    parts = ch.synthesize_chunks(lambda c: fake_vc_from_frames(c.frame_start, c.frame_end), spans, owner, 0, TGT_SR, cfg)
    wave = ch.gather_stream(parts, len(spans))
    assert wave.shape == ref_wave.shape and np.array_equal(wave, ref_wave)


def test_uniform_plan_config5():
    """BASELINE config 5: 600 s stream, split exactly every 38 s, 1 s pad -> 15 chunks of 4000 frames + one of 3200."""
    n = 600 * 16000
    spans = ch.chunk_spans(n, ch.uniform_split_points(n))
    assert len(spans) == 16
    assert [c.frames for c in spans] == [4000] * 15 + [3200]
    assert ch.uniform_split_points(30 * 16000) == []


def test_sharding_is_balanced_and_complete():
    n = 600 * 16000
    spans = ch.chunk_spans(n, ch.uniform_split_points(n))
    for world in (1, 2, 4, 8):
        owner = ch.shard_chunks(spans, world)
        assert sorted(owner) == list(range(len(spans)))
        loads = [sum(c.frames for c in spans if owner[c.index] == r) for r in range(world)]
        assert max(loads) - min(loads) <= 4000
    with pytest.raises(RuntimeError):
        ch.gather_stream({0: np.zeros(3)}, 2)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    audio = make_audio(200.3)
    cfg = ch.ChunkConfig()
    spans = ch.chunk_spans(audio.shape[0], ch.split_points(audio, cfg), cfg)
    owner = ch.shard_chunks(spans, world)
    parts = ch.synthesize_chunks(lambda c: fake_vc_from_frames(c.frame_start, c.frame_end), spans, owner, rank, TGT_SR, cfg)
    assert all(owner[i] == rank for i in parts)
    wave = ch.gather_stream(parts, len(spans), rank, world)
    if rank == 0:
        np.save(out_path, wave)
    else:
        assert wave is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_stream_equals_single_process(tmp_path):
    out = str(tmp_path / "wave.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    audio = make_audio(200.3)
    ref_wave, _ = reference_chunk_loop(audio, lambda a, f0, f1: fake_vc_from_frames(f0, f1), TGT_SR)
    assert np.array_equal(np.load(out), ref_wave)
