"""Host logic of the chunk pipeline (no GPU): batching, piece layout, and the oracle restatement of convert.py:361-378."""
import numpy as np
import pytest
import torch

from oracle import pipeline_oracle
from rvc_maker_b200 import chunks as ch
from rvc_maker_b200 import pipeline as pl


def test_batch_chunks_groups_equal_shapes():
    n = 600 * ch.SAMPLE_RATE
    spans = ch.chunk_spans(n, ch.uniform_split_points(n))
    groups = pl.batch_chunks(spans, 2)
    assert [len(g) for g in groups] == [2] * 7 + [1, 1]
    assert all(len({c.frames for c in g}) == 1 for g in groups)
    assert sorted(c.index for g in groups for c in g) == list(range(16))
    assert [len(g) for g in pl.batch_chunks(spans, 1)] == [1] * 16
    assert [len(g) for g in pl.batch_chunks(spans, 4)] == [4, 4, 4, 3, 1]
    assert pl.batch_chunks([], 2) == []


def test_batch_chunks_buckets_energy_split_plans():
    """The reference moves every split point to the quietest sample within +-6 s (convert.py:406-412): on real audio no two chunks
    share a length.  Buckets of 512 frames bring a 10-minute stream back to a handful of shapes and let unequal chunks share a batch."""
    rng = np.random.Generator(np.random.PCG64(5))
    n = 600 * ch.SAMPLE_RATE
    audio = (rng.standard_normal(n) * 0.1).astype(np.float32)
    spans = ch.chunk_spans(n, ch.split_points(audio))
    assert len({c.frames for c in spans}) >= len(spans) - 2
    exact = pl.batch_chunks(spans, 2)
    assert len(exact) >= len(spans) - 2
    groups = pl.batch_chunks(spans, 2, bucket=512)
    assert sorted(c.index for g in groups for c in g) == list(range(len(spans)))
    assert all(len({-(-c.frames // 512) for c in g}) == 1 and len(g) <= 2 for g in groups)
    assert len(groups) < len(exact)
    shapes = {(len(g), -(-g[0].frames // 512) * 512) for g in groups}
    assert len(shapes) <= 8        # fits the engine's plan cache (8 most recently used shapes)
    pad = sum(-(-g[0].frames // 512) * 512 * len(g) for g in groups) / sum(c.frames for c in spans)
    assert pad < 1.12              # padded work stays within 12 % of the real frames


def test_stream_converter_rejects_odd_bucket():
    with pytest.raises((ValueError, RuntimeError)):
        pl.StreamConverter(torch.nn.Linear(1, 1), 48000, 480, bucket=3)


def test_piece_layout_covers_the_stream():
    for seconds in (10, 45, 600):
        n = seconds * ch.SAMPLE_RATE
        spans = ch.chunk_spans(n, ch.uniform_split_points(n))
        offs, lens, total = pl.piece_layout(spans, 480, 48000)
        assert total == seconds * 48000
        assert offs[0] == 0 and all(offs[i] + lens[i] == offs[i + 1] for i in range(len(spans) - 1))


def test_reference_feature_side_semantics():
    """The restated statements: nearest x2, p_len clamp, protect only where pitchf < 1, untouched when protect >= 0.5."""
    g = torch.Generator().manual_seed(0)
    feats, feats0 = torch.randn(1, 50, 8, generator=g), torch.randn(1, 50, 8, generator=g)
    pitchf = torch.tensor([[0.0, 0.5, 1.0, 220.0] * 26])[:, :101]
    seen = {}

    def infer(ph, ln, pi, pf, sid):
        seen.update(phone=ph, lengths=ln, pitch=pi, pitchf=pf)
        return (torch.zeros(1, 1, 4),)

    pipeline_oracle.reference_voice_conversion(infer, feats, feats0, 101 * 160, torch.ones(1, 101).long(), pitchf, None, 0.33)
    ph = seen["phone"]
    assert ph.shape == (1, 100, 8) and int(seen["lengths"]) == 100 and seen["pitch"].shape == (1, 100)
    up, up0 = feats.repeat_interleave(2, 1), feats0.repeat_interleave(2, 1)
    voiced = pitchf[0, :100] >= 1
    assert torch.equal(ph[0, voiced], up[0, voiced] * 1 + up0[0, voiced] * 0)
    assert torch.equal(ph[0, ~voiced], up[0, ~voiced] * torch.tensor(0.33) + up0[0, ~voiced] * (1 - torch.tensor(0.33)))
    pipeline_oracle.reference_voice_conversion(infer, feats, feats0, 101 * 160, torch.ones(1, 101).long(), pitchf, None, 0.5)
    assert torch.equal(seen["phone"], up)


def test_stream_converter_refuses_cpu_model():
    net = torch.nn.Linear(2, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        pl.StreamConverter(net, 48000, 480)
    with pytest.raises(RuntimeError, match="CUDA"):
        pl.feature_prep(torch.zeros(1, 4, 8), None, None, 0.5)


# ---------------------------------------------------------------------------------------------- 2 ranks over gloo
import os  # noqa: E402

import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _piece(c, n):
    return torch.arange(n, dtype=torch.float32) * 1e-3 + c.index


def _cut(c):
    """Pieces 1 and 4 come out one frame short (an odd chunk through a 50 fps embedder): the stitched stream closes the gaps."""
    return 480 if c.index in (1, 4) else 0


def _gather_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 300 * ch.SAMPLE_RATE
    spans = ch.chunk_spans(n, ch.uniform_split_points(n))
    owner = ch.shard_chunks(spans, world)
    offs, lens, total = pl.piece_layout(spans, 480, 48000)
    pieces = {c.index: _piece(c, lens[c.index] - _cut(c)) for c in spans if owner[c.index] == rank}
    peak = torch.tensor([float(max(p.abs().max() for p in pieces.values()))])
    host = torch.zeros(total)
    got_peak, got_total = pl.gather_pieces(pieces, peak, spans, owner, offs, lens, rank, world, host, "cpu")
    if rank == 0:
        assert got_total == total - sum(_cut(c) for c in spans)
        np.save(out_path, np.concatenate([host.numpy()[:got_total], [got_peak]]))
    else:
        assert got_peak is None and got_total is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_piece_gather(tmp_path):
    """The stream workload's only exchange step: one gather of every rank's packed pieces, laid out by chunk offsets on rank 0."""
    out = str(tmp_path / "stream.npy")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_gather_worker, args=(2, port, out), nprocs=2, join=True)
    n = 300 * ch.SAMPLE_RATE
    spans = ch.chunk_spans(n, ch.uniform_split_points(n))
    offs, lens, total = pl.piece_layout(spans, 480, 48000)
    want = torch.cat([_piece(c, lens[c.index] - _cut(c)) for c in spans]).numpy()
    got = np.load(out)
    assert np.array_equal(got[:-1].astype(np.float32), want)
    assert got[-1] == np.float32(want.max())


def test_filter_constants_match_scipy():
    """Host side of pipeline.filtfilt: zi is scipy.signal.lfilter_zi to the bit, and M^block's columns are the states scipy's own lfilter reaches from
    the unit states after `block` zero-input steps (bit-identical: the same recurrence, run rather than exponentiated)."""
    import numpy as np
    from scipy import signal

    from rvc_maker_b200 import pipeline as pl

    for order, wn, btype in ((5, 48, "high"), (3, 200, "high"), (8, 4000, "low")):
        b, a = signal.butter(N=order, Wn=wn, btype=btype, fs=16000)
        b64, a64, zi, ML = pl._filter_constants(b, a, 4096)
        assert np.array_equal(zi, signal.lfilter_zi(b, a)) and ML.shape == (order, order)
        for j in range(order):
            z0 = np.zeros(order)
            z0[j] = 1.0
            _, zf = signal.lfilter(b, a, np.zeros(4096), zi=z0)
            assert np.array_equal(zf, ML[:, j]), (order, j)
    with pytest.raises(ValueError):
        pl._filter_constants([1.0, 2.0], [1.0], 4096)
