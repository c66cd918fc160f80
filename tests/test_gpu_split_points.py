"""pipeline.split_points (csrc/split_search.cu) against the literal restatement of convert.py:404-412 (oracle/chunk_oracle.py) and the host planner
(chunks.split_points).  Index work: the bar is bit-exact — the sums are formed in numpy's order and dtype."""
import numpy as np
import pytest
import torch
from scipy import signal

from oracle import chunk_oracle
from rvc_maker_b200 import chunks as ch
from rvc_maker_b200 import pipeline as pl

pytestmark = pytest.mark.gpu


def reference_opt_ts(audio, **kw):
    return chunk_oracle.reference_chunk_loop(audio, lambda a, f0, f1: np.zeros(3 * 48000, np.float32), 48000, **kw)[1]


def make_audio(seconds, seed, dtype):
    rng = np.random.Generator(np.random.PCG64(seed))
    n = seconds * ch.SAMPLE_RATE + int(rng.integers(0, 999))
    t = np.arange(n) / ch.SAMPLE_RATE
    env = 0.55 + 0.45 * np.sin(2 * np.pi * t / 7.3 + rng.random())          # slow loudness swings: real minima to find
    x = (env * (0.3 * np.sin(2 * np.pi * 220 * t) + 0.05 * rng.standard_normal(n))).astype(np.float32)
    if dtype == np.float64:                                                   # the reference's own path: high-pass filtered, float64 (:402)
        bh, ah = signal.butter(N=5, Wn=48, btype="high", fs=16000)            # convert.py:29
        x = np.ascontiguousarray(signal.filtfilt(bh, ah, x))                   # (filtfilt hands back a reversed view)
    return x


@pytest.mark.parametrize("seconds,seed,dtype", [(50, 1, np.float64), (100, 2, np.float64), (100, 3, np.float32), (200, 4, np.float64), (41, 5, np.float32),
                                                (77, 6, np.float64), (600, 7, np.float64)])
def test_split_points_match_reference(seconds, seed, dtype):
    audio = make_audio(seconds, seed, dtype)
    want = reference_opt_ts(audio)
    got = pl.split_points(torch.from_numpy(audio).cuda())
    assert got == want and len(want) == (audio.shape[0] - 1) // (38 * 16000)
    assert ch.split_points(audio) == want                                     # the host planner agrees too


def test_split_points_edges():
    # silence: every |sum| is 0 -> the first sample of each window (np.where(...)[0][0])
    z = np.zeros(100 * 16000, np.float32)
    assert pl.split_points(torch.from_numpy(z).cuda()) == reference_opt_ts(z) == [32 * 16000, 70 * 16000]
    # signed zeros and exact ties between distant samples
    x = make_audio(90, 9, np.float64)
    x[40 * 16000: 40 * 16000 + 400] = 0.0
    x[35 * 16000: 35 * 16000 + 400] = -0.0
    assert pl.split_points(torch.from_numpy(x).cuda()) == reference_opt_ts(x)
    # the last window clipped by the end of the signal (t + t_query > n), a candidate whose window holds a single block
    for secs in (77, 115):
        y = make_audio(secs, 10 + secs, np.float64)[: secs * 16000]
        assert pl.split_points(torch.from_numpy(y).cuda()) == reference_opt_ts(y)
    # below the reference's guard (audio_pad.shape[0] > t_max): no split
    s = make_audio(40, 11, np.float32)[: 41 * 16000 - 160]
    assert pl.split_points(torch.from_numpy(s).cuda()) == reference_opt_ts(s) == []
    # other chunk constants (the reference's low-memory configuration x_query 5, x_center 30, x_max 32: config.py:81)
    cfg = ch.ChunkConfig(x_pad=1, x_query=5, x_center=30, x_max=32)
    w = make_audio(100, 12, np.float64)
    assert pl.split_points(torch.from_numpy(w).cuda(), cfg) == reference_opt_ts(w, x_pad=1, x_query=5, x_center=30, x_max=32)
    with pytest.raises(RuntimeError, match="CUDA"):
        pl.split_points(torch.from_numpy(w))


def test_split_points_feed_the_converter():
    """The device-made plan drives StreamConverter to the same chunks as the host-made one."""
    audio = make_audio(100, 13, np.float64)
    assert ch.chunk_spans(audio.shape[0], pl.split_points(torch.from_numpy(audio).cuda())) == ch.chunk_spans(audio.shape[0], ch.split_points(audio))
