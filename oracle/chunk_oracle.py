"""Literal numpy restatement of the reference's chunk loop (TEST INFRASTRUCTURE) — main/inference/convert.py:403-448.

`convert.py` itself cannot be imported here (librosa / faiss / onnxruntime are absent), so the chunk plan is pinned by
restating exactly the slicing statements the reference executes, with `voice_conversion` replaced by a caller-supplied
function of (audio slice, pitch-frame slice bounds).  parity unpinned by reference tests: the reference has none.
"""
import numpy as np


def reference_chunk_loop(audio, vc, tgt_sr, x_pad=1, x_query=6, x_center=38, x_max=41):
    sample_rate, window = 16000, 160                                   # :186-187
    t_pad, t_pad_tgt = sample_rate * x_pad, tgt_sr * x_pad             # :188-189
    t_pad2 = t_pad * 2
    t_query, t_center, t_max = sample_rate * x_query, sample_rate * x_center, sample_rate * x_max
    opt_ts, audio_opt = [], []
    audio_pad = np.pad(audio, (window // 2, window // 2), mode="reflect")          # :404
    if audio_pad.shape[0] > t_max:                                                   # :406
        audio_sum = np.zeros_like(audio)
        for i in range(window):
            audio_sum += audio_pad[i: i - window]
        for t in range(t_center, audio.shape[0], t_center):
            opt_ts.append(t - t_query + np.where(np.abs(audio_sum[t - t_query: t + t_query]) == np.abs(audio_sum[t - t_query: t + t_query]).min())[0][0])
    s = 0
    t = None
    audio_pad = np.pad(audio, (t_pad, t_pad), mode="reflect")                        # :416
    p_len = audio_pad.shape[0] // window
    for t in opt_ts:                                                                 # :442-445
        t = t // window * window
        audio_opt.append(vc(audio_pad[s: t + t_pad2 + window], s // window, (t + t_pad2) // window)[t_pad_tgt: -t_pad_tgt])
        s = t
    audio_opt.append(vc(audio_pad[t:], (t // window) if t is not None else 0, p_len)[t_pad_tgt: -t_pad_tgt])   # :447
    return np.concatenate(audio_opt), [int(x) for x in opt_ts]
