"""Restatement of the reference's per-chunk feature preparation and stream stitching (TEST INFRASTRUCTURE).

Follows main/inference/convert.py statement by statement with the same torch / numpy calls the reference makes:
``voice_conversion`` from the embedder output on (:361-381) and the tail of ``pipeline`` (:448-452).  `convert.py` cannot be
imported in this container (librosa / faiss / onnxruntime / soundfile are absent) and the reference has no test for it, so this
file is pinned only by being a literal restatement: parity unpinned by reference tests.  Only tests/ and bench.py's
reference leg may import it.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .chunk_oracle import reference_chunk_loop


def reference_retrieval_blend(feats, big_npy, score, ix, index_rate):
    """convert.py:354-359 with (score, ix) = index.search(npy, k=8) given: feats [1, T, D] fp32 torch (CPU), returns the blended feats."""
    weight = np.square(1 / score)                                                                # :354
    weight /= weight.sum(axis=1, keepdims=True)                                                  # :355
    npy = np.sum(big_npy[ix] * np.expand_dims(weight, axis=2), axis=1)                           # :356
    return torch.from_numpy(npy).unsqueeze(0) * index_rate + (1 - index_rate) * feats            # :359


def reference_voice_conversion(infer_fn, feats, feats0, n_audio_samples, pitch, pitchf, sid, protect, window=160):
    """feats [1, Th, D] = embedder output after the retrieval blend, feats0 = before it (None when protect >= 0.5 or no pitch)."""
    pitch_guidance = pitch is not None and pitchf is not None                                   # :329
    feats = F.interpolate(feats.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)              # :361
    if protect < 0.5 and pitch_guidance:
        feats0 = F.interpolate(feats0.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)        # :362
    p_len = n_audio_samples // window                                                           # :363
    if feats.shape[1] < p_len:                                                                  # :365-369
        p_len = feats.shape[1]
        if pitch_guidance:
            pitch = pitch[:, :p_len]
            pitchf = pitchf[:, :p_len]
    if protect < 0.5 and pitch_guidance:                                                        # :372-378
        pitchff = pitchf.clone()
        pitchff[pitchf > 0] = 1
        pitchff[pitchf < 1] = protect
        pitchff = pitchff.unsqueeze(-1)
        feats = (feats * pitchff + feats0 * (1 - pitchff)).to(feats0.dtype)
    p_len = torch.tensor([p_len]).long()                                                        # :380
    out = infer_fn(feats.float(), p_len, pitch if pitch_guidance else None, pitchf.float() if pitch_guidance else None, sid)
    return out[0][0, 0].data.cpu().float().numpy(), feats                                       # :381


def reference_stream(audio, embed_fn, infer_fn, pitch, pitchf, sid, tgt_sr, protect, x_pad=1, x_query=6, x_center=38, x_max=41):
    """VC.pipeline from :403 on: chunk loop (chunk_oracle), concatenation and the peak normalisation (:448-452).
    embed_fn(audio_slice, frame_start, frame_end) -> (feats, feats0); pitch / pitchf: [1, frames] tensors."""
    def vc(audio_slice, f0, f1):
        feats, feats0 = embed_fn(audio_slice, f0, f1)
        return reference_voice_conversion(infer_fn, feats, feats0, audio_slice.shape[0], pitch[:, f0:f1], pitchf[:, f0:f1], sid, protect)[0]

    audio_opt, opt_ts = reference_chunk_loop(audio, vc, tgt_sr, x_pad, x_query, x_center, x_max)
    audio_max = np.abs(audio_opt).max() / 0.99                                                  # :451
    if audio_max > 1:
        audio_opt = audio_opt / audio_max                                                       # :452
    return audio_opt, opt_ts
