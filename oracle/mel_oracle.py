"""Restatement of the RMVPE mel front end (TEST INFRASTRUCTURE): ``MelSpectrogram.forward`` (main/library/predictors/RMVPE.py:141-181, the
``keyshift == 0``, ``speed == 1`` branch that ``infer_from_audio`` takes, :224) and the frame padding of ``mel2hidden`` (:211-214).

The filterbank is a BUFFER of the reference module (``register_buffer("mel_basis", ...)``, :146-148) that it fills from
``librosa.filters.mel(sr=16000, n_fft=1024, n_mels=128, fmin=30, fmax=8000, htk=True)``; librosa (pinned ``librosa>=0.10`` by the reference's
requirements.txt) is absent from this container, so ``mel_filterbank`` restates librosa's published algorithm (triangular filters on the HTK mel
scale, Slaney area normalisation) — *parity unpinned* for the filterbank itself.  The device kernel takes the filterbank as an input, like a weight, so
its parity does not depend on this restatement.  ``mel_spectrogram`` is pinned by tests/golden/mel_front.npz, written by oracle/make_mel_golden.py
from the UNMODIFIED reference class (RMVPE.py imported with ``librosa.filters.mel`` stubbed by ``mel_filterbank``).  Only tests/ and bench.py's
cpu_baseline leg may import this file.
"""
import numpy as np
import torch
import torch.nn.functional as F

N_MELS, SR, N_FFT, HOP, FMIN, FMAX, CLAMP = 128, 16000, 1024, 160, 30, 8000, 1e-5     # RMVPE.py:7, :204


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX, htk=True, **_):
    """librosa.filters.mel (0.10) with ``htk=True``, ``norm="slaney"``, float32: restated from the library's documentation."""
    assert htk
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    hz_to_mel = lambda f: 2595.0 * np.log10(1.0 + np.asanyarray(f) / 700.0)
    mel_to_hz = lambda m: 700.0 * (10.0 ** (np.asanyarray(m) / 2595.0) - 1.0)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def mel_spectrogram(audio, mel_basis, n_fft=N_FFT, hop=HOP, clamp=CLAMP):
    """RMVPE.py:161-181 for keyshift = 0, speed = 1, center = True, fp32: audio [1, n] -> log-mel [1, 128, 1 + n // hop]."""
    window = torch.hann_window(n_fft).to(audio.device)                                                                     # :165
    fft = torch.stft(audio, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=window, center=True, return_complex=True)   # :167
    magnitude = torch.sqrt(fft.real.pow(2) + fft.imag.pow(2))                                                # :168
    mel_output = torch.matmul(mel_basis, magnitude)                                                          # :178
    return torch.log(torch.clamp(mel_output, min=clamp))                                                     # :181


def pad_frames(mel):
    """mel2hidden's padding (RMVPE.py:212-213): the frame axis reflect-padded to a multiple of 32."""
    n_frames = mel.shape[-1]
    return F.pad(mel, (0, 32 * ((n_frames - 1) // 32 + 1) - n_frames), mode="reflect")


def synthetic_audio(n, seed):
    """Seeded 16 kHz test signal in [-1, 1]: a gliding harmonic tone + noise, a silent stretch (the clamp branch), a full-scale click."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / SR
    f0 = 110.0 * 2.0 ** (np.sin(2 * np.pi * 0.3 * t) + 0.5 * rng.random())
    ph = 2 * np.pi * np.cumsum(f0) / SR
    x = 0.4 * np.sin(ph) + 0.2 * np.sin(2 * ph) + 0.1 * np.sin(3 * ph) + 0.02 * rng.standard_normal(n)
    if n > 4000:
        x[n // 3: n // 3 + min(n // 5, 2400)] = 0.0
        x[n // 2] = 1.0
    return torch.from_numpy(x.astype(np.float32))[None]
