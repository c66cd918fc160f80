"""The f0 tail in front of ``Synthesizer.infer``: RMVPE salience -> f0 -> (``pitch``, ``pitchf``).

Mirrors, by name and argument meaning, the reference functions between the RMVPE network and the synthesiser (post-network part):

* ``mel_spectrogram(audio, mel_basis)`` — ``MelSpectrogram.forward`` + the frame padding of ``mel2hidden`` (RMVPE.py:161-181, :212-213): the input of
  the RMVPE network, one launch instead of the reference's pad / FFT / abs / GEMM / clamp / log / pad;
* ``decode(hidden, thred)`` / ``decode_with_pitch`` — ``RMVPE.decode`` / ``to_local_average_cents`` / the legacy gate
  (main/library/predictors/RMVPE.py:217-221, :229-233, :237-252), on a CUDA salience map ``[T, 360]``;
* ``get_f0_tail(f0, pitch, f0_autotune, f0_autotune_strength)`` — the tail of ``VC.get_f0`` (main/inference/convert.py:308-323);
* ``rmvpe_pitch(hidden, ...)`` — both in one launch: what ``VC.pipeline`` (:436-439) hands to ``infer`` (``pitch`` int64, ``pitchf`` float32),
  without the device -> host -> device round trip of the reference (``hidden.cpu().numpy()`` at RMVPE.py:226, ``torch.tensor(..., device=)`` at :439).

The device functions need CUDA tensors and ``librvc_b200.so`` (no CPU fallback).  ``coarse_pitch`` is the host-side numpy quantiser kept for callers
that already hold a host f0 track (other f0 methods); it is the same formula as the reference and bit-exact by construction.
"""
import numpy as np

F0_MIN, F0_MAX = 50.0, 1100.0                                     # convert.py:196-197
# convert.py:200
REF_FREQS = (49.00, 51.91, 55.00, 58.27, 61.74, 65.41, 69.30, 73.42, 77.78, 82.41, 87.31, 92.50, 98.00, 103.83, 110.00, 116.54, 123.47, 130.81,
             138.59, 146.83, 155.56, 164.81, 174.61, 185.00, 196.00, 207.65, 220.00, 233.08, 246.94, 261.63, 277.18, 293.66, 311.13, 329.63,
             349.23, 369.99, 392.00, 415.30, 440.00, 466.16, 493.88, 523.25, 554.37, 587.33, 622.25, 659.25, 698.46, 739.99, 783.99, 830.61,
             880.00, 932.33, 987.77, 1046.50)
N_CLASS = 360


def _mel_limits():
    return float(1127 * np.log(1 + F0_MIN / 700)), float(1127 * np.log(1 + F0_MAX / 700))   # convert.py:198-199


def coarse_pitch(f0):
    """Host-side coarse quantiser (convert.py:318-323): mel scale in float64, affine map to 1..255, ``np.rint`` (half to even)."""
    f0 = np.asarray(f0, dtype=np.float64)
    mel_min, mel_max = _mel_limits()
    mel = 1127.0 * np.log(1.0 + f0 / 700.0)
    voiced = mel > 0
    mel[voiced] = (mel[voiced] - mel_min) * 254.0 / (mel_max - mel_min) + 1.0
    mel[mel <= 1] = 1
    mel[mel > 255] = 255
    return np.rint(mel).astype(np.int32)


def _cuda(t, what, dtype):
    import torch

    if not isinstance(t, torch.Tensor) or t.device.type != "cuda":
        raise RuntimeError(f"rvc_maker_b200.pitch.{what} needs a CUDA tensor (no CPU fallback)")
    return t.to(dtype)


_notes_cache = {}


def _notes(device, note_dict):
    import torch

    key = (str(device), tuple(note_dict))
    if key not in _notes_cache:
        _notes_cache[key] = torch.tensor(list(note_dict), dtype=torch.float64, device=device)
    return _notes_cache[key]


def _salience(hidden, what):
    import torch

    hidden = _cuda(hidden, what, torch.float32)
    if hidden.dim() == 3 and hidden.shape[0] == 1:                # the network's [1, T, 360] (RMVPE.py:226 squeezes it)
        hidden = hidden[0]
    if hidden.dim() != 2 or hidden.shape[1] != N_CLASS:
        raise ValueError(f"salience must be [T, {N_CLASS}], got {tuple(hidden.shape)}")
    if hidden.stride(1) != 1:
        hidden = hidden.contiguous()
    return hidden


_window_cache = {}
_ranges_cache = {}


def mel_spectrogram(audio, mel_basis, clamp=1e-5, pad_frames=True, sparse=True):
    """``MelSpectrogram.forward(audio, center=True)`` (keyshift 0, speed 1, fp32; RMVPE.py:161-181) in one launch, with — when ``pad_frames`` — the
    reflect padding of the frame axis to a multiple of 32 that ``mel2hidden`` applies before the network (:212-213).

    audio: [n] or [1, n] fp32 CUDA tensor (16 kHz); mel_basis: the reference module's ``mel_basis`` buffer [128, 513] on the same device.
    ``sparse``: multiply only the non-zero bins of every (triangular) filterbank row — same values, the zeros contribute nothing.
    Returns ``(mel [1, 128, Tp], n_frames)``: feed ``mel`` to the network and cut its output to ``[:, :n_frames]`` as ``mel2hidden`` does (:214)."""
    import torch

    from . import _lib

    audio = _cuda(audio, "mel_spectrogram", torch.float32).reshape(-1).contiguous()
    mel_basis = _cuda(mel_basis, "mel_spectrogram", torch.float32).contiguous()
    if tuple(mel_basis.shape) != (128, 513):
        raise ValueError(f"mel_basis must be [128, 513] (RMVPE: 128 mels, n_fft 1024), got {tuple(mel_basis.shape)}")
    dev = audio.device
    if str(dev) not in _window_cache:
        _window_cache[str(dev)] = torch.hann_window(1024, device=dev)      # RMVPE.py:165
    n = audio.numel()
    T = 1 + n // 160
    Tp = 32 * ((T - 1) // 32 + 1) if pad_frames else T
    out = torch.empty(1, 128, Tp, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    lib = _lib.load()
    rng = None
    if sparse:
        # the non-zero support of every filterbank row, computed once per buffer (keyed on storage and in-place version)
        key = (mel_basis.data_ptr(), mel_basis._version, str(dev))
        if _ranges_cache.get("key") != key:
            r = torch.empty(128, 2, dtype=torch.int32, device=dev)
            _lib.check(lib.rvc_mel_basis_ranges(mel_basis.data_ptr(), r.data_ptr(), st), "rvc_mel_basis_ranges")
            _ranges_cache.update(key=key, ranges=r, basis=mel_basis)       # (holding the tensor keeps its address from being reused)
        rng = _ranges_cache["ranges"]
    _lib.check(lib.rvc_mel_spectrogram_f32(audio.data_ptr(), n, _window_cache[str(dev)].data_ptr(), mel_basis.data_ptr(),
                                           rng.data_ptr() if rng is not None else None, out.data_ptr(), Tp, T, Tp, float(clamp), st),
               "rvc_mel_spectrogram_f32")
    return out, T


def decode(hidden, thred=0.03, f0_min=None, f0_max=None):
    """``RMVPE.decode`` on the device: salience [T, 360] (or [1, T, 360]) -> f0 [T] float64 CUDA tensor (0 = unvoiced)."""
    import torch

    from . import _lib

    hidden = _salience(hidden, "decode")
    T = hidden.shape[0]
    f0 = torch.empty(T, dtype=torch.float64, device=hidden.device)
    lo, hi = (0.0, 0.0) if f0_max is None else (float(f0_min), float(f0_max))
    st = torch.cuda.current_stream(hidden.device).cuda_stream
    _lib.check(_lib.load().rvc_rmvpe_decode_f32(hidden.data_ptr(), hidden.stride(0), T, float(thred), lo, hi, f0.data_ptr(), st), "rvc_rmvpe_decode_f32")
    return f0


def decode_with_pitch(hidden, thred=0.03, f0_min=50, f0_max=1100):
    """The tail of ``RMVPE.infer_from_audio_with_pitch`` (the "rmvpe-legacy" method): ``decode`` + the [f0_min, f0_max] gate."""
    return decode(hidden, thred, f0_min, f0_max)


def get_f0_tail(f0, pitch=0, f0_autotune=False, f0_autotune_strength=1.0, note_dict=REF_FREQS, return_f0=False):
    """The tail of ``VC.get_f0`` on a float64 CUDA f0 track [n]: returns (``pitch`` int64 [n], ``pitchf`` float32 [n]) (+ the shifted float64 f0)."""
    import torch

    from . import _lib

    f0 = _cuda(f0, "get_f0_tail", torch.float64).contiguous()
    n = f0.numel()
    dev = f0.device
    coarse = torch.empty(f0.shape, dtype=torch.int64, device=dev)
    pitchf = torch.empty(f0.shape, dtype=torch.float32, device=dev)
    f0s = torch.empty_like(f0) if return_f0 else None
    notes = _notes(dev, note_dict) if f0_autotune else None
    mel_min, mel_max = _mel_limits()
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.load().rvc_f0_coarse_f64(f0.data_ptr(), n, notes.data_ptr() if f0_autotune else None, len(note_dict) if f0_autotune else 0,
                                             float(f0_autotune_strength), float(pow(2, pitch / 12)), mel_min, mel_max, coarse.data_ptr(),
                                             pitchf.data_ptr(), f0s.data_ptr() if return_f0 else None, st), "rvc_f0_coarse_f64")
    return (coarse, pitchf, f0s) if return_f0 else (coarse, pitchf)


def rmvpe_pitch(hidden, pitch=0, thred=0.03, f0_autotune=False, f0_autotune_strength=1.0, legacy=False, note_dict=REF_FREQS, return_f0=False):
    """salience -> (``pitch`` int64 [T], ``pitchf`` float32 [T]) in one launch: ``get_f0("rmvpe" | "rmvpe-legacy")`` from the network output on."""
    import torch

    from . import _lib

    hidden = _salience(hidden, "rmvpe_pitch")
    T, dev = hidden.shape[0], hidden.device
    coarse = torch.empty(T, dtype=torch.int64, device=dev)
    pitchf = torch.empty(T, dtype=torch.float32, device=dev)
    f0s = torch.empty(T, dtype=torch.float64, device=dev) if return_f0 else None
    notes = _notes(dev, note_dict) if f0_autotune else None
    mel_min, mel_max = _mel_limits()
    lo, hi = (F0_MIN, F0_MAX) if legacy else (0.0, 0.0)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.load().rvc_rmvpe_pitch_f32(hidden.data_ptr(), hidden.stride(0), T, float(thred), lo, hi,
                                               notes.data_ptr() if f0_autotune else None, len(note_dict) if f0_autotune else 0,
                                               float(f0_autotune_strength), float(pow(2, pitch / 12)), mel_min, mel_max, coarse.data_ptr(),
                                               pitchf.data_ptr(), f0s.data_ptr() if return_f0 else None, st), "rvc_rmvpe_pitch_f32")
    return (coarse, pitchf, f0s) if return_f0 else (coarse, pitchf)
