"""The RMVPE mel front end on the device (csrc/mel_front.cu through rvc_maker_b200.pitch.mel_spectrogram) against the reference's golden output and
the oracle.  Floating point, fp32 like the reference; the FFT and the filterbank product sum in a different order than the reference's libraries:
bar |log-mel - ref| <= 2e-4 wherever the mel energy is above 100x the clamp (relative error of a sum of 513 fp32 terms), and <= 2e-2 below it, where
the band energy is rounding noise of the FFT itself (cuFFT and pocketfft differ there by as much); the clamp floor must be hit exactly where the
reference hits it in silent frames."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle
from oracle.make_mel_golden import CASES

pytestmark = pytest.mark.gpu
LOG_CLAMP = float(np.log(np.float32(1e-5)))


def check(got, ref):
    assert got.shape == ref.shape
    loud = ref > LOG_CLAMP + np.log(100.0)
    d = (got - ref).abs()
    assert float(d[loud].max()) <= 2e-4, float(d[loud].max())
    assert float(d.max()) <= 2e-2, float(d.max())
    return float(d[loud].max()), float(d.max())


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "mel_front.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_mel_matches_reference_golden(name, gold):
    from rvc_maker_b200 import pitch as P

    n, seed = CASES[name]
    audio = mel_oracle.synthetic_audio(n, seed).cuda()
    basis = torch.from_numpy(gold["mel_basis"]).cuda()
    mel, T = P.mel_spectrogram(audio, basis, pad_frames=False)
    ref = torch.from_numpy(gold[name])
    assert T == ref.shape[-1]
    e = check(mel.cpu(), ref)
    silent = ref == LOG_CLAMP                      # all-zero windows: the reference sits exactly on the clamp
    assert torch.equal(mel.cpu()[silent], ref[silent])
    print(f"[mel {name}] max err loud / all: {e[0]:.2e} / {e[1]:.2e}; clamped bins {int(silent.sum())}")
    dense, _ = P.mel_spectrogram(audio, basis, pad_frames=False, sparse=False)     # the dense filterbank product: same terms, another order
    check(dense.cpu(), ref)
    assert float((dense - mel).abs().max()) <= 1e-4 and torch.equal(dense.cpu()[silent], ref[silent])
    if T > 17:                                     # mel2hidden's reflect padding of the frame axis (the reference raises for shorter clips)
        melp, T2 = P.mel_spectrogram(audio[0], basis)
        assert T2 == T and melp.shape[-1] % 32 == 0 and 0 <= melp.shape[-1] - T < 32
        assert torch.equal(melp[..., :T], mel)
        assert torch.equal(melp.cpu(), mel_oracle.pad_frames(mel.cpu()))


def test_mel_long_audio_against_oracle():
    """602 s of audio (BASELINE configs[4]'s ten-minute stream with its padding: 60 201 frames) against the torch CPU restatement."""
    from rvc_maker_b200 import pitch as P

    audio = mel_oracle.synthetic_audio(602 * 16000, 7)
    basis = torch.from_numpy(mel_oracle.mel_filterbank())
    ref = mel_oracle.pad_frames(mel_oracle.mel_spectrogram(audio, basis))
    mel, T = P.mel_spectrogram(audio.cuda(), basis.cuda())
    assert T == 60201 and mel.shape == ref.shape
    check(mel.cpu(), ref)


def test_mel_argument_errors():
    from rvc_maker_b200 import pitch as P

    basis = torch.zeros(128, 513, device="cuda")
    with pytest.raises(RuntimeError, match="512 samples"):
        P.mel_spectrogram(torch.zeros(400, device="cuda"), basis)
    with pytest.raises(RuntimeError, match="Tp - T < T"):
        P.mel_spectrogram(torch.zeros(700, device="cuda"), basis)          # 5 frames cannot be reflect-padded to 32 (the reference raises too)
    with pytest.raises(ValueError):
        P.mel_spectrogram(torch.zeros(16000, device="cuda"), basis[:, :512])
    with pytest.raises(RuntimeError, match="CUDA"):
        P.mel_spectrogram(torch.zeros(16000), basis)
