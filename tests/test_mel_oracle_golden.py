"""Pin oracle/mel_oracle.py (the RMVPE mel front end) against tests/golden/mel_front.npz, written from the UNMODIFIED reference MelSpectrogram
(oracle/make_mel_golden.py).  Same torch calls on the same host: equal up to the FFT library's rounding (bar 1e-5 on the log-mel)."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle
from oracle.make_mel_golden import CASES


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "mel_front.npz"))


def test_filterbank_is_the_fixture_buffer(gold):
    fb = mel_oracle.mel_filterbank()
    assert fb.dtype == np.float32 and fb.shape == (128, 513) and np.array_equal(fb, gold["mel_basis"])
    # triangular, Slaney-normalised: every band has a contiguous support, neighbours overlap, nothing below fmin or above fmax
    freqs = np.linspace(0, 8000, 513)
    for m in range(128):
        nz = np.nonzero(fb[m])[0]
        assert len(nz) >= 1 and np.array_equal(nz, np.arange(nz[0], nz[-1] + 1))
    assert fb[:, freqs < 30].sum() == 0 and fb[:, -1].sum() < 1e-12          # (the ramp at exactly fmax leaves rounding dust)


@pytest.mark.parametrize("name", sorted(CASES))
def test_mel_oracle_matches_reference(name, gold):
    n, seed = CASES[name]
    audio = mel_oracle.synthetic_audio(n, seed)
    mel = mel_oracle.mel_spectrogram(audio, torch.from_numpy(gold["mel_basis"]))
    g = torch.from_numpy(gold[name])
    assert mel.shape == g.shape == (1, 128, 1 + n // 160)
    assert float((mel - g).abs().max()) <= 1e-5
    assert float(g.min()) == pytest.approx(np.log(1e-5), abs=1e-6) if n > 4000 else True      # the clamp branch is exercised (silent stretch)
    if g.shape[-1] > 17:
        p = mel_oracle.pad_frames(mel)
        T = mel.shape[-1]
        assert p.shape[-1] % 32 == 0 and p.shape[-1] - T < 32 and torch.equal(p[..., T], mel[..., T - 2])
