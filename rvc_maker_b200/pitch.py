"""Host-side coarse-pitch quantiser — the integer `pitch` input of Synthesizer.infer.

Mirrors VC.get_f0's tail (main/inference/convert.py:318-323, f0_min=50, f0_max=1100 at :196-199): mel scale in float64,
affine map to 1..255, ``np.rint`` (half-to-even).  Bit-exact by construction (tests/ compare it with the CPU restatement).
"""
import numpy as np

F0_MIN, F0_MAX = 50.0, 1100.0


def coarse_pitch(f0):
    f0 = np.asarray(f0, dtype=np.float64)
    mel_min = 1127.0 * np.log(1.0 + F0_MIN / 700.0)
    mel_max = 1127.0 * np.log(1.0 + F0_MAX / 700.0)
    mel = 1127.0 * np.log(1.0 + f0 / 700.0)
    voiced = mel > 0
    mel[voiced] = (mel[voiced] - mel_min) * 254.0 / (mel_max - mel_min) + 1.0
    mel[mel <= 1] = 1
    mel[mel > 255] = 255
    return np.rint(mel).astype(np.int32)
