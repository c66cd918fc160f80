"""Restatement of the f0 tail between the RMVPE network and ``Synthesizer.infer`` (TEST INFRASTRUCTURE).

Two pieces of the reference, restated with the numpy calls the reference makes (the post-network part):

* ``RMVPE.decode`` / ``to_local_average_cents`` (main/library/predictors/RMVPE.py:217-221, :237-252): salience ``[T, 360]`` fp32 ->
  f0 ``[T]`` float64 (Hz, 0 = unvoiced);
* the tail of ``VC.get_f0`` (main/inference/convert.py:308-323) with ``Autotune.autotune_f0`` (:173-179): autotune, the pitch shift
  ``f0 *= 2^(pitch/12)``, the coarse mel quantiser -> ``(pitch int32 1..255, f0 float64)``, and the two casts the pipeline applies
  before ``infer`` (:437-439: ``.long()`` / ``.float()``).

Pinned by tests/golden/f0_tail.npz, written by oracle/make_f0_golden.py from the UNMODIFIED reference code objects (RMVPE.py imported
with a stub ``librosa.filters``; ``VC.get_f0`` / ``Autotune.autotune_f0`` taken as code objects out of the compiled convert.py, which
cannot be imported here).  Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline leg may import this file.
"""
import numpy as np

N_CLASS = 360
F0_MIN, F0_MAX = 50, 1100                                    # convert.py:196-197
# convert.py:200 — the 54 note frequencies `Autotune` snaps to
REF_FREQS = [49.00, 51.91, 55.00, 58.27, 61.74, 65.41, 69.30, 73.42, 77.78, 82.41, 87.31, 92.50, 98.00, 103.83, 110.00, 116.54, 123.47,
             130.81, 138.59, 146.83, 155.56, 164.81, 174.61, 185.00, 196.00, 207.65, 220.00, 233.08, 246.94, 261.63, 277.18, 293.66,
             311.13, 329.63, 349.23, 369.99, 392.00, 415.30, 440.00, 466.16, 493.88, 523.25, 554.37, 587.33, 622.25, 659.25, 698.46,
             739.99, 783.99, 830.61, 880.00, 932.33, 987.77, 1046.50]


def cents_mapping():
    """RMVPE.py:207-208: 360 bins of 20 cents from 1997.379..., zero-padded by 4 on both sides."""
    return np.pad(20 * np.arange(N_CLASS) + 1997.3794084376191, (4, 4))


def to_local_average_cents(salience, thred=0.05):
    """RMVPE.py:237-252, vectorised: the 9 bins around each frame's arg-max, salience-weighted mean of their cents.

    The products are float32 x float64 -> float64, summed over 9 elements by numpy's pairwise reduction; the denominator is the
    float32 sum of the same 9 saliences (same reduction order, float32 accumulators); the quotient is float64."""
    salience = np.asarray(salience)
    center = np.argmax(salience, axis=1)                      # :238 (first maximum)
    padded = np.pad(salience, ((0, 0), (4, 4)))               # :239
    idx = center[:, None] + np.arange(9)[None, :]             # :240-242: padded columns center-4+4 .. center+4+4
    todo_salience = np.take_along_axis(padded, idx, axis=1)   # :245-246
    todo_cents = cents_mapping()[idx]                         # :247
    devided = np.sum(todo_salience * todo_cents, 1) / np.sum(todo_salience, 1)   # :250
    devided[np.max(padded, axis=1) <= thred] = 0              # :251
    return devided


def decode(hidden, thred=0.03):
    """RMVPE.py:217-221."""
    f0 = 10 * (2 ** (to_local_average_cents(hidden, thred=thred) / 1200))
    f0[f0 == 10] = 0
    return f0


def decode_with_pitch(hidden, thred=0.03, f0_min=F0_MIN, f0_max=F0_MAX):
    """RMVPE.infer_from_audio_with_pitch's tail (RMVPE.py:229-233, the "rmvpe-legacy" method)."""
    f0 = decode(hidden, thred)
    f0[(f0 < f0_min) | (f0 > f0_max)] = 0
    return f0


def autotune_f0(f0, strength, note_dict=REF_FREQS):
    """convert.py:173-179: every frame moves towards the nearest note (first one on ties), unvoiced frames included."""
    out = np.zeros_like(f0)
    for i, freq in enumerate(f0):
        out[i] = freq + (min(note_dict, key=lambda x: abs(x - freq)) - freq) * strength
    return out


def get_f0_tail(f0, pitch=0, f0_autotune=False, f0_autotune_strength=1.0):
    """convert.py:308-323 (without the ``inp_f0`` file override, :314-316): returns (coarse pitch int32, f0 float64)."""
    f0 = np.array(f0, dtype=np.float64)
    f0_mel_min = 1127 * np.log(1 + F0_MIN / 700)              # :198
    f0_mel_max = 1127 * np.log(1 + F0_MAX / 700)              # :199
    if f0_autotune:
        f0 = autotune_f0(f0, f0_autotune_strength)            # :308
    f0 *= pow(2, pitch / 12)                                  # :311
    f0_mel = 1127 * np.log(1 + f0 / 700)                      # :318
    f0_mel[f0_mel > 0] = (f0_mel[f0_mel > 0] - f0_mel_min) * 254 / (f0_mel_max - f0_mel_min) + 1   # :319
    f0_mel[f0_mel <= 1] = 1                                   # :320
    f0_mel[f0_mel > 255] = 255                                # :321
    return np.rint(f0_mel).astype(np.int32), f0.copy()        # :323


def infer_inputs(pitch_i32, f0_f64, p_len=None):
    """convert.py:437-439: what ``infer`` receives — ``pitch`` int64 and ``pitchf`` float32, cut to p_len frames."""
    if p_len is not None:
        pitch_i32, f0_f64 = pitch_i32[:p_len], f0_f64[:p_len]
    return pitch_i32.astype(np.int64), f0_f64.astype(np.float32)


def synthetic_salience(T, seed, voiced=0.7, sharp=True):
    """A seeded RMVPE-like salience map [T, 360] fp32: a smooth peak wandering over the bins, runs of unvoiced frames whose maximum
    stays below the threshold, peaks at both edges (windows reaching the zero padding), exact ties and a few all-zero frames."""
    rng = np.random.default_rng(seed)
    bins = np.arange(N_CLASS, dtype=np.float64)[None, :]
    walk = np.cumsum(rng.normal(0, 2.0, T)) + rng.uniform(40, 300)
    centre = np.abs(np.mod(walk, 2 * 359.0) - 359.0)         # reflect into [0, 359]
    width = rng.uniform(0.8, 3.0, T)[:, None] if sharp else rng.uniform(3.0, 12.0, T)[:, None]
    height = rng.uniform(0.2, 0.98, T)[:, None]
    sal = height * np.exp(-0.5 * ((bins - centre[:, None]) / width) ** 2) + rng.uniform(0, 0.02, (T, N_CLASS))
    t = 0
    while t < T:                                             # unvoiced runs
        run = int(rng.integers(5, 40))
        if rng.random() > voiced:
            sal[t:t + run] = rng.uniform(0, 0.029, (min(run, T - t), N_CLASS))
        t += run
    sal = sal.astype(np.float32)
    if T >= 16:
        sal[1, :] = 0; sal[1, 0] = 0.9; sal[1, 1] = 0.5          # peak at bin 0: window reaches the left padding
        sal[2, :] = 0; sal[2, 359] = 0.8; sal[2, 357] = 0.3      # peak at bin 359
        sal[3, :] = 0                                            # all zero: arg-max 0, 0/0 -> NaN, then thresholded to 0
        sal[4, :] = 0; sal[4, 100] = 0.5; sal[4, 200] = 0.5      # exact tie: first maximum wins
        sal[5, :] = 0; sal[5, 50] = np.float32(0.03)             # maximum == thred: unvoiced (<=)
        sal[6, :] = 0; sal[6, 50] = np.nextafter(np.float32(0.03), np.float32(1))
        sal[7, :] = 0; sal[7, 3] = 0.7; sal[7, 7] = 0.69         # window partly in the padding
    return sal


def synthetic_f0(T, seed):
    """Seeded float64 f0 track: log-random-walk contour 40..1300 Hz (beyond both quantiser limits), unvoiced runs, special values."""
    rng = np.random.default_rng(seed)
    f0 = np.exp(np.clip(np.log(200.0) + np.cumsum(rng.normal(0, 0.03, T)), np.log(40.0), np.log(1300.0)))
    t = 0
    while t < T:
        run = int(rng.integers(5, 40))
        if rng.random() < 0.3:
            f0[t:t + run] = 0
        t += run
    if T >= 8:
        f0[0], f0[1], f0[2], f0[3], f0[4] = 50.0, 1100.0, 49.0, 1101.0, 1e-9
    return f0
