"""Pin oracle/f0_oracle.py (RMVPE.decode + the tail of VC.get_f0) against tests/golden/f0_tail.npz, which oracle/make_f0_golden.py wrote
from the UNMODIFIED reference code objects.  Everything here is numpy on both sides: the bar is bit-exact (float64 values compared by
bit pattern, NaN-free by construction), except ``2 ** x`` and ``log`` whose last bit may depend on the host's libm / numpy SIMD build:
float64 f0 within 4 ulp, the integer pitch and the float32 ``pitchf`` exact."""
import os

import numpy as np
import pytest

from oracle import f0_oracle
from oracle.make_f0_golden import SALIENCE_CASES, TAIL_CASES


def ulp_close(a, b, ulps=4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b))))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "f0_tail.npz"))


@pytest.mark.parametrize("name", sorted(SALIENCE_CASES))
def test_decode_matches_reference(name, gold):
    T, seed, sharp = SALIENCE_CASES[name]
    sal = f0_oracle.synthetic_salience(T, seed, sharp=sharp)
    with np.errstate(all="ignore"):
        cents = f0_oracle.to_local_average_cents(sal, thred=0.03)
        f0 = f0_oracle.decode(sal, thred=0.03)
        f0_legacy = f0_oracle.decode_with_pitch(sal, thred=0.03)
    assert np.array_equal(cents.view(np.int64), gold[name + ".cents"].view(np.int64))     # sums and quotient: bit-exact
    assert ulp_close(f0, gold[name + ".f0"]) and np.array_equal(f0 == 0, gold[name + ".f0"] == 0)
    assert ulp_close(f0_legacy, gold[name + ".f0_legacy"]) and np.array_equal(f0_legacy == 0, gold[name + ".f0_legacy"] == 0)
    assert (f0 > 0).sum() > T // 4 and (f0 == 0).sum() >= 3                                # both branches exercised
    pitch, pitchf = f0_oracle.get_f0_tail(f0, pitch=2)
    assert np.array_equal(pitch, gold[name + ".pitch"]) and pitch.dtype == np.int32
    assert np.array_equal(pitchf.astype(np.float32), gold[name + ".pitchf64"].astype(np.float32))


@pytest.mark.parametrize("name", sorted(TAIL_CASES))
def test_get_f0_tail_matches_reference(name, gold):
    T, seed, shift, tune, strength = TAIL_CASES[name]
    f0 = f0_oracle.synthetic_f0(T, seed)
    pitch, pitchf = f0_oracle.get_f0_tail(f0, pitch=shift, f0_autotune=tune, f0_autotune_strength=strength)
    assert np.array_equal(pitch, gold[name + ".pitch"])
    assert ulp_close(pitchf, gold[name + ".pitchf64"], ulps=1)
    assert pitch.min() == 1 and pitch.max() == (255 if shift >= 0 else pitch.max())
    p64, pf32 = f0_oracle.infer_inputs(pitch, pitchf, p_len=T - 3)
    assert p64.dtype == np.int64 and pf32.dtype == np.float32 and len(p64) == T - 3


def test_host_coarse_pitch_equals_oracle():
    from rvc_maker_b200.pitch import coarse_pitch

    f0 = f0_oracle.synthetic_f0(4000, 77)
    assert np.array_equal(coarse_pitch(f0), f0_oracle.get_f0_tail(f0)[0])


def test_pairwise_sum_order_is_numpy():
    """The device kernel adds the 9 window terms as ((t0+t1)+(t2+t3))+((t4+t5)+(t6+t7)) then +t8 — numpy's pairwise reduction for n = 9."""
    rng = np.random.default_rng(5)
    for dt in (np.float32, np.float64):
        a = (rng.random((5000, 9)) * rng.choice([1e-3, 1.0, 1e3], size=(5000, 9))).astype(dt)
        r = ((a[:, 0] + a[:, 1]) + (a[:, 2] + a[:, 3])) + ((a[:, 4] + a[:, 5]) + (a[:, 6] + a[:, 7])) + a[:, 8]
        assert np.array_equal(r, np.sum(a, 1))
