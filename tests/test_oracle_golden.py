"""Pin the oracle restatement (oracle/synth_oracle.py) against outputs of the UNMODIFIED reference module.

tests/golden/*.npz were written by oracle/make_golden.py in the build container (reference imported from
/root/reference).  Tolerances: fp32 oracle vs fp32 reference on CPU differ only by operation-order rounding
(fp32 floor 2.4e-7 vs fp64), so max-abs <= 5e-6 x max(1, max|ref|).
"""
import os

import numpy as np
import pytest
import torch

from oracle import cases, synth_oracle
from oracle.make_golden import CASES

TOL = 5e-6


def run_oracle_case(name, dtype=torch.float32, taps=None):
    cname, use_f0, heated, disk, B, T, ragged, rate, wseed, iseed, _ = CASES[name]
    cfg = cases.CONFIGS[cname]
    sd = cases.synth_state_dict(cfg, use_f0=use_f0, seed=wseed, heated=heated)
    if disk:
        sd = cases.to_disk_form(sd)
    inp = cases.synth_inputs(cfg, B, T, seed=iseed, ragged=ragged, use_f0=use_f0)
    o, x_mask, (z, z_p, m_p, logs_p) = synth_oracle.infer(
        sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"], rate,
        eps_z=inp["eps_z"], eps_src=inp["eps_src"], use_f0=use_f0, dtype=dtype, taps=taps)
    return dict(o=o, x_mask=x_mask, z=z, z_p=z_p, m_p=m_p, logs_p=logs_p)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name, golden_dir):
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    taps = {}
    out = run_oracle_case(name, taps=taps)
    for k, v in out.items():
        g = torch.from_numpy(gold[k])
        assert tuple(v.shape) == tuple(g.shape), (k, v.shape, g.shape)
        m = cases.metrics(v.float(), g)
        assert m["max_abs"] <= TOL * max(1.0, float(g.abs().max())), (name, k, m)
    if "tap_har_source" in gold.files:
        m = cases.metrics(taps["har_source"], torch.from_numpy(gold["tap_har_source"]))
        assert m["max_abs"] <= TOL, (name, "har_source", m)


def test_oracle_fp64_floor(golden_dir):
    """fp64 run of the restatement vs the fp32 reference golden: the reference's own rounding floor."""
    gold = np.load(os.path.join(golden_dir, "v2_48k_f0.npz"))
    out = run_oracle_case("v2_48k_f0", dtype=torch.float64)
    m = cases.metrics(out["o"].float(), torch.from_numpy(gold["o"]))
    assert m["max_abs"] <= TOL and m["snr_db"] >= 90, m


def test_coarse_pitch_known_answers():
    """convert.py:318-323 — integer-exact known answers computed by hand from the formula."""
    f0 = np.array([0.0, 49.0, 50.0, 1100.0, 2000.0, 440.0, 100.0])
    got = cases.coarse_pitch(f0)
    mel = lambda f: 1127.0 * np.log(1.0 + f / 700.0)
    lo, hi = mel(50.0), mel(1100.0)
    want = [1, 1, 1, 255, 255, int(np.rint((mel(440.0) - lo) * 254 / (hi - lo) + 1)), int(np.rint((mel(100.0) - lo) * 254 / (hi - lo) + 1))]
    assert got.tolist() == want
    assert got.dtype == np.int32 and got.min() >= 1 and got.max() <= 255


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference_midsize():
    """Live check at a mid-size the fixtures do not store (T=200, v2-40k, ragged batch)."""
    from oracle.make_golden import build_reference, run_reference

    cfg = cases.CONFIGS["v2_40k"]
    sd = cases.synth_state_dict(cfg, seed=77, heated=True)
    inp = cases.synth_inputs(cfg, 2, 200, seed=9, ragged=True)
    ref = run_reference(build_reference(cfg, sd, True), inp, True)
    o, _, (z, z_p, m_p, logs_p) = synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"],
                                                      inp["sid"], eps_z=inp["eps_z"], eps_src=inp["eps_src"])
    for a, b, k in ((o, ref["o"], "o"), (z, ref["z"], "z"), (m_p, ref["m_p"], "m_p")):
        m = cases.metrics(a, b)
        assert m["max_abs"] <= TOL * max(1.0, float(b.abs().max())), (k, m)
