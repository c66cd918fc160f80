"""The f0 tail on the device (csrc/f0_post.cu through rvc_maker_b200.pitch) against the oracle and the reference's golden outputs.

Bars: ``cents`` path bit-exact by construction, so the voiced / unvoiced decision and the coarse ``pitch`` (int64) are compared exactly; ``pitchf``
(float32) exactly; the float64 f0 within 4 ulp (exp2 / log: CUDA's routines vs the host libm, both <= 1 ulp)."""
import os

import numpy as np
import pytest
import torch

from oracle import f0_oracle
from oracle.make_f0_golden import SALIENCE_CASES, TAIL_CASES

pytestmark = pytest.mark.gpu


def ulp_close(a, b, ulps=4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b))))


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "f0_tail.npz"))


@pytest.mark.parametrize("name", sorted(SALIENCE_CASES))
def test_decode_matches_reference_golden(name, gold):
    from rvc_maker_b200 import pitch as P

    T, seed, sharp = SALIENCE_CASES[name]
    sal = torch.from_numpy(f0_oracle.synthetic_salience(T, seed, sharp=sharp)).cuda()
    f0 = P.decode(sal, thred=0.03).cpu().numpy()
    g = gold[name + ".f0"]
    assert f0.dtype == np.float64 and np.array_equal(f0 == 0, g == 0) and ulp_close(f0, g)
    f0l = P.decode_with_pitch(sal.unsqueeze(0), thred=0.03).cpu().numpy()          # the network's [1, T, 360] shape
    gl = gold[name + ".f0_legacy"]
    assert np.array_equal(f0l == 0, gl == 0) and ulp_close(f0l, gl)
    # fused: salience -> pitch / pitchf (+2 semitones, as the fixture)
    p, pf, f0s = P.rmvpe_pitch(sal, pitch=2, return_f0=True)
    assert p.dtype == torch.int64 and pf.dtype == torch.float32
    assert np.array_equal(p.cpu().numpy(), gold[name + ".pitch"].astype(np.int64))
    assert np.array_equal(pf.cpu().numpy(), gold[name + ".pitchf64"].astype(np.float32))
    assert ulp_close(f0s.cpu().numpy(), gold[name + ".pitchf64"])
    # two launches == one launch, bit for bit
    p2, pf2 = P.get_f0_tail(P.decode(sal), pitch=2)
    assert torch.equal(p2, p) and torch.equal(pf2, pf)


@pytest.mark.parametrize("name", sorted(TAIL_CASES))
def test_get_f0_tail_matches_reference_golden(name, gold):
    from rvc_maker_b200 import pitch as P

    T, seed, shift, tune, strength = TAIL_CASES[name]
    f0 = torch.from_numpy(f0_oracle.synthetic_f0(T, seed)).cuda()
    p, pf, f0s = P.get_f0_tail(f0, pitch=shift, f0_autotune=tune, f0_autotune_strength=strength, return_f0=True)
    assert np.array_equal(p.cpu().numpy(), gold[name + ".pitch"].astype(np.int64))
    assert np.array_equal(pf.cpu().numpy(), gold[name + ".pitchf64"].astype(np.float32))
    assert ulp_close(f0s.cpu().numpy(), gold[name + ".pitchf64"], ulps=1)


@pytest.mark.parametrize("T,legacy,tune", [(1, False, False), (7, True, True), (60001, False, True), (60001, True, False), (200003, False, False)])
def test_rmvpe_pitch_against_oracle_at_size(T, legacy, tune):
    """Sizes beyond the fixtures (10 minutes of audio at 100 frames/s; 200 003 frames for the eight-frames-per-warp variant) against the numpy oracle
    on the same seeded salience."""
    from rvc_maker_b200 import pitch as P

    sal = f0_oracle.synthetic_salience(T, 1000 + T, sharp=bool(T & 1))
    with np.errstate(all="ignore"):
        f0 = f0_oracle.decode_with_pitch(sal) if legacy else f0_oracle.decode(sal)
    want_p, want_f = f0_oracle.get_f0_tail(f0, pitch=-5, f0_autotune=tune, f0_autotune_strength=0.75)
    p, pf = P.rmvpe_pitch(torch.from_numpy(sal).cuda(), pitch=-5, f0_autotune=tune, f0_autotune_strength=0.75, legacy=legacy)
    assert np.array_equal(p.cpu().numpy(), want_p.astype(np.int64))
    got_f = pf.cpu().numpy()
    assert np.array_equal(got_f == 0, want_f == 0)
    assert np.array_equal(got_f, want_f.astype(np.float32))


def test_strided_rows_and_empty():
    from rvc_maker_b200 import pitch as P

    sal = torch.from_numpy(f0_oracle.synthetic_salience(300, 9)).cuda()
    wide = torch.zeros(300, 512, device="cuda")
    wide[:, :360] = sal
    assert torch.equal(P.decode(wide[:, :360]), P.decode(sal))                        # row stride 512
    assert P.decode(sal[:0]).shape == (0,)
    p, pf = P.rmvpe_pitch(sal[:0])
    assert p.shape == (0,) and pf.shape == (0,)


def test_pitch_feeds_infer_like_the_reference():
    """The device pitch / pitchf drive Synthesizer.infer to the same output as the host-made ones (convert.py:436-439 then :381)."""
    from oracle import cases
    from rvc_maker_b200 import Synthesizer
    from rvc_maker_b200 import pitch as P

    cfg = cases.CONFIGS["tiny"]
    T = 64
    sal = f0_oracle.synthetic_salience(T, 4)
    with np.errstate(all="ignore"):
        hp, hf = f0_oracle.infer_inputs(*f0_oracle.get_f0_tail(f0_oracle.decode(sal), pitch=3))
    dp, df = P.rmvpe_pitch(torch.from_numpy(sal).cuda(), pitch=3)
    assert np.array_equal(dp.cpu().numpy(), hp) and np.array_equal(df.cpu().numpy(), hf)
    net = Synthesizer(*cases.ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"])
    del net.enc_q                                                                     # convert.py:565
    missing, unexpected = net.load_state_dict(cases.synth_state_dict(cfg, use_f0=True, seed=3, heated=True), strict=False)
    assert not missing and not unexpected
    net = net.eval().cuda()
    inp = cases.synth_inputs(cfg, 1, T, seed=3, ragged=False, use_f0=True)
    net._noise_override = (inp["eps_z"], inp["eps_src"])
    try:
        a = net.infer(inp["phone"].cuda(), inp["phone_lengths"].cuda(), dp[None], df[None], inp["sid"].cuda())[0]
        b = net.infer(inp["phone"].cuda(), inp["phone_lengths"].cuda(), torch.from_numpy(hp)[None].cuda(), torch.from_numpy(hf)[None].cuda(),
                      inp["sid"].cuda())[0]
    finally:
        net._noise_override = None
    assert torch.equal(a, b)
