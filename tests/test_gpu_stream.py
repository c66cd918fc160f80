"""BASELINE config 5 at reduced length: a synthetic stream cut by the reference's chunk plan, synthesised chunk by chunk on the
GPU and stitched, against the oracle run on the identical plan (per-chunk results depend on the plan, so parity is
defined per chunk).  Exercises T = 4000 frames (attention key length 4096) and two plan shapes in one engine."""
import numpy as np
import pytest
import torch

from oracle import cases, synth_oracle
from rvc_maker_b200 import chunks as ch

pytestmark = pytest.mark.gpu


def test_stream_chunks_match_oracle_per_chunk():
    from rvc_maker_b200 import Synthesizer

    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=1234, heated=True)
    seconds = 45
    n = seconds * ch.SAMPLE_RATE
    spans = ch.chunk_spans(n, ch.uniform_split_points(n))
    assert [c.frames for c in spans] == [4000, 900]
    total_frames = spans[-1].frame_end
    inp = cases.synth_inputs(cfg, 1, total_frames, seed=2)
    net = Synthesizer(*cases.ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision="tc")
    del net.enc_q
    net.load_state_dict(sd, strict=True)
    net = net.eval().to("cuda:0")
    upp = cases.upp_of(cfg)
    rng = np.random.Generator(np.random.PCG64(5))
    folded = synth_oracle.fold_state_dict(sd)
    sr = cfg["sr"]

    def piece(c, use_gpu):
        sl = slice(c.frame_start, c.frame_end)
        T = c.frames
        eps_z = torch.from_numpy(np.random.Generator(np.random.PCG64(100 + c.index)).standard_normal((1, cfg["inter_channels"], T), dtype=np.float32))
        eps_s = torch.from_numpy(np.random.Generator(np.random.PCG64(200 + c.index)).standard_normal((1, T * upp, 1), dtype=np.float32))
        ph, pi, f0 = inp["phone"][:, sl], inp["pitch"][:, sl], inp["nsff0"][:, sl]
        ln = torch.tensor([T])
        if use_gpu:
            net._noise_override = (eps_z, eps_s)
            o = net.infer(ph.cuda(), ln.cuda(), pi.cuda(), f0.cuda(), inp["sid"].cuda())[0]
            return o[0, 0].float().cpu().numpy()
        o = synth_oracle.infer(sd, cfg, ph, ln, pi, f0, inp["sid"], eps_z=eps_z, eps_src=eps_s, folded=folded)[0]
        return o[0, 0].numpy()

    owner = ch.shard_chunks(spans, 1)
    got = ch.gather_stream(ch.synthesize_chunks(lambda c: piece(c, True), spans, owner, 0, sr), len(spans))
    want = ch.gather_stream(ch.synthesize_chunks(lambda c: piece(c, False), spans, owner, 0, sr), len(spans))
    assert got.shape == want.shape == (seconds * sr,)
    m = cases.metrics(torch.from_numpy(got), torch.from_numpy(want))
    print(f"\n[stream 45 s, chunks 4000+900 frames, tc] max_abs={m['max_abs']:.3e} snr={m['snr_db']:.1f} dB")
    assert m["max_abs"] <= 1e-3 and m["snr_db"] >= 40.0, m
