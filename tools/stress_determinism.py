"""Stress loop for timing-dependent hazards: run one case many times in ONE process — fresh engines, eager / captured / replayed calls,
other shapes in between — and report every call whose waveform is not bit-identical to the first (no oracle needed: every kernel of
the path is deterministic by construction, so any difference is a race or a read of stale memory).

usage (on a B200): python tools/stress_determinism.py [--B 40 --T 64 --rounds 30 --precision tc]
Written for the open issue in DESIGN.md section 8 item 0 (the B = 40 x T = 64 parity case failed once inside a long test process).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rvc_maker_b200 import Synthesizer, configs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=40)
ap.add_argument("--T", type=int, default=64)
ap.add_argument("--rounds", type=int, default=30)
ap.add_argument("--precision", default="tc")
ap.add_argument("--config", default="v2_48k")
args = ap.parse_args()

cfg = configs.CONFIGS[args.config]
dev = torch.device("cuda:0")
torch.manual_seed(1234)
proto = Synthesizer(*configs.ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision=args.precision)
sd = {k: v.clone() for k, v in proto.state_dict().items()}
inp = {k: v.to(dev) for k, v in bench.synth_inputs(cfg, args.B, args.T, seed=3).items()}
inp["phone_lengths"] = torch.randint(args.T // 2, args.T + 1, (args.B,), generator=torch.Generator().manual_seed(5)).to(dev)
other = {k: v.to(dev) for k, v in bench.synth_inputs(cfg, 2, 333, seed=4).items()}
g = torch.Generator().manual_seed(7)
upp = 1
for u in cfg["upsample_rates"]:
    upp *= u
noise = (torch.randn(args.B, cfg["inter_channels"], args.T, generator=g), torch.randn(args.B, args.T * upp, 1, generator=g))

first, bad = None, 0
for r in range(args.rounds):
    net = Synthesizer(*configs.ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision=args.precision)
    net.load_state_dict(sd)
    del net.enc_q
    net = net.eval().to(dev)
    for call in range(4):
        net._noise_override = noise
        o = net.infer(inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"])[0]
        if call == 1:   # another shape in between (evicts nothing, but moves the allocator and the L2 contents)
            net._noise_override = None
            net.infer(other["phone"], other["phone_lengths"], other["pitch"], other["nsff0"], other["sid"])
        torch.cuda.synchronize()
        if first is None:
            first = o.clone()
            continue
        if not torch.equal(o, first):
            bad += 1
            d = (o.float() - first.float()).abs()
            b = int(d.amax(dim=(1, 2)).argmax())
            print(f"round {r} call {call}: DIFFERS  max|d|={float(d.max()):.3e} nan={int(torch.isnan(o).sum())} item {b} at sample {int(d[b, 0].argmax())}", flush=True)
print(f"{args.rounds} rounds x 4 calls of B={args.B} T={args.T} ({args.precision}): {bad} calls differ from the first")
sys.exit(1 if bad else 0)
