"""Does any kernel read workspace memory it never wrote?  Run a wide-batch case in a fresh process, then again after filling the
caching allocator's free blocks with NaN, and report where the outputs differ (batch item, position, encoder vs decoder).

usage (on a B200): python tools/poison_probe.py [B T]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases, synth_oracle  # noqa: E402  (diagnostic tool: the oracle is the checker here)
from rvc_maker_b200 import Synthesizer  # noqa: E402

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 64)
cfg = cases.CONFIGS["v2_48k"]
sd = cases.synth_state_dict(cfg, seed=778, heated=True)
inp = cases.synth_inputs(cfg, B, T, seed=B + T, ragged=B > 1)
ro, _, (rz, rzp, rm, rl) = synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"],
                                           eps_z=inp["eps_z"], eps_src=inp["eps_src"])


def run(tag):
    net = Synthesizer(*cases.ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision="tc")
    del net.enc_q
    net.load_state_dict(sd, strict=True)
    net = net.eval().to("cuda:0")
    net._noise_override = (inp["eps_z"], inp["eps_src"])
    outs = []
    for rep in range(3):   # eager, graph capture, graph replay
        o, _, (z, z_p, m_p, logs_p) = net.infer(*(inp[k].cuda() for k in ("phone", "phone_lengths", "pitch", "nsff0", "sid")))
        torch.cuda.synchronize()
        if rep == 0:
            for k, v, r in (("o", o, ro), ("z", z, rz), ("z_p", z_p, rzp), ("m_p", m_p, rm), ("logs_p", logs_p, rl)):
                mm = cases.metrics(v.float().cpu(), r)
                print(f"   {k}: max_abs={mm['max_abs']:.4e} max|ref|={float(r.abs().max()):.3f} rel={mm['max_abs'] / max(1.0, float(r.abs().max())):.3e} snr={mm['snr_db']:.1f} dB")
        outs.append((o.float().cpu(), z.float().cpu(), m_p.float().cpu()))
    for rep, (o, z, m) in enumerate(outs):
        eo = (o - ro).abs()
        per_item = eo.amax(dim=(1, 2))
        b = int(per_item.argmax())
        pos = int(eo[b, 0].argmax())
        print(f"[{tag} call {rep}] o: max_abs={float(eo.max()):.3e} nan={int(torch.isnan(o).sum())} worst item {b} (len {int(inp['phone_lengths'][b])}) at sample {pos} "
              f"(frame {pos // 480}); items > 1e-3: {[int(i) for i in (per_item > 1e-3).nonzero().flatten()][:12]}; "
              f"z: {float((z - rz).abs().max()):.3e}/{float(rz.abs().max()):.1f} m_p: {float((m - rm).abs().max()):.3e}; "
              f"same as call 0: {torch.equal(o, outs[0][0])}", flush=True)
    return outs[0][0]


a = run("fresh process")
junk = torch.full(((24 << 30) // 4,), float("nan"), device="cuda:0")   # poison what the allocator will hand out next
del junk
b = run("after NaN poisoning of freed blocks")
print("fresh == poisoned:", torch.equal(a, b))
