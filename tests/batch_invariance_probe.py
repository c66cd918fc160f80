"""Is Synthesizer.infer batch-invariant?  Two B = 1 items run one by one against the same two items stacked as one B = 2 call (pinned noise):
prints the max |difference| of every output.  Not a test: `python tests/batch_invariance_probe.py [T ...]` on a GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import cases  # noqa: E402
from rvc_maker_b200 import Synthesizer  # noqa: E402


def main():
    Ts = [int(a) for a in sys.argv[1:]] or [200, 1000]
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=5, heated=True)
    for prec in ("tc", "fp32"):
        net = Synthesizer(*cases.ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision=prec)
        del net.enc_q
        net.load_state_dict(sd, strict=False)
        net = net.eval().cuda()
        for T in Ts:
            a, b = cases.synth_inputs(cfg, 1, T, seed=1), cases.synth_inputs(cfg, 1, T, seed=2)
            outs = []
            for inp in (a, b):
                net._noise_override = (inp["eps_z"], inp["eps_src"])
                o, m, (z, z_p, m_p, logs_p) = net.infer(inp["phone"].cuda(), inp["phone_lengths"].cuda(), inp["pitch"].cuda(), inp["nsff0"].cuda(), inp["sid"].cuda())
                outs.append([t.clone() for t in (o, z, z_p, m_p, logs_p)])
            cat = {k: torch.cat([a[k], b[k]]) for k in ("phone", "phone_lengths", "pitch", "nsff0", "sid", "eps_z", "eps_src")}
            net._noise_override = (cat["eps_z"], cat["eps_src"])
            o, m, (z, z_p, m_p, logs_p) = net.infer(cat["phone"].cuda(), cat["phone_lengths"].cuda(), cat["pitch"].cuda(), cat["nsff0"].cuda(), cat["sid"].cuda())
            net._noise_override = None
            for i in range(2):
                d = {n: float((t[i:i + 1] - r).abs().max()) for n, t, r in zip(("o", "z", "z_p", "m_p", "logs_p"), (o, z, z_p, m_p, logs_p), outs[i])}
                print(prec, "T", T, "item", i, d)


if __name__ == "__main__":
    main()
