"""Shapes for `compute-sanitizer --tool memcheck|initcheck python tests/memcheck_cases.py`: the cases where a 512-row tile
or the transposed mode's extra tile used to run past the packed rows (B = 40 x T = 64, T = 1000, T = 512), eager launches only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import cases  # noqa: E402
from rvc_maker_b200 import Synthesizer  # noqa: E402


def main():
    cfg = cases.CONFIGS["v2_48k"]
    sd = cases.synth_state_dict(cfg, seed=778, heated=True)
    shapes = [(40, 64), (1, 1000), (2, 512)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for prec in ("tc", "fp32"):
        net = Synthesizer(*cases.ctor_args(cfg), use_f0=True, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision=prec, use_cuda_graph=False)
        del net.enc_q
        net.load_state_dict(sd, strict=True)
        net = net.eval().to("cuda:0").float()
        for B, T in shapes:
            inp = cases.synth_inputs(cfg, B, T, seed=B + T, ragged=True)
            net._noise_override = (inp["eps_z"], inp["eps_src"])
            o = net.infer(inp["phone"].cuda(), inp["phone_lengths"].cuda(), inp["pitch"].cuda(), inp["nsff0"].cuda(), inp["sid"].cuda())[0]
            torch.cuda.synchronize()
            print(f"{prec} B={B} T={T}: o{tuple(o.shape)} finite={bool(torch.isfinite(o).all())} absmax={float(o.abs().max()):.4f}", flush=True)


if __name__ == "__main__":
    main()
