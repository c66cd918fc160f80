"""Where the tensor-core path's waveform error comes from: the same case with the encoder and / or the flow (+ conv_pre) moved back to the
fp32 CUDA-core kernels (RVC_B200_ENC_TC / RVC_B200_FLOW_TC = 0), and with / without the split-fp16 residual stream.  Not a test:
`python tests/error_budget.py [case ...]` on a GPU box prints one line per (case, mode)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import cases, synth_oracle  # noqa: E402

CASES = {"v2_48k_b2": ("v2_48k", True, 2, 1000, 778, 5), "v1_32k_b16": ("v1_32k", True, 16, 1000, 4321, 16), "v2_40k_nof0_b4": ("v2_40k", False, 4, 1000, 4321, 32),
         "v2_48k_b40_t64": ("v2_48k", True, 40, 64, 778, 104)}
MODES = {"tc": {}, "enc_fp32": {"RVC_B200_ENC_TC": "0"}, "flow_fp32": {"RVC_B200_FLOW_TC": "0"}, "encflow_fp32": {"RVC_B200_ENC_TC": "0", "RVC_B200_FLOW_TC": "0"},
         "nosplit": {"RVC_B200_PAIR_SPLIT": "0"}, "enc_nosplit": {"RVC_B200_ENC_SPLIT": "0"}, "fp32": {"precision": "fp32"}}


def main():
    from rvc_maker_b200 import Synthesizer

    names = sys.argv[1:] or list(CASES)
    for name in names:
        cname, use_f0, B, T, wseed, iseed = CASES[name]
        cfg = cases.CONFIGS[cname]
        sd = cases.synth_state_dict(cfg, use_f0=use_f0, seed=wseed, heated=True)
        inp = cases.synth_inputs(cfg, B, T, seed=iseed, ragged=True, use_f0=use_f0)
        ro, _, (rz, rzp, rm, rl) = synth_oracle.infer(sd, cfg, inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"],
                                                      eps_z=inp["eps_z"], eps_src=inp["eps_src"], use_f0=use_f0)
        for mode, env in MODES.items():
            env = dict(env)
            prec = env.pop("precision", "tc")
            for k in ("RVC_B200_ENC_TC", "RVC_B200_FLOW_TC", "RVC_B200_PAIR_SPLIT", "RVC_B200_ENC_SPLIT"):
                os.environ.pop(k, None)
            os.environ.update(env)
            net = Synthesizer(*cases.ctor_args(cfg), use_f0=use_f0, text_enc_hidden_dim=cfg["text_enc_hidden_dim"], precision=prec)
            del net.enc_q
            net.load_state_dict(sd, strict=True)
            net = net.eval().to("cuda:0").float()
            net._noise_override = (inp["eps_z"], inp["eps_src"])
            mv = lambda t: None if t is None else t.cuda()
            o, _, (z, zp, m, l) = net.infer(mv(inp["phone"]), mv(inp["phone_lengths"]), mv(inp["pitch"]), mv(inp["nsff0"]), mv(inp["sid"]))
            torch.cuda.synchronize()
            mo, mz, mm = cases.metrics(o.cpu(), ro), cases.metrics(z.cpu(), rz), cases.metrics(m.cpu(), rm)
            print(f"{name:16s} {mode:13s} o max_abs={mo['max_abs']:.3e} snr={mo['snr_db']:.1f} ac={mo['snr_ac_db']:.1f} | z max_abs={mz['max_abs']:.3e} "
                  f"(max|z|={float(rz.abs().max()):.2f}) snr={mz['snr_db']:.1f} | m_p max_abs={mm['max_abs']:.3e} snr={mm['snr_db']:.1f}", flush=True)
            del net
    for k in ("RVC_B200_ENC_TC", "RVC_B200_FLOW_TC", "RVC_B200_PAIR_SPLIT", "RVC_B200_ENC_SPLIT"):
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()
