"""Generate tests/golden/*.npz by running the UNMODIFIED reference module (TEST INFRASTRUCTURE).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python -m oracle.make_golden

Each fixture stores only the case description (config name, seeds, shapes) and the reference's outputs; weights and
inputs are regenerated bit-identically from the seeds by ``oracle/cases.py``.  The reference's noise draws
(synthesizers.py:449 ``randn_like(m_p)``, :98 ``randn_like(sine_waves)``) are replaced by the case's recorded
noise tensors by patching ``torch.randn_like`` for the duration of the call; ``torch.rand`` (:87) is left alone since
its value is zeroed at :88.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import cases

REF_ROOT = "/root/reference"
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (config, use_f0, heated, disk_form, B, T, ragged, rate, weight_seed, input_seed, store_taps)
CASES = {
    "tiny_f0_ragged": ("tiny", True, True, False, 3, 40, True, None, 11, 1, True),
    "tiny_f0_cold": ("tiny", True, False, False, 1, 33, False, None, 12, 2, False),
    "tiny_nof0_ragged": ("tiny", False, True, False, 2, 36, True, None, 13, 3, True),
    "tiny_f0_rate": ("tiny", True, True, False, 2, 120, False, 0.75, 14, 4, False),
    "tiny_f0_diskform": ("tiny", True, True, True, 2, 24, False, None, 15, 5, False),
    "v2_48k_f0": ("v2_48k", True, True, False, 1, 48, False, None, 1234, 0, False),
    "v2_40k_f0_cold": ("v2_40k", True, False, False, 1, 40, False, None, 1234, 6, False),
    "v1_32k_f0_ragged": ("v1_32k", True, True, False, 2, 36, True, None, 1234, 7, False),
    "v2_40k_nof0": ("v2_40k", False, True, False, 2, 30, True, None, 1234, 8, False),
}


def import_reference():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present (only available in the build container)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sys.dont_write_bytecode = True
    from main.library.algorithm import synthesizers  # noqa

    # SURVEY 8c defect 1: the non-f0 Generator.forward is nested inside __init__; bind its code object (no source copy).
    G = synthesizers.Generator
    if "forward" not in G.__dict__:
        code = [c for c in G.__init__.__code__.co_consts if getattr(c, "co_name", "") == "forward"][0]
        G.forward = types.FunctionType(code, synthesizers.__dict__, "forward", (None,))
    return synthesizers


def build_reference(cfg, sd, use_f0):
    syn = import_reference()
    net = syn.Synthesizer(*cases.ctor_args(cfg), use_f0=use_f0, text_enc_hidden_dim=cfg["text_enc_hidden_dim"])
    del net.enc_q  # convert.py:565
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return net.eval().float()


def run_reference(net, inp, use_f0, rate=None):
    draws = [inp["eps_z"]] + ([inp["eps_src"]] if use_f0 else [])
    real = torch.randn_like
    it = iter(draws)

    def fake(t, *a, **k):
        n = next(it)
        if n.dim() == 3 and n.shape[1] > t.shape[1] and n.shape[2] == 1:
            n = n[:, : t.shape[1]]  # `rate` trims the head frames before the source draw (synthesizers.py:451-456)
        assert tuple(n.shape) == tuple(t.shape), (n.shape, t.shape)
        return n.to(t.dtype)

    torch.randn_like = fake
    try:
        with torch.no_grad():
            r = None if rate is None else torch.tensor(rate)
            o, x_mask, (z, z_p, m_p, logs_p) = net.infer(inp["phone"], inp["phone_lengths"], inp["pitch"], inp["nsff0"], inp["sid"], r)
    finally:
        torch.randn_like = real
    return dict(o=o, x_mask=x_mask, z=z, z_p=z_p, m_p=m_p, logs_p=logs_p)


def reference_taps(net, inp, use_f0):
    """Stage-level tensors through forward hooks."""
    taps = {}
    hooks = []
    dec = net.dec
    hooks.append(dec.conv_pre.register_forward_hook(lambda m, i, o: taps.__setitem__("conv_pre_nocond", o.detach().clone())))
    if use_f0:
        hooks.append(dec.m_source.register_forward_hook(lambda m, i, o: taps.__setitem__("har_source", o.detach().transpose(1, 2).clone())))
        for i, rb in enumerate(dec.resblocks):
            hooks.append(rb.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(f"rb{i}", o.detach().clone())))
    out = run_reference(net, inp, use_f0)
    for h in hooks:
        h.remove()
    return out, taps


def make_case(name):
    cname, use_f0, heated, disk, B, T, ragged, rate, wseed, iseed, store_taps = CASES[name]
    cfg = cases.CONFIGS[cname]
    sd = cases.synth_state_dict(cfg, use_f0=use_f0, seed=wseed, heated=heated)
    if disk:
        sd = cases.to_disk_form(sd)
    inp = cases.synth_inputs(cfg, B, T, seed=iseed, ragged=ragged, use_f0=use_f0)
    net = build_reference(cfg, sd, use_f0)
    if store_taps:
        out, taps = reference_taps(net, inp, use_f0)
    else:
        out, taps = run_reference(net, inp, use_f0, rate), {}
    arrays = {k: v.float().numpy() for k, v in out.items()}
    arrays.update({"tap_" + k: v.float().numpy() for k, v in taps.items()})
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
    return arrays


def main():
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    for name in CASES:
        a = make_case(name)
        o = a["o"]
        print(f"{name}: o{o.shape} mean={o.mean():+.4f} std={o.std():.4f} absmax={np.abs(o).max():.4f}")


if __name__ == "__main__":
    main()
