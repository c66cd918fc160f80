"""Write tests/golden/module_boundary.json from the UNMODIFIED reference class (TEST INFRASTRUCTURE).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python -m oracle.make_boundary_fixture

Records what the drop-in must mirror of ``Synthesizer`` (main/library/algorithm/synthesizers.py:396-432) without shipping any
reference source: the constructor's parameter list with defaults (``inspect.signature``), the ``infer`` parameter list, and the
``state_dict`` key -> shape map (including the train-only ``enc_q`` that the caller deletes, convert.py:565) for every preset
the parity suite uses.
"""
from __future__ import annotations

import inspect
import json
import os

from . import cases
from .make_golden import GOLDEN_DIR, import_reference

# (config, use_f0)
PRESETS = [("v2_40k", True), ("v2_48k", True), ("v1_32k", True), ("v2_40k", False), ("tiny", True), ("tiny", False)]


def _params(fn):
    out = []
    for p in list(inspect.signature(fn).parameters.values())[1:]:   # drop self
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([p.name, p.kind.name, d])
    return out


def main():
    syn = import_reference()
    fx = {"ctor": _params(syn.Synthesizer.__init__), "infer": _params(syn.Synthesizer.infer), "state_dict": {}}
    for cname, use_f0 in PRESETS:
        cfg = cases.CONFIGS[cname]
        net = syn.Synthesizer(*cases.ctor_args(cfg), use_f0=use_f0, text_enc_hidden_dim=cfg["text_enc_hidden_dim"])
        fx["state_dict"][f"{cname}/{'f0' if use_f0 else 'nof0'}"] = {k: list(v.shape) for k, v in net.state_dict().items()}
    path = os.path.join(GOLDEN_DIR, "module_boundary.json")
    with open(path, "w") as f:
        json.dump(fx, f, indent=0, sort_keys=True)
    print(path, {k: len(v) for k, v in fx["state_dict"].items()})


if __name__ == "__main__":
    main()
