"""Generate tests/golden/mel_front.npz from the UNMODIFIED reference ``MelSpectrogram`` (TEST INFRASTRUCTURE; build container only).

    python -m oracle.make_mel_golden

main/library/predictors/RMVPE.py is imported as it lies; its missing dependency ``librosa.filters.mel`` (which only fills the module's ``mel_basis``
buffer, :146-148) is stubbed with oracle/mel_oracle.mel_filterbank.  The module is built exactly as ``RMVPE.__init__`` builds it (:204) and called as
``infer_from_audio`` calls it (:224).  The fixture stores the seeds and the reference's log-mel output; inputs are regenerated from the seeds.
"""
import os
import sys
import types

import numpy as np
import torch

from . import mel_oracle

REF_ROOT = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mel_front.npz")
# name -> (samples, seed): 1.0 s, a length that is not a multiple of the hop, shorter than one FFT window's half (reflect padding both ends), 5 s
CASES = {"mel_1s": (16000, 41), "mel_ragged": (23117, 42), "mel_short": (700, 43), "mel_5s": (80000, 44)}


def reference_mel_module():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sys.dont_write_bytecode = True
    lib, filt = types.ModuleType("librosa"), types.ModuleType("librosa.filters")
    filt.mel = mel_oracle.mel_filterbank
    lib.filters = filt
    sys.modules["librosa"], sys.modules["librosa.filters"] = lib, filt
    sys.modules.pop("main.library.predictors.RMVPE", None)       # (a previous import may have bound another stub)
    from main.library.predictors import RMVPE as mod
    return mod.MelSpectrogram(False, mod.N_MELS, 16000, 1024, 160, None, 30, 8000)      # RMVPE.py:204


def main():
    m = reference_mel_module()
    out = {"mel_basis": m.mel_basis.numpy()}
    for name, (n, seed) in CASES.items():
        audio = mel_oracle.synthetic_audio(n, seed)
        with torch.no_grad():
            out[name] = m(audio, center=True).numpy()
        out[name + ".case"] = np.array([n, seed])
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
