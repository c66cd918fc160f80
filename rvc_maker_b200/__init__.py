"""rvc_maker_b200 — B200-native (sm_100a) implementation of RVC-MAKER's ``Synthesizer.infer`` hot path.

``Synthesizer`` is a drop-in for ``main.library.algorithm.synthesizers.Synthesizer`` on the inference path (same
constructor, ``state_dict`` keys and ``infer`` signature); its arithmetic runs in hand-written CUDA kernels reached through
the C ABI declared in ``include/rvc_b200.h`` (``librvc_b200.so``).  No CPU fallback.
"""
from .synthesizer import Synthesizer  # noqa: F401
from . import configs  # noqa: F401

__all__ = ["Synthesizer", "configs"]
