"""Import shim for running the UNMODIFIED upstream reference in the build container.

TEST INFRASTRUCTURE ONLY (golden-vector generation and oracle validation).
The reference lives read-only at /root/reference and does not exist on the GPU
box, so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call
``load_reference()``; callers must check ``reference_available()`` first.

Why a shim: unitspeech/util.py imports librosa, matplotlib, phonemizer and
fairseq-backed modules at top level (unitspeech/util.py:8-18) and
conf/hydra_config.py does not import on Python >= 3.11; none of that is on the
decoder path, so those modules are replaced by inert stubs.
"""

from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("UNITSPEECH_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "unitspeech", "unitspeech.py"))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return None


_STUBBED = [
    "librosa", "librosa.util", "librosa.filters", "matplotlib", "matplotlib.pyplot", "matplotlib.pylab",
    "phonemizer", "phonemizer.backend", "conf", "conf.hydra_config",
    "unitspeech.speaker_encoder.ecapa_tdnn", "unitspeech.textlesslib.textless.data.speech_encoder",
]


def load_reference():
    """Returns the reference module ``unitspeech.unitspeech`` (classes UnitSpeech, GradLogPEstimator2d...)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in _STUBBED:
        if name not in sys.modules:
            m = _Stub(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            m.__path__ = []
            sys.modules[name] = m
    return importlib.import_module("unitspeech.unitspeech")


def run_reference_per_utterance(dec, z, mask, cond, spk, noise, n_steps, tg, sg, trace=None):
    """The reference's own reverse_diffusion, one utterance (B=1) at a time, with the
    per-step noise injected by temporarily replacing ``torch.randn`` (the only randn
    inside reverse_diffusion is unitspeech/unitspeech.py:367)."""
    import torch

    U = sys.modules["unitspeech.unitspeech"]
    outs = []
    real_randn = torch.randn
    for b in range(z.shape[0]):
        queue = [noise[i, b:b + 1] for i in range(n_steps)]

        def fake_randn(*a, **k):
            return queue.pop(0).clone()

        U.torch.randn = fake_randn
        try:
            outs.append(dec.reverse_diffusion(z[b:b + 1], mask[b:b + 1], cond[b:b + 1], spk[b:b + 1], n_steps,
                                              text_gradient_scale=tg, spk_gradient_scale=sg))
        finally:
            U.torch.randn = real_randn
    return torch.cat(outs, 0)
