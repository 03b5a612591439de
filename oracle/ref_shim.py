"""Import shim for running the UNMODIFIED upstream reference.

TEST / BASELINE INFRASTRUCTURE ONLY (golden-vector generation, oracle validation, the CPU
reference arm of bench.py).  The reference lives read-only at /root/reference in the build
container and does not exist on the GPU box; scripts/install_ref.py puts the unmodified
files of the decoder path under the git-ignored baseline/_ref/, which does travel.  Nothing
may read /root/reference at run time on the GPU box: bench.py passes
``root=INSTALLED_REF`` explicitly; callers must check ``reference_available()`` /
``installed_reference_available()`` first.  The product (unitspeech_b200/) never imports this.

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

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED_REF = os.path.join(_REPO, "baseline", "_ref")      # scripts/install_ref.py: the unmodified files, git-ignored


def _has_decoder(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "unitspeech", "unitspeech.py"))


def _default_root() -> str:
    env = os.environ.get("UNITSPEECH_REFERENCE")
    if env:
        return env
    return "/root/reference" if _has_decoder("/root/reference") else INSTALLED_REF


REF_ROOT = _default_root()


def reference_available() -> bool:
    return _has_decoder(REF_ROOT)


def installed_reference_available() -> bool:
    """True when scripts/install_ref.py has put the unmodified reference under baseline/_ref (travels to the GPU box)."""
    return _has_decoder(INSTALLED_REF)


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


def load_reference(root: str = None):
    """Returns the reference module ``unitspeech.unitspeech`` (classes UnitSpeech, GradLogPEstimator2d...), imported
    from ``root`` (default: $UNITSPEECH_REFERENCE, /root/reference, else baseline/_ref)."""
    root = root or REF_ROOT
    if not _has_decoder(root):
        raise RuntimeError(f"reference not found under {root}")
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
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
