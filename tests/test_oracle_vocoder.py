"""Pins oracle/bigvgan_oracle.py against outputs of the unmodified reference BigVGAN (tests/golden/make_golden_vocoder.py)."""

import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from oracle import bigvgan_oracle as V  # noqa: E402

from make_golden_vocoder import CONFIGS  # noqa: E402


@pytest.mark.parametrize("name", list(CONFIGS))
def test_vocoder_oracle_matches_reference(golden_dir, name):
    h, B, T = CONFIGS[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    p = V.harness_params(h)
    gen = torch.Generator().manual_seed(17)
    mel = torch.randn(B, h["num_mels"], T, generator=gen) * 2 - 4
    out = V.bigvgan_forward(p, mel, h)
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 2e-5


def test_kaiser_sinc_filter_known_answer(golden_dir):
    g = np.load(os.path.join(golden_dir, "vocoder_small.npz"))
    f = V.kaiser_sinc_filter1d(0.25, 0.3, 12)
    assert np.array_equal(f.numpy(), g["filt"])
    assert abs(float(f.sum()) - 1.0) < 1e-6 and torch.allclose(f, f.flip(0))
