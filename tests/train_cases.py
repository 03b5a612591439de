"""Shared definitions of the fine-tune parity cases (used by tests/golden/make_golden_grads.py and the tests)."""

import torch

from oracle import unitspeech_oracle as O

# name: (dim, dim_mults, B, T, lengths, t values, out_scale, lr, K iterations)
CASES = {
    "grads_d64": (64, (1, 2), 2, 16, (16, 11), (0.37, 0.81), 4.0, 2e-5, 4),
    "grads_full": (128, (1, 2, 4, 8), 2, 24, (24, 19), (0.05, 0.6), 4.0, 2e-5, 3),
}
# gradients stored in full in the golden files
FULL_KEYS = ["estimator.final_conv.weight", "estimator.final_conv.bias", "estimator.downs.0.0.res_conv.weight",
             "estimator.downs.0.2.fn.g", "estimator.mid_attn.fn.g", "estimator.mlp.2.bias",
             "estimator.downs.0.0.block1.block.1.weight"]


def case_inputs(B, T, lengths, seed=21):
    _, mask, cond, spk, _ = O.harness_inputs(B, T, 2, seed=seed, lengths=lengths)
    g = torch.Generator().manual_seed(seed + 1)
    x0 = (torch.randn(B, 80, T, generator=g) * 0.5).clamp(-1, 1) * mask
    return x0, mask, cond, spk


def reference_z(shape, i):
    """The draw forward_diffusion makes (unitspeech/unitspeech.py:381) when iteration i runs under manual_seed(77 + i)."""
    torch.manual_seed(77 + i)
    return torch.randn(shape)


def projection(name, shape):
    """Seeded N(0,1) tensor per parameter name: <grad, projection> is stored instead of the full gradient."""
    seed = sum(ord(c) * (i + 1) for i, c in enumerate(name)) % (2 ** 31)
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))
