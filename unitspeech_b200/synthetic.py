"""Seeded synthetic weights and inputs for benchmarks and smoke runs (there is no network for checkpoints or data).

Random-init weights of the reference architecture with the three fixes that make a default init usable
(SURVEY F3/F4): Rezero g ~ N(0, 0.1) (the reference initialises it to 0, which disables attention), spk_uncon ~ N(0, 1)
(zero-init gives 0/0 under CFG, unitspeech/unitspeech.py:358), text_uncon ~ N(0, 0.5); final_conv is scaled so that the
expansive untrained sampler stays O(1).  tests/test_synthetic.py checks these are the same tensors the parity
harness (oracle/) uses.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch


def random_init_state_dict(decoder: torch.nn.Module, seed: int = 1234, out_scale: float = 1.0 / 512) -> Dict[str, torch.Tensor]:
    """state_dict for `decoder` (a unitspeech_b200.UnitSpeech), generated in state_dict key order."""
    g = torch.Generator().manual_seed(seed)
    # canonical generation order (forward order of the network): uncon vectors, time MLP, down path, middle, up path, final
    order = ["text_uncon", "spk_uncon", "estimator.mlp.", "estimator.downs.", "estimator.mid_block1.",
             "estimator.mid_attn.", "estimator.mid_block2.", "estimator.ups.", "estimator.final_block.",
             "estimator.final_conv."]
    rank = lambda k: next(i for i, pre in enumerate(order) if k.startswith(pre))  # noqa: E731
    sd = decoder.state_dict()
    shapes = {k: tuple(sd[k].shape) for k in sorted(sd, key=rank)}  # stable: keeps registration order inside a group
    p: Dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        if name == "text_uncon":
            p[name] = torch.randn(shape, generator=g) * 0.5
        elif name == "spk_uncon":
            p[name] = torch.randn(shape, generator=g)
        elif name.endswith(".fn.g"):
            p[name] = torch.randn(shape, generator=g) * 0.1
        elif ".block.1." in name:  # GroupNorm affine
            base = 1.0 if name.endswith("weight") else 0.0
            p[name] = base + 0.1 * torch.randn(shape, generator=g)
        else:
            wshape = shapes.get(name[:-4] + "weight", shape) if name.endswith("bias") else shape
            if len(wshape) == 4 and wshape[2] == 4:  # ConvTranspose2d weight is (Cin, Cout, 4, 4)
                fan_in = wshape[1] * wshape[2] * wshape[3]
            else:
                fan_in = 1
                for d in wshape[1:]:
                    fan_in *= d
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    p["estimator.final_conv.weight"] = p["estimator.final_conv.weight"] * out_scale
    p["estimator.final_conv.bias"] = p["estimator.final_conv.bias"] * out_scale
    return p


def synthetic_inputs(B: int, T: int, n_steps: int, n_feats: int = 80, spk_emb_dim: int = 256, seed: int = 0,
                     scale: float = 1.0 / 512, lengths: Optional[Sequence[int]] = None):
    """cond ~ N(0,1).clamp(-1,1); spk L2-normalised; z and per-step noise ~ scale*N(0,1); mask from lengths."""
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn(B, n_feats, T, generator=g).clamp(-1, 1)
    spk = torch.randn(B, 1, spk_emb_dim, generator=g)
    spk = spk / spk.norm(dim=-1, keepdim=True)
    z = torch.randn(B, n_feats, T, generator=g) * scale
    noise = torch.randn(n_steps, B, n_feats, T, generator=g) * scale
    if lengths is None:
        mask = torch.ones(B, 1, T)
    else:
        ar = torch.arange(T).unsqueeze(0)
        mask = (ar < torch.tensor(list(lengths)).unsqueeze(1)).float().unsqueeze(1)
    return z, mask, cond, spk, noise


# Hyper-parameters of the public 22 kHz / 80-band BigVGAN generator (the reference ships no vocoder config.json,
# README.md:63-64 points at the public checkpoint).
PUBLIC_VOCODER_CONFIG = dict(
    num_mels=80, upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
    upsample_initial_channel=1536, resblock="1", resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], activation="snakebeta", snake_logscale=True)


def vocoder_state(h: dict, seed: int = 4321) -> Dict[str, torch.Tensor]:
    """Seeded generator weights in the folded (post remove_weight_norm) form: convs U(+-1/sqrt(fan_in)) so the
    activations stay O(1), snake alpha/beta ~ N(0, 0.3) in log scale (or 1 + N(0, 0.1) linear).  The same tensors as
    the parity harness (oracle/bigvgan_oracle.harness_params; tests/test_synthetic.py checks)."""
    from .vocoder import param_shapes
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(h).items():
        if name.endswith(".alpha") or name.endswith(".beta"):
            p[name] = torch.randn(shape, generator=g) * 0.3 if h.get("snake_logscale") else 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".weight"):
            if name.startswith("ups."):
                fan_in = shape[0] * shape[2] / h["upsample_rates"][int(name.split(".")[1])]
            else:
                fan_in = shape[1] * shape[2]
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1.0))
        else:
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return p
