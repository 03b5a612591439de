"""CPU oracle for the BigVGAN generator (optional second stage).

TEST INFRASTRUCTURE ONLY (see oracle/unitspeech_oracle.py for the rules).  Functional fp32 restatement of
unitspeech/vocoder/models.py:121-191 (BigVGAN.forward), :18-69 (AMPBlock1), :78-112 (AMPBlock2),
activations.py:9-59,62-120 (Snake / SnakeBeta), alias_free_torch/act.py:8-28, resample.py:10-49, filter.py:28-95.
Parameters travel as a dict keyed by the reference generator's state_dict names AFTER remove_weight_norm()
(conv_pre.weight, ups.0.0.weight, resblocks.0.convs1.0.weight, resblocks.0.activations.0.act.alpha, ...), which is the
form unitspeech/util.py:174-181 (get_vocoder) runs.

Parity pinning: tests/golden/make_golden_vocoder.py runs the unmodified reference class (public 22 kHz / 80-band
hyper-parameters and two reduced configs) and commits the outputs; tests/test_oracle_vocoder.py checks this file
against them.  The reference ships no config file (README.md:63-64), so the hyper-parameters are the public
bigvgan_22khz_80band ones -- not pinned by the reference itself.
"""

from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

PUBLIC_22KHZ_80BAND = dict(
    num_mels=80, upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
    upsample_initial_channel=1536, resblock="1", resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], activation="snakebeta", snake_logscale=True)


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """alias_free_torch/filter.py:28-57 -> (kernel_size,)"""
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    time = (torch.arange(-half_size, half_size) + 0.5) if even else (torch.arange(kernel_size) - half_size)
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return filt / filt.sum()


def upsample2(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """UpSample1d(ratio 2, kernel 12) -- resample.py:10-33.  x: (B, C, T) -> (B, C, 2T)"""
    K, ratio = filt.numel(), 2
    pad = K // ratio - 1
    pad_left = pad * ratio + (K - ratio) // 2
    pad_right = pad * ratio + (K - ratio + 1) // 2
    C = x.shape[1]
    xp = F.pad(x, (pad, pad), mode="replicate")
    y = ratio * F.conv_transpose1d(xp, filt.view(1, 1, K).expand(C, -1, -1), stride=ratio, groups=C)
    return y[..., pad_left:-pad_right]


def downsample2(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """DownSample1d / LowPassFilter1d(stride 2, kernel 12, replicate padding) -- resample.py:36-49, filter.py:60-95."""
    K = filt.numel()
    pad_left, pad_right = K // 2 - int(K % 2 == 0), K // 2
    C = x.shape[1]
    xp = F.pad(x, (pad_left, pad_right), mode="replicate")
    return F.conv1d(xp, filt.view(1, 1, K).expand(C, -1, -1), stride=2, groups=C)


def snake(x: torch.Tensor, alpha: torch.Tensor, beta: torch.Tensor, logscale: bool, is_beta: bool) -> torch.Tensor:
    """activations.py:47-59 (Snake) / :107-120 (SnakeBeta)."""
    a = alpha.view(1, -1, 1)
    b = beta.view(1, -1, 1) if is_beta else a
    if logscale:
        a = torch.exp(a)
        b = torch.exp(b) if is_beta else a
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a) ** 2


def activation1d(p: Params, prefix: str, x: torch.Tensor, filt: torch.Tensor, h: dict) -> torch.Tensor:
    """Activation1d: up x2 -> snake -> down x2 -- alias_free_torch/act.py:23-28."""
    is_beta = h["activation"] == "snakebeta"
    alpha = p[prefix + ".act.alpha"]
    beta = p[prefix + ".act.beta"] if is_beta else alpha
    return downsample2(snake(upsample2(x, filt), alpha, beta, h["snake_logscale"], is_beta), filt)


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    return int((kernel_size * dilation - dilation) / 2)


def amp_block(p: Params, prefix: str, x: torch.Tensor, k: int, dil: Sequence[int], filt: torch.Tensor, h: dict) -> torch.Tensor:
    if h["resblock"] == "1":   # models.py:60-69
        for i, d in enumerate(dil):
            xt = activation1d(p, f"{prefix}.activations.{2 * i}", x, filt, h)
            xt = F.conv1d(xt, p[f"{prefix}.convs1.{i}.weight"], p[f"{prefix}.convs1.{i}.bias"], dilation=d,
                          padding=get_padding(k, d))
            xt = activation1d(p, f"{prefix}.activations.{2 * i + 1}", xt, filt, h)
            xt = F.conv1d(xt, p[f"{prefix}.convs2.{i}.weight"], p[f"{prefix}.convs2.{i}.bias"], padding=get_padding(k, 1))
            x = xt + x
        return x
    for i, d in enumerate(dil):    # AMPBlock2, models.py:105-112
        xt = activation1d(p, f"{prefix}.activations.{i}", x, filt, h)
        xt = F.conv1d(xt, p[f"{prefix}.convs.{i}.weight"], p[f"{prefix}.convs.{i}.bias"], dilation=d, padding=get_padding(k, d))
        x = xt + x
    return x


@torch.no_grad()
def bigvgan_forward(p: Params, mel: torch.Tensor, h: dict) -> torch.Tensor:
    """BigVGAN.forward -- models.py:169-191.  mel: (B, num_mels, T) -> (B, 1, T * prod(upsample_rates))"""
    filt = kaiser_sinc_filter1d(0.25, 0.3, 12)
    nk = len(h["resblock_kernel_sizes"])
    x = F.conv1d(mel, p["conv_pre.weight"], p["conv_pre.bias"], padding=3)
    for i, (u, ku) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(x, p[f"ups.{i}.0.weight"], p[f"ups.{i}.0.bias"], stride=u, padding=(ku - u) // 2)
        xs = None
        for j, (k, d) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            r = amp_block(p, f"resblocks.{i * nk + j}", x, k, d, filt, h)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = activation1d(p, "activation_post", x, filt, h)
    x = F.conv1d(x, p["conv_post.weight"], p["conv_post.bias"], padding=3)
    return torch.tanh(x)


def param_shapes(h: dict) -> Dict[str, tuple]:
    """state_dict of the generator after remove_weight_norm()."""
    s: Dict[str, tuple] = {}
    c0 = h["upsample_initial_channel"]
    s["conv_pre.weight"] = (c0, h["num_mels"], 7); s["conv_pre.bias"] = (c0,)
    nk = len(h["resblock_kernel_sizes"])
    ch = c0
    for i, (u, ku) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, ch = c0 // 2 ** i, c0 // 2 ** (i + 1)
        s[f"ups.{i}.0.weight"] = (cin, ch, ku); s[f"ups.{i}.0.bias"] = (ch,)
        for j, (k, d) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            pre = f"resblocks.{i * nk + j}"
            names = ("convs1", "convs2") if h["resblock"] == "1" else ("convs",)
            for nm in names:
                for l in range(len(d)):
                    s[f"{pre}.{nm}.{l}.weight"] = (ch, ch, k); s[f"{pre}.{nm}.{l}.bias"] = (ch,)
            for l in range(len(d) * len(names)):
                s[f"{pre}.activations.{l}.act.alpha"] = (ch,)
                if h["activation"] == "snakebeta":
                    s[f"{pre}.activations.{l}.act.beta"] = (ch,)
    s["activation_post.act.alpha"] = (ch,)
    if h["activation"] == "snakebeta":
        s["activation_post.act.beta"] = (ch,)
    s["conv_post.weight"] = (1, ch, 7); s["conv_post.bias"] = (1,)
    return s


def harness_params(h: dict, seed: int = 4321) -> Params:
    """Seeded weights: U(+-1/sqrt(fan_in)) convs (so activations stay O(1)), alpha/beta ~ N(0, 0.3) (log scale) or
    1 + N(0, 0.1) (linear), covering non-trivial snake frequencies."""
    g = torch.Generator().manual_seed(seed)
    p: Params = {}
    for name, shape in param_shapes(h).items():
        if name.endswith(".alpha") or name.endswith(".beta"):
            p[name] = torch.randn(shape, generator=g) * 0.3 if h["snake_logscale"] else 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".weight"):
            if name.startswith("ups."):
                fan_in = shape[0] * shape[2] / h["upsample_rates"][int(name.split(".")[1])]
            else:
                fan_in = shape[1] * shape[2]
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan_in, 1.0))
        else:
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return p
