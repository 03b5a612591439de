"""Drop-in for the reference's vocoder class ``unitspeech.vocoder.models.BigVGAN`` (models.py:121-201).

Same constructor (``BigVGAN(h)`` with the reference's config.json fields), the reference checkpoint's
``state_dict`` keys load unchanged (weight-norm ``weight_g`` / ``weight_v`` pairs are folded on load, the kaiser-sinc
``*.filter`` buffers are checked and dropped), ``remove_weight_norm()`` is accepted, and ``forward(mel)`` returns the
``(B, 1, T * hop)`` waveform.  The computation is ``usb_vocoder_forward`` of libunitspeech_b200.so; there is no
PyTorch or CPU fallback.
"""

from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional

import torch
from torch import nn

from . import abi


class AttrDict(dict):
    """unitspeech/vocoder/env.py:7-10"""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def param_shapes(h) -> Dict[str, tuple]:
    """state_dict of the reference generator after remove_weight_norm() (models.py:133-167), without buffers."""
    s: Dict[str, tuple] = {}
    c0 = h["upsample_initial_channel"]
    s["conv_pre.weight"] = (c0, h["num_mels"], 7)
    s["conv_pre.bias"] = (c0,)
    nk = len(h["resblock_kernel_sizes"])
    beta = h["activation"] == "snakebeta"
    ch = c0
    for i, ku in enumerate(h["upsample_kernel_sizes"]):
        cin, ch = c0 // 2 ** i, c0 // 2 ** (i + 1)
        s[f"ups.{i}.0.weight"] = (cin, ch, ku)
        s[f"ups.{i}.0.bias"] = (ch,)
        for j, (k, d) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            pre = f"resblocks.{i * nk + j}"
            names = ("convs1", "convs2") if str(h["resblock"]) == "1" else ("convs",)
            for nm in names:
                for l in range(len(d)):
                    s[f"{pre}.{nm}.{l}.weight"] = (ch, ch, k)
                    s[f"{pre}.{nm}.{l}.bias"] = (ch,)
            for l in range(len(d) * len(names)):
                s[f"{pre}.activations.{l}.act.alpha"] = (ch,)
                if beta:
                    s[f"{pre}.activations.{l}.act.beta"] = (ch,)
    s["activation_post.act.alpha"] = (ch,)
    if beta:
        s["activation_post.act.beta"] = (ch,)
    s["conv_post.weight"] = (1, ch, 7)
    s["conv_post.bias"] = (1,)
    return s


def fold_weight_norm(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """weight = g * v / ||v|| with the norm over every dim but 0 (torch.nn.utils.weight_norm, dim=0), i.e. what
    remove_weight_norm() leaves in ``weight`` (models.py:193-201); ``*.filter`` buffers are dropped."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in state.items():
        if k.endswith(".weight_g") or k.endswith(".filter"):
            continue
        if k.endswith(".weight_v"):
            g = state[k[:-1] + "g"].float()
            vf = v.float()
            norm = vf.flatten(1).norm(dim=1).view(-1, *([1] * (vf.dim() - 1)))
            out[k[:-2]] = vf * (g / norm)
        else:
            out[k] = v
    return out


def _attach(root: nn.Module, dotted: str, p: nn.Parameter) -> None:
    parts = dotted.split(".")
    m = root
    for name in parts[:-1]:
        if name not in m._modules:
            m.add_module(name, nn.Module())
        m = m._modules[name]
    m.register_parameter(parts[-1], p)


class BigVGAN(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.h = h if isinstance(h, dict) else AttrDict(vars(h))
        h = self.h
        if h["activation"] not in ("snake", "snakebeta"):
            raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")
        self.num_kernels = len(h["resblock_kernel_sizes"])
        self.num_upsamples = len(h["upsample_rates"])
        if len({len(d) for d in h["resblock_dilation_sizes"]}) != 1:
            raise ValueError("every resblock must have the same number of dilations")
        self.hop = int(math.prod(h["upsample_rates"]))
        for name, shape in param_shapes(h).items():
            if name.endswith(".alpha") or name.endswith(".beta"):   # activations.py:34-43,92-101 initial values
                init = torch.zeros(shape) if h.get("snake_logscale", False) else torch.ones(shape)
            elif name.endswith(".weight"):
                init = torch.randn(shape) * 0.01                        # init_weights, models.py:161,166
            else:
                init = torch.zeros(shape)
            _attach(self, name, nn.Parameter(init, requires_grad=False))
        self._handle: Optional[ctypes.c_void_p] = None
        self._handle_device = -1
        self._dirty = True
        self.max_frames_per_call = 8192      # utterances are grouped so that B*T stays below this (workspace bound)

    # ------------------------------------------------------------------ weights
    def remove_weight_norm(self):
        """No-op: weights are stored folded (models.py:193-201)."""
        return None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        folded = fold_weight_norm(dict(state_dict))
        self._dirty = True
        return super().load_state_dict(folded, strict=strict)

    def _apply(self, fn, *a, **kw):
        self._dirty = True
        return super()._apply(fn, *a, **kw)

    def _release(self):
        if self._handle is not None:
            abi.load_library().usb_vocoder_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_handle(self, dev: int):
        lib = abi.load_library()
        if self._handle is not None and not self._dirty and self._handle_device == dev:
            return self._handle
        self._release()
        h = self.h
        cfg = abi.UsbVocoderConfig()
        cfg.num_mels, cfg.n_upsamples = int(h["num_mels"]), self.num_upsamples
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            cfg.upsample_rates[i], cfg.upsample_kernel_sizes[i] = int(u), int(k)
        cfg.upsample_initial_channel = int(h["upsample_initial_channel"])
        cfg.resblock_type = int(h["resblock"])
        cfg.n_resblock_kernels = self.num_kernels
        cfg.n_dilations = len(h["resblock_dilation_sizes"][0])
        for j, (k, ds) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            cfg.resblock_kernel_sizes[j] = int(k)
            for l, d in enumerate(ds):
                cfg.resblock_dilations[j][l] = int(d)
        cfg.activation = 1 if h["activation"] == "snakebeta" else 0
        cfg.snake_logscale = int(bool(h.get("snake_logscale", False)))
        cfg.device = dev
        hp = ctypes.c_void_p()
        abi.check(lib.usb_vocoder_create(ctypes.byref(cfg), ctypes.byref(hp)))
        try:
            for key, t in self.state_dict().items():
                t = t.detach().float().cpu().contiguous()
                shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                abi.check(lib.usb_vocoder_load_param(hp, key.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()))
            abi.check(lib.usb_vocoder_finalize_params(hp))
        except Exception:
            lib.usb_vocoder_destroy(hp)
            raise
        self._handle, self._handle_device, self._dirty = hp, dev, False
        return hp

    @property
    def launch_count(self) -> int:
        return int(abi.load_library().usb_vocoder_launch_count(self._handle)) if self._handle is not None else 0

    @property
    def workspace_bytes(self) -> int:
        return int(abi.load_library().usb_vocoder_workspace_bytes(self._handle)) if self._handle is not None else 0

    @property
    def flops_per_call(self) -> float:
        return float(abi.load_library().usb_vocoder_flops_per_call(self._handle)) if self._handle is not None else 0.0

    def set_profiling(self, on: bool) -> None:
        """Per-kernel-class CUDA-event timing of the next forward calls (adds a sync; not for timed runs)."""
        if self._handle is None:
            raise abi.UsbError("run one forward before profiling")
        abi.check(abi.load_library().usb_vocoder_set_profiling(self._handle, int(bool(on))))

    def get_profile(self):
        """{class: (ms, algorithmic work, launches)} accumulated since set_profiling(True)."""
        ms, work, n = (ctypes.c_double * 3)(), (ctypes.c_double * 3)(), (ctypes.c_int64 * 3)()
        abi.check(abi.load_library().usb_vocoder_get_profile(self._handle, ms, work, n))
        return {k: (ms[i], work[i], int(n[i])) for i, k in enumerate(("conv_igemm", "snake_act", "other"))}

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x: torch.Tensor, mel_min: torch.Tensor = None, mel_max: torch.Tensor = None) -> torch.Tensor:
        """BigVGAN.forward (models.py:169-191).  x: (B, num_mels, T) -> (B, 1, T * hop), on x's device.

        With ``mel_min`` / ``mel_max`` (per-bin, any shape holding num_mels values, as stored in the decoder checkpoint)
        ``x`` is the decoder's NORMALISED output and the callers' de-normalisation (inference.py:140) is applied while
        the input is packed, i.e. ``voc(y, mel_min, mel_max) == voc((y + 1) / 2 * (mel_max - mel_min) + mel_min)``."""
        if x.dim() != 3 or x.shape[1] != self.h["num_mels"]:
            raise ValueError(f"expected (B, {self.h['num_mels']}, T) mel, got {tuple(x.shape)}")
        lib = abi.load_library()
        B, M, T = x.shape
        if not torch.cuda.is_available():
            raise abi.UsbError("unitspeech_b200 needs a B200 GPU; there is no CPU fallback")
        if x.is_cuda:
            dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        else:
            p = next(self.parameters())
            dev = p.device.index if p.is_cuda and p.device.index is not None else torch.cuda.current_device()
        hp = self._ensure_handle(dev)
        if (mel_min is None) != (mel_max is None):
            raise ValueError("mel_min and mel_max must be given together")
        if mel_min is not None:
            lo = mel_min.detach().float().reshape(-1).cpu().contiguous()
            hi = mel_max.detach().float().reshape(-1).cpu().contiguous()
            if lo.numel() != M or hi.numel() != M:
                raise ValueError(f"mel_min / mel_max must hold {M} values")
            abi.check(lib.usb_vocoder_set_input_denorm(hp, lo.data_ptr(), hi.data_ptr()))
        else:
            abi.check(lib.usb_vocoder_set_input_denorm(hp, None, None))
        xf = x.detach().float().contiguous()
        out = torch.empty(B, 1, T * self.hop, dtype=torch.float32, device=x.device)
        per = max(1, self.max_frames_per_call // max(T, 1))
        for b0 in range(0, B, per):
            b1 = min(B, b0 + per)
            if x.is_cuda:
                with torch.cuda.device(dev):
                    abi.check(lib.usb_vocoder_forward(hp, xf[b0:b1].data_ptr(), b1 - b0, T, out[b0:b1].data_ptr(),
                                                      int(torch.cuda.current_stream(dev).cuda_stream)))
            else:
                abi.check(lib.usb_vocoder_forward_host(hp, xf[b0:b1].data_ptr(), b1 - b0, T, out[b0:b1].data_ptr()))
        return out.to(x.dtype)


def get_vocoder(config_path, checkpoint, device):
    """unitspeech/util.py:174-181"""
    import json
    with open(config_path) as f:
        hps = AttrDict(json.load(f))
    vocoder = BigVGAN(hps)
    vocoder.load_state_dict(torch.load(checkpoint, map_location="cpu")["generator"])
    _ = vocoder.to(device).eval()
    vocoder.remove_weight_norm()
    return vocoder
