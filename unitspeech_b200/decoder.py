"""Drop-in `UnitSpeech` decoder for the reverse-diffusion path, backed by libunitspeech_b200.so (sm_100a CUDA).

Mirrors the call surface of the reference class unitspeech.unitspeech.UnitSpeech (unitspeech/unitspeech.py:220):
same constructor, same `state_dict` key names and shapes (so reference checkpoints load with
``load_state_dict(ckpt["model"])``, inference.py:66-73), and the entry points the reference's callers use on this
path: ``forward`` / ``reverse_diffusion`` (scripts/text_to_speech.py:42), ``execute_text_to_speech``
(inference.py:128, evaluation/evaluation_generate_samples.py:325) and ``estimator(x, mask, mu, t, spk_emb)``.

PyTorch is used for tensor plumbing only: parameters are handed to the C library, every FLOP of the estimator and
the sampler update runs in the hand-written kernels.  There is no fallback path.

Differences from the reference, all deliberate:
  * a batch is sampled as "the reference's batch-1 call per utterance" (the reference's own B>1 schedule is wrong and
    its CFG path crashes for B>1);
  * ``noise=`` injects the per-step randn draws for reproducibility; with ``noise=None`` they are drawn with
    ``torch.randn`` in the reference's order (one (B, n_feats, T) draw per step);
  * tensors may live on the CPU: the call then goes through the host-buffer entry of the library (copies included)
    and returns a CPU tensor;
  * ``n_timesteps=1`` raises ValueError (the reference crashes with an indexing error).
``forward_diffusion`` / ``loss_t`` / ``compute_loss`` / ``fine_tune`` evaluate the fine-tuning objective on the same
kernels; with autograd enabled the loss back-propagates through the CUDA backward pass (unitspeech_b200/training.py), so the
reference's fine-tune loop (finetune.py:131-165) runs unchanged, and ``fused_finetuner()`` is the all-in-library fast path
(backward + clip_grad_norm_ + Adam).
"""

from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence

import numpy as np
import torch

from . import abi, schedule
from .util import fix_len_compatibility, generate_path, sequence_mask


class _Holder(torch.nn.Module):
    """Parameter container; exists only so state_dict() has the reference's key names."""


def _uniform(shape, bound):
    return torch.nn.Parameter((torch.rand(shape) * 2 - 1) * bound)


def _conv_params(mod: torch.nn.Module, cout: int, cin: int, k: int, bias: bool = True, transposed: bool = False):
    # torch default init of Conv2d/ConvTranspose2d: U(+-1/sqrt(fan_in)) for weight and bias
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    fan_in = shape[1] * k * k
    bound = 1.0 / math.sqrt(fan_in)
    mod.weight = _uniform(shape, bound)
    if bias:
        mod.bias = _uniform((cout,), bound)
    return mod


def _linear_params(mod: torch.nn.Module, cout: int, cin: int):
    bound = 1.0 / math.sqrt(cin)
    mod.weight = _uniform((cout, cin), bound)
    mod.bias = _uniform((cout,), bound)
    return mod


def _gn_params(mod: torch.nn.Module, c: int):
    mod.weight = torch.nn.Parameter(torch.ones(c))
    mod.bias = torch.nn.Parameter(torch.zeros(c))
    return mod


def _seq(**children) -> _Holder:
    h = _Holder()
    for name, child in children.items():
        h.add_module(name.lstrip("_"), child)
    return h


def _block(cin: int, cout: int) -> _Holder:
    # Block.block = Sequential(Conv2d, GroupNorm, Mish) -> keys block.0.*, block.1.*   (unitspeech.py:46-51)
    return _seq(block=_seq(_0=_conv_params(_Holder(), cout, cin, 3), _1=_gn_params(_Holder(), cout)))


def _resnet(cin: int, cout: int, temb: int) -> _Holder:
    r = _seq(mlp=_seq(_1=_linear_params(_Holder(), cout, temb)), block1=_block(cin, cout), block2=_block(cout, cout))
    if cin != cout:
        r.add_module("res_conv", _conv_params(_Holder(), cout, cin, 1))
    return r


def _attn(c: int, hidden: int = 128) -> _Holder:
    # Residual(Rezero(LinearAttention)) -> keys fn.g, fn.fn.to_qkv.weight, fn.fn.to_out.{weight,bias}
    la = _seq(to_qkv=_conv_params(_Holder(), hidden * 3, c, 1, bias=False), to_out=_conv_params(_Holder(), c, hidden, 1))
    rz = _seq(fn=la)
    rz.g = torch.nn.Parameter(torch.zeros(1))
    return _seq(fn=rz)


class GradLogPEstimator2d(torch.nn.Module):
    """Parameter tree of the reference U-Net (unitspeech/unitspeech.py:124-162); forward runs on the CUDA library."""

    def __init__(self, dim, dim_mults=(1, 2, 4), groups=8, pe_scale=1000, spk_emb_dim=0):
        super().__init__()
        self.dim, self.dim_mults, self.groups, self.pe_scale = dim, tuple(dim_mults), groups, pe_scale
        temb = dim + spk_emb_dim
        self.mlp = _seq(_0=_linear_params(_Holder(), dim * 4, dim), _2=_linear_params(_Holder(), dim, dim * 4))
        dims = [2] + [dim * m for m in dim_mults]
        in_out = list(zip(dims[:-1], dims[1:]))
        self.downs = torch.nn.ModuleList()
        self.ups = torch.nn.ModuleList()
        for ind, (ci, co) in enumerate(in_out):
            last = ind >= len(in_out) - 1
            down = _Holder() if last else _seq(conv=_conv_params(_Holder(), co, co, 3))
            self.downs.append(torch.nn.ModuleList([_resnet(ci, co, temb), _resnet(co, co, temb), _attn(co), down]))
        mid = dims[-1]
        self.mid_block1 = _resnet(mid, mid, temb)
        self.mid_attn = _attn(mid)
        self.mid_block2 = _resnet(mid, mid, temb)
        for ci, co in reversed(in_out[1:]):
            up = _seq(conv=_conv_params(_Holder(), ci, ci, 4, transposed=True))
            self.ups.append(torch.nn.ModuleList([_resnet(co * 2, ci, temb), _resnet(ci, ci, temb), _attn(ci), up]))
        self.final_block = _block(dim, dim)
        self.final_conv = _conv_params(_Holder(), 1, dim, 1)
        object.__setattr__(self, "_owner", None)

    @torch.no_grad()
    def forward(self, x, mask, mu, t, spk_emb=None):
        """GradLogPEstimator2d.forward (unitspeech/unitspeech.py:164-201): (B,F,T),(B,1,T),(B,F,T),(B,),(B,1,S)."""
        owner = self._owner
        if owner is None:
            raise RuntimeError("estimator is not attached to a UnitSpeech decoder")
        return owner._estimator_forward(x, mask, mu, t, spk_emb)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class UnitSpeech(torch.nn.Module):
    def __init__(self, n_feats, dim, dim_mults, beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=0):
        super().__init__()
        self.n_feats = n_feats
        self.dim = dim
        self.dim_mults = tuple(dim_mults)
        self.beta_min = beta_min
        self.beta_max = beta_max
        self.pe_scale = pe_scale
        self.spk_emb_dim = spk_emb_dim
        self.text_uncon = torch.nn.Parameter(torch.zeros(1, n_feats, 1))
        self.spk_uncon = torch.nn.Parameter(torch.zeros(1, 1, spk_emb_dim))
        self.estimator = GradLogPEstimator2d(dim, dim_mults=dim_mults, pe_scale=pe_scale, spk_emb_dim=spk_emb_dim)
        object.__setattr__(self.estimator, "_owner", self)
        self.max_rows_frames = 96 * 1000   # (CFG branches x utterances) x frames per library call; see reverse_diffusion
        self._graph_mode = -1              # CUDA-graph replay of the sampler step: -1 auto (small calls), 0 off, 1 on
        self._splitk_mode = -1             # split-K of the few-tile convolutions: same encoding
        self._handle = None
        self._handle_device = None
        self._weights_version = 0
        self._synced_version = -1
        self._tuner = None

    # ------------------------------------------------------------------ reference conveniences (unitspeech/base.py)
    @property
    def nparams(self):
        return sum(int(np.prod(p.detach().cpu().numpy().shape)) for _, p in self.named_parameters() if p.requires_grad)

    def relocate_input(self, x: list):
        device = next(self.parameters()).device
        return [t.to(device) if isinstance(t, torch.Tensor) and t.device != device else t for t in x]

    # ------------------------------------------------------------------ native handle management
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._weights_version += 1
        return out

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._weights_version += 1
        return out

    def mark_weights_changed(self):
        """Call after modifying parameters in place; the next call re-uploads them."""
        self._weights_version += 1

    def _release(self):
        if self._handle is not None:
            abi.load_library().usb_destroy(self._handle)
            self._handle = None
            self._denorm_key = self._denorm_handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _device_index(self, like: Optional[torch.Tensor] = None) -> int:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            if like is not None and like.device.type == "cuda":
                dev = like.device
            elif torch.cuda.is_available():
                dev = torch.device("cuda", torch.cuda.current_device())
            else:
                raise abi.UsbError("unitspeech_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        return dev.index if dev.index is not None else torch.cuda.current_device()

    def _ensure_handle(self, like: Optional[torch.Tensor] = None):
        lib = abi.load_library()
        dev = self._device_index(like)
        if self._handle is not None and self._synced_version == self._weights_version and self._handle_device == dev:
            return self._handle
        self._release()
        cfg = abi.UsbConfig()
        cfg.n_feats, cfg.dim, cfg.n_mults = self.n_feats, self.dim, len(self.dim_mults)
        for i, m in enumerate(self.dim_mults):
            cfg.dim_mults[i] = int(m)
        cfg.groups, cfg.spk_emb_dim = 8, self.spk_emb_dim
        cfg.pe_scale, cfg.beta_min, cfg.beta_max, cfg.device = float(self.pe_scale), self.beta_min, self.beta_max, dev
        hp = ctypes.c_void_p()
        abi.check(lib.usb_create(ctypes.byref(cfg), ctypes.byref(hp)))
        try:
            params = {k: _f32c(v).cpu() for k, v in self.state_dict().items()}
            params["__posemb_freqs"] = schedule.posemb_freqs(self.dim)
            for key, t in params.items():
                shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                abi.check(lib.usb_load_param(hp, key.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()))
            abi.check(lib.usb_finalize_params(hp))
            abi.check(lib.usb_set_graph_mode(hp, int(self._graph_mode)))
            abi.check(lib.usb_set_splitk_mode(hp, int(self._splitk_mode)))
        except Exception:
            lib.usb_destroy(hp)
            raise
        self._handle, self._handle_device, self._synced_version = hp, dev, self._weights_version
        return hp

    @staticmethod
    def _stream(device_index: int) -> int:
        return int(torch.cuda.current_stream(device_index).cuda_stream)

    @property
    def launch_count(self) -> int:
        return int(abi.load_library().usb_launch_count(self._handle)) if self._handle is not None else 0

    @property
    def workspace_bytes(self) -> int:
        return int(abi.load_library().usb_workspace_bytes(self._handle)) if self._handle is not None else 0

    @property
    def graph_mode(self) -> int:
        """-1: replay one captured CUDA graph per diffusion step when the call is small enough to be launch-bound
        ((CFG branches x utterances) x frames <= 12288, e.g. the reference callers' one-utterance calls); 0: never; 1: always.
        Results are bit-identical either way (same kernels, same arguments)."""
        return self._graph_mode

    @graph_mode.setter
    def graph_mode(self, mode: int) -> None:
        if int(mode) not in (-1, 0, 1):
            raise ValueError("graph_mode must be -1 (auto), 0 (off) or 1 (on)")
        self._graph_mode = int(mode)
        if self._handle is not None:
            abi.check(abi.load_library().usb_set_graph_mode(self._handle, self._graph_mode))

    @property
    def splitk_mode(self) -> int:
        """-1: small (latency-bound) calls split the K range of the level-2/3 convolutions over otherwise idle SMs, with a
        fixed-order reduction (deterministic; the split depends on the layer geometry only); 0: never; 1: always.  A call is
        bit-identical to the same utterances sampled alone as long as both run in the same mode (auto: both
        (CFG branches x utterances) x frames <= 12288, or both above); the two modes differ in the last fp32 bits."""
        return self._splitk_mode

    @splitk_mode.setter
    def splitk_mode(self, mode: int) -> None:
        if int(mode) not in (-1, 0, 1):
            raise ValueError("splitk_mode must be -1 (auto), 0 (off) or 1 (on)")
        self._splitk_mode = int(mode)
        if self._handle is not None:
            abi.check(abi.load_library().usb_set_splitk_mode(self._handle, self._splitk_mode))

    SPLITK_AUTO_ROWS_FRAMES = 3072      # kSplitKAutoRowsFrames of csrc/engine.cu

    def job_splitk_mode(self, rows_frames: int) -> int:
        """The explicit mode (0 / 1) the auto rule picks for a job of `rows_frames` = (CFG branches x utterances) x frames.
        Used to pin one decision for all parts of a job that is cut into micro-batches or shards."""
        return self._splitk_mode if self._splitk_mode != -1 else (1 if rows_frames <= self.SPLITK_AUTO_ROWS_FRAMES else 0)

    @property
    def graph_steps(self) -> int:
        """Diffusion steps executed as graph replays since the handle was created."""
        return int(abi.load_library().usb_graph_steps(self._handle)) if self._handle is not None else 0

    def saturation_count(self, reset: bool = False) -> int:
        """Number of kernel epilogue threads that clamped a value to the fp16 limit (+-65504) since the handle was
        created / last reset (saturate and flag instead of emitting inf).  Non-zero means the fp16 operand
        format was out of range for this checkpoint or input.  Synchronises with the device."""
        if self._handle is None:
            return 0
        n = ctypes.c_int64(0)
        abi.check(abi.load_library().usb_saturation_count(self._handle, ctypes.byref(n), int(bool(reset))))
        return int(n.value)

    def set_profiling(self, on: bool) -> None:
        """Per-kernel-class CUDA-event timing of the next reverse_diffusion calls (adds a sync; not for timed runs)."""
        abi.check(abi.load_library().usb_set_profiling(self._ensure_handle(), int(bool(on))))

    def get_profile(self):
        """{class: (ms, algorithmic work, launches)} accumulated since set_profiling()."""
        ms, work, n = (ctypes.c_double * 4)(), (ctypes.c_double * 4)(), (ctypes.c_int64 * 4)()
        abi.check(abi.load_library().usb_get_profile(self._ensure_handle(), ms, work, n))
        names = ("conv_igemm", "gn_apply", "attention", "other")
        return {k: (ms[i], work[i], int(n[i])) for i, k in enumerate(names)}

    # ------------------------------------------------------------------ estimator
    @torch.no_grad()
    def _estimator_forward(self, x, mask, mu, t, spk_emb):
        lib = abi.load_library()
        h = self._ensure_handle(x)
        dev = torch.device("cuda", self._handle_device)
        B, F, T = x.shape
        xd, mud = _f32c(x).to(dev), _f32c(mu).to(dev)
        md = _f32c(mask).to(dev).reshape(B, T)
        td = _f32c(t).to(dev).reshape(B)
        sd = _f32c(spk_emb).to(dev).reshape(B, self.spk_emb_dim)
        out = torch.empty_like(xd)
        with torch.cuda.device(dev):
            abi.check(lib.usb_estimator_forward(h, xd.data_ptr(), mud.data_ptr(), md.data_ptr(), td.data_ptr(),
                                                sd.data_ptr(), out.data_ptr(), B, T, self._stream(dev.index)))
        return out.to(x.device)

    # ------------------------------------------------------------------ sampler
    @torch.no_grad()
    def reverse_diffusion(self, z, mask, cond, spk_emb, n_timesteps, text_gradient_scale=0.0, spk_gradient_scale=0.0,
                          noise: Optional[torch.Tensor] = None, trace: bool = False,
                          denorm: Optional[Sequence[torch.Tensor]] = None):
        """UnitSpeech.reverse_diffusion (unitspeech/unitspeech.py:334-374), per-utterance batch-1 semantics.

        z, cond: (B, n_feats, T); mask: (B, 1, T); spk_emb: (B, 1, spk_emb_dim); noise: optional (n, B, n_feats, T).
        Returns (B, n_feats, T) on z's device; with trace=True also the (n, B, n_feats, T) x_t after every step.
        ``denorm=(mel_min, mel_max)`` (per-bin, n_feats values each, as stored in the decoder checkpoint) fuses the
        callers' mel de-normalisation ``(y + 1) / 2 * (mel_max - mel_min) + mel_min`` (inference.py:140) into the last
        sampler step: the returned tensor is then the log-mel the vocoder takes (the trace stays normalised).
        """
        if n_timesteps < 2:
            raise ValueError("n_timesteps must be >= 2 (the reference fails for 1)")
        lib = abi.load_library()
        B, F, T = z.shape
        if F != self.n_feats:
            raise ValueError(f"expected {self.n_feats} mel bins, got {F}")
        if T % (2 ** (len(self.dim_mults) - 1)):
            raise ValueError("T must be a multiple of 2**(len(dim_mults)-1) (use fix_len_compatibility)")
        if noise is None:
            # the reference draws one randn per step from the global generator of z's device (:367)
            noise = torch.stack([torch.randn(z.shape, dtype=z.dtype, device=z.device) for _ in range(n_timesteps)])
        if tuple(noise.shape) != (n_timesteps, B, F, T):
            raise ValueError("noise must have shape (n_timesteps, B, n_feats, T)")
        # micro-batching: the workspace grows with (CFG branches x utterances x frames); split large batches so one call
        # stays within `max_rows_frames` (about 35 GB at the default) -- utterances are independent, so this is exact
        nb = 1 + (1 if float(text_gradient_scale) > 0 else 0) + (1 if float(spk_gradient_scale) > 0 else 0)
        max_b = max(1, int(self.max_rows_frames) // (nb * T))
        if B > max_b:
            outs, traces = [], []
            # the split-K decision (which fixes the fp32 summation order) is taken for the JOB, so that every micro-batch
            # -- including a short last one -- rounds like the others
            saved = self._splitk_mode
            self.splitk_mode = self.job_splitk_mode(nb * B * T)
            try:
                for b0 in range(0, B, max_b):
                    sl = slice(b0, min(B, b0 + max_b))
                    r = self.reverse_diffusion(z[sl], mask[sl], cond[sl], spk_emb[sl], n_timesteps, text_gradient_scale,
                                               spk_gradient_scale, noise=noise[:, sl], trace=trace, denorm=denorm)
                    outs.append(r[0] if trace else r)
                    if trace:
                        traces.append(r[1])
            finally:
                self.splitk_mode = saved
            out = torch.cat(outs, 0)
            return (out, torch.cat(traces, 1)) if trace else out
        coef = schedule.step_coefficients(n_timesteps, self.beta_min, self.beta_max).contiguous()
        times = schedule.step_times(n_timesteps).contiguous()
        tg, sg = float(text_gradient_scale), float(spk_gradient_scale)
        h = self._ensure_handle(z)
        dev = torch.device("cuda", self._handle_device)
        on_host = z.device.type != "cuda"
        self._set_denorm(h, denorm)
        with torch.cuda.device(dev):
            stream = self._stream(dev.index)
            if on_host and not trace:
                zc, cc, nc = _f32c(z), _f32c(cond), _f32c(noise)
                mc = _f32c(mask).reshape(B, T)
                sc = _f32c(spk_emb).reshape(B, self.spk_emb_dim)
                out = torch.empty_like(zc)
                abi.check(lib.usb_reverse_diffusion_host(h, zc.data_ptr(), cc.data_ptr(), mc.data_ptr(), sc.data_ptr(),
                                                         nc.data_ptr(), coef.data_ptr(), times.data_ptr(), n_timesteps,
                                                         tg, sg, out.data_ptr(), B, T, stream))
                return out
            zd, cd, nd = _f32c(z).to(dev), _f32c(cond).to(dev), _f32c(noise).to(dev)
            md = _f32c(mask).to(dev).reshape(B, T)
            sd = _f32c(spk_emb).to(dev).reshape(B, self.spk_emb_dim)
            out = torch.empty_like(zd)
            tr = torch.empty_like(nd) if trace else None
            abi.check(lib.usb_reverse_diffusion(h, zd.data_ptr(), cd.data_ptr(), md.data_ptr(), sd.data_ptr(),
                                                nd.data_ptr(), coef.data_ptr(), times.data_ptr(), n_timesteps, tg, sg,
                                                out.data_ptr(), tr.data_ptr() if trace else None, B, T, stream))
        if on_host:
            out = out.cpu()
            tr = tr.cpu() if trace else None
        return (out, tr) if trace else out

    def _set_denorm(self, h, denorm) -> None:
        key = None
        if denorm is not None:
            lo, hi = (t.detach().to(torch.float32).reshape(-1).cpu().contiguous() for t in denorm)
            if lo.numel() != self.n_feats or hi.numel() != self.n_feats:
                raise ValueError(f"denorm=(mel_min, mel_max) must hold {self.n_feats} values each")
            key = (lo.numpy().tobytes(), hi.numpy().tobytes())
        if key == getattr(self, "_denorm_key", None) and getattr(self, "_denorm_handle", None) == h.value:
            return
        lib = abi.load_library()
        if denorm is None:
            abi.check(lib.usb_set_output_denorm(h, None, None))
        else:
            abi.check(lib.usb_set_output_denorm(h, lo.data_ptr(), hi.data_ptr()))
        self._denorm_key, self._denorm_handle = key, h.value

    @torch.no_grad()
    def forward(self, z, mask, cond, spk_emb, n_timesteps, text_gradient_scale=0.0, spk_gradient_scale=0.0, **kw):
        """unitspeech/unitspeech.py:387-391."""
        return self.reverse_diffusion(z, mask, cond, spk_emb, n_timesteps, text_gradient_scale=text_gradient_scale,
                                      spk_gradient_scale=spk_gradient_scale, **kw)

    @torch.no_grad()
    def execute_text_to_speech(self, phoneme, phoneme_lengths, spk_emb, text_encoder, duration_predictor,
                               num_downsamplings_in_unet, diffusion_steps=50, length_scale=1.0,
                               text_gradient_scale=1.0, spk_gradient_scale=1.0, noise: Optional[torch.Tensor] = None,
                               max_frames: Optional[int] = None, denorm: Optional[Sequence[torch.Tensor]] = None):
        """unitspeech/unitspeech.py:414-450: encoder -> durations -> alignment -> z -> reverse diffusion -> crop.

        ``max_frames=None`` follows the reference exactly, including its host round trip ``int(y_lengths.max())`` (:428).
        With ``max_frames`` (rounded up by fix_len_compatibility) the glue between the duration predictor and the sampler
        runs in one CUDA kernel (usb_align_expand) at that fixed frame capacity and nothing synchronises with the host:
        the three outputs come back padded to the capacity (frames past an utterance's length are zero) and the
        per-utterance frame counts are left in ``self.last_y_lengths`` (device int64, capped at the capacity; utterances
        that did not fit are flagged in ``self.last_y_overflow``, device bool)."""
        cond_x, x, x_mask = text_encoder(phoneme, phoneme_lengths)
        logw = duration_predictor(x, x_mask, w=None, g=spk_emb, reverse=True)
        w = torch.exp(logw) * x_mask
        w_ceil = torch.ceil(w) * length_scale
        if max_frames is not None:
            return self._tts_on_device(cond_x, x_mask, w_ceil, spk_emb, num_downsamplings_in_unet, int(max_frames),
                                       diffusion_steps, text_gradient_scale, spk_gradient_scale, noise, denorm)
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        y_max_length = int(y_lengths.max())
        y_max_length_ = fix_len_compatibility(y_max_length, num_downsamplings_in_unet)
        y_mask = sequence_mask(y_lengths, y_max_length_).unsqueeze(1).to(x_mask.dtype)
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        attn = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1)).unsqueeze(1)
        cond_y = torch.matmul(attn.squeeze(1).transpose(1, 2).contiguous(), cond_x.transpose(1, 2).contiguous())
        cond_y = cond_y.transpose(1, 2).contiguous()
        encoder_outputs = cond_y[:, :, :y_max_length]
        z = torch.randn_like(cond_y, device=cond_y.device)
        decoder_outputs = self.forward(z, y_mask, cond_y, spk_emb, n_timesteps=diffusion_steps,
                                       text_gradient_scale=text_gradient_scale, spk_gradient_scale=spk_gradient_scale,
                                       noise=noise, denorm=denorm)
        decoder_outputs = decoder_outputs[:, :, :y_max_length]
        return encoder_outputs, decoder_outputs, attn[:, :, :y_max_length]

    def _tts_on_device(self, cond_x, x_mask, w_ceil, spk_emb, n_down, max_frames, diffusion_steps, tg, sg, noise,
                       denorm=None):
        lib = abi.load_library()
        h = self._ensure_handle(cond_x)
        dev = torch.device("cuda", self._handle_device)
        B, F, Tx = cond_x.shape
        T = fix_len_compatibility(max_frames, n_down)
        wd, xm, cx = _f32c(w_ceil).to(dev).reshape(B, Tx), _f32c(x_mask).to(dev).reshape(B, Tx), _f32c(cond_x).to(dev)
        y_lengths = torch.empty(B, dtype=torch.int64, device=dev)
        y_mask = torch.empty(B, 1, T, dtype=torch.float32, device=dev)
        attn = torch.empty(B, 1, Tx, T, dtype=torch.float32, device=dev)
        cond_y = torch.empty(B, F, T, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            abi.check(lib.usb_align_expand(h, wd.data_ptr(), xm.data_ptr(), cx.data_ptr(), B, Tx, F, T, y_lengths.data_ptr(),
                                           y_mask.data_ptr(), attn.data_ptr(), cond_y.data_ptr(), self._stream(dev.index)))
        self.last_y_lengths = y_lengths
        # utterances whose predicted durations exceed the frame capacity are truncated at T (device bool, no host sync)
        self.last_y_overflow = wd.sum(dim=1) > T
        z = torch.randn_like(cond_y)
        dec = self.forward(z, y_mask, cond_y, spk_emb.to(dev), n_timesteps=diffusion_steps, text_gradient_scale=tg,
                           spk_gradient_scale=sg, noise=noise, denorm=denorm)
        return cond_y, dec, attn

    # ------------------------------------------------------------------ training objective
    def _dev_tensors(self, like, *tensors):
        self._ensure_handle(like)
        dev = torch.device("cuda", self._handle_device)
        return dev, [_f32c(t).to(dev) for t in tensors]

    @torch.no_grad()
    def forward_diffusion(self, x0, mask, t):
        """UnitSpeech.forward_diffusion (unitspeech/unitspeech.py:376-384): returns (xt * mask, z * mask); z is drawn
        with torch.randn(x0.shape) from x0's device generator exactly like the reference (:381)."""
        lib = abi.load_library()
        B, F, T = x0.shape
        z = torch.randn(x0.shape, dtype=x0.dtype, device=x0.device, requires_grad=False)
        dev, (xd, md, td, zd) = self._dev_tensors(x0, x0, mask.reshape(B, T), t.reshape(B), z)
        xt, zm = torch.empty_like(xd), torch.empty_like(xd)
        with torch.cuda.device(dev):
            abi.check(lib.usb_forward_diffusion(self._handle, xd.data_ptr(), md.data_ptr(), td.data_ptr(), zd.data_ptr(),
                                                xt.data_ptr(), zm.data_ptr(), B, T, self._stream(dev.index)))
        return xt.to(x0.device), zm.to(x0.device)

    def _wants_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def _ensure_tuner(self, like: Optional[torch.Tensor] = None):
        """The fine-tune engine (unitspeech_b200/training.py) behind loss.backward(): created on first use, its fp32
        master copy refreshed from the module parameters before every forward (the caller's optimizer updates the
        module parameters in place)."""
        from .training import FineTuner
        dev = self._device_index(like)
        if self._tuner is None or self._tuner.dev.index != dev:
            self._tuner = FineTuner(n_feats=self.n_feats, dim=self.dim, dim_mults=self.dim_mults, beta_min=self.beta_min,
                                    beta_max=self.beta_max, pe_scale=self.pe_scale, spk_emb_dim=self.spk_emb_dim, device=dev)
        # the optimizer updates parameters in place (tensor._version counts that); an unchanged set needs no re-upload
        stamp = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if stamp != getattr(self._tuner, "_source_stamp", None):
            self._tuner.load_state_dict(dict(self.state_dict()), strict=True)
            self._tuner._source_stamp = stamp
        return self._tuner

    def fused_finetuner(self, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, loss_scale=8192.0):
        """Fast path of the fine-tune loop (finetune.py:131-165): returns a FineTuner holding a copy of this decoder's
        weights whose ``fine_tune(...)`` runs zero_grad + loss + backward + clip_grad_norm_ + Adam entirely in the CUDA
        library (two launches for clip + Adam over all parameters).  Copy the result back with
        ``decoder.load_state_dict(tuner.state_dict())``."""
        from .training import FineTuner
        ft = FineTuner(n_feats=self.n_feats, dim=self.dim, dim_mults=self.dim_mults, beta_min=self.beta_min,
                       beta_max=self.beta_max, pe_scale=self.pe_scale, spk_emb_dim=self.spk_emb_dim, lr=lr, betas=betas, eps=eps,
                       max_norm=max_norm, loss_scale=loss_scale, device=self._device_index())
        ft.load_state_dict(dict(self.state_dict()), strict=True)
        return ft

    def loss_t(self, x0, mask, cond, t, spk_emb):
        """UnitSpeech.loss_t (unitspeech/unitspeech.py:393-405) -> (loss, xt).  With autograd enabled and trainable
        parameters the returned loss back-propagates through the CUDA backward pass into ``parameter.grad`` (so the
        reference loop ``loss.backward(); clip_grad_norm_(...); optimizer.step()`` works unchanged, finetune.py:163-165);
        under ``torch.no_grad()`` it is the forward value only."""
        B, F, T = x0.shape
        if T % (2 ** (len(self.dim_mults) - 1)):
            raise ValueError("T must be a multiple of 2**(len(dim_mults)-1) (use fix_len_compatibility)")
        z = torch.randn(x0.shape, dtype=x0.dtype, device=x0.device, requires_grad=False)
        if torch.is_grad_enabled() and any(isinstance(v, torch.Tensor) and v.requires_grad for v in (x0, cond, spk_emb)):
            # the reference back-propagates through the estimator into cond / x0 / spk_emb (train_STEP1.py:381,
            # train_STEP2.py:299 train an encoder that way); the CUDA backward pass only produces the decoder's parameter
            # gradients, so refuse instead of silently training the caller's encoder on a constant
            raise NotImplementedError(
                "unitspeech_b200.UnitSpeech.loss_t computes gradients for the decoder parameters only (the fine-tuning "
                "path, finetune.py:131-165); x0 / cond / spk_emb require grad here, which needs the reference decoder "
                "(detach them or wrap the call in torch.no_grad())")
        if self._wants_grad():
            ft = self._ensure_tuner(x0)
            names = [k for k, _ in self.named_parameters()]
            loss = _DiffusionLoss.apply(ft, names, x0, mask, cond, t, spk_emb, z, *[p for _, p in self.named_parameters()])
            # an optimizer is about to update the parameters in place: the sampler's copy of the weights (fp16 operands
            # inside the library handle) must be rebuilt before the next inference call
            self._weights_version += 1
            return loss.to(x0.device), ft.xt.detach().clone().to(x0.device)
        with torch.no_grad():
            lib = abi.load_library()
            dev, (xd, md, cd, td, sd, zd) = self._dev_tensors(x0, x0, mask.reshape(B, T), cond, t.reshape(B),
                                                             spk_emb.reshape(B, self.spk_emb_dim), z)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            xt = torch.empty_like(xd)
            with torch.cuda.device(dev):
                abi.check(lib.usb_loss_t(self._handle, xd.data_ptr(), cd.data_ptr(), md.data_ptr(), td.data_ptr(),
                                         sd.data_ptr(), zd.data_ptr(), loss.data_ptr(), xt.data_ptr(), B, T,
                                         self._stream(dev.index)))
            return loss.to(x0.device), xt.to(x0.device)

    def compute_loss(self, x0, mask, cond, spk_emb=None, offset=1e-5):
        """UnitSpeech.compute_loss (unitspeech/unitspeech.py:407-411): t ~ U(offset, 1 - offset) per utterance."""
        t = torch.rand(x0.shape[0], dtype=x0.dtype, device=x0.device, requires_grad=False)
        t = torch.clamp(t, offset, 1.0 - offset)
        return self.loss_t(x0, mask, cond, t, spk_emb)

    def fine_tune(self, cond_x, y, y_mask, y_lengths, y_max_length, attn, spk_emb, segment_size, n_feats):
        """UnitSpeech.fine_tune (unitspeech/unitspeech.py:452-492): random segment crop (Python `random`, as the
        reference), alignment of the encoder output to the crop, then the diffusion loss (see loss_t for gradients)."""
        y_cut, y_cut_mask, cond_y = crop_segments(cond_x, y, y_mask, y_lengths, y_max_length, attn, segment_size, n_feats)
        diff_loss, _ = self.compute_loss(y_cut, y_cut_mask, cond_y, spk_emb=spk_emb)
        return diff_loss


class _DiffusionLoss(torch.autograd.Function):
    """loss_t with the CUDA backward pass behind torch.autograd: the parameters are inputs of the node, so
    ``loss.backward()`` accumulates into ``parameter.grad`` like the reference's eager graph."""

    @staticmethod
    def forward(ctx, ft, names, x0, mask, cond, t, spk_emb, z, *params):
        loss = ft.forward(x0, mask, cond, t, spk_emb, z)
        ctx.ft, ctx.names, ctx.devs = ft, names, [p.device for p in params]
        ctx.generation = ft.generation
        return loss.detach().clone().reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        ft = ctx.ft
        if ft.generation != ctx.generation:
            # the engine keeps ONE set of activations: a second loss_t before this backward has overwritten them
            raise RuntimeError("unitspeech_b200: the activations of this loss were overwritten by a later loss_t / fine_tune "
                               "call; call backward() before evaluating the next loss (gradient accumulation over several "
                               "losses is not supported by the CUDA fine-tune engine)")
        ft.zero_grad()
        ft.backward()
        scale = grad_out.to(ft.dev, torch.float32) / ft.loss_scale
        grads = tuple((ft.grad_reference(k) * scale).to(d) for k, d in zip(ctx.names, ctx.devs))
        return (None,) * 8 + grads


def crop_segments(cond_x, y, y_mask, y_lengths, y_max_length, attn, segment_size, n_feats):
    """The segment crop + alignment half of UnitSpeech.fine_tune (unitspeech/unitspeech.py:453-487)."""
    import random
    if y_max_length < segment_size:
        pad_size = segment_size - y_max_length
        y = torch.cat([y, torch.zeros_like(y)[:, :, :pad_size]], dim=-1)
        y_mask = torch.cat([y_mask, torch.zeros_like(y_mask)[:, :, :pad_size]], dim=-1)
    max_offset = (y_lengths - segment_size).clamp(0)
    out_offset = [random.choice(range(0, int(end))) if int(end) > 0 else 0 for end in max_offset.cpu().numpy()]
    attn_cut = torch.zeros(attn.shape[0], attn.shape[1], segment_size, dtype=attn.dtype, device=attn.device)
    y_cut = torch.zeros(y.shape[0], n_feats, segment_size, dtype=y.dtype, device=y.device)
    y_cut_lengths = []
    for i, lower in enumerate(out_offset):
        cut_len = segment_size + int((y_lengths[i] - segment_size).clamp(None, 0))
        y_cut_lengths.append(cut_len)
        y_cut[i, :, :cut_len] = y[i, :, lower:lower + cut_len]
        attn_cut[i, :, :cut_len] = attn[i, :, lower:lower + cut_len]
    y_cut_mask = sequence_mask(torch.LongTensor(y_cut_lengths)).unsqueeze(1).to(y_mask)
    if y_cut_mask.shape[-1] < segment_size:
        y_cut_mask = torch.nn.functional.pad(y_cut_mask, (0, segment_size - y_cut_mask.shape[-1]))
    cond_y = torch.matmul(attn_cut.squeeze(1).transpose(1, 2).contiguous(), cond_x.transpose(1, 2).contiguous())
    cond_y = cond_y.transpose(1, 2).contiguous() * y_cut_mask
    return y_cut, y_cut_mask, cond_y


def denormalize_mel(y: torch.Tensor, mel_min: torch.Tensor, mel_max: torch.Tensor) -> torch.Tensor:
    """The mel normalisation contract of the callers (inference.py:140): [-1, 1] -> log-mel."""
    return (y + 1) / 2 * (mel_max - mel_min) + mel_min


def normalize_mel(mel: torch.Tensor, mel_min: torch.Tensor, mel_max: torch.Tensor) -> torch.Tensor:
    """Inverse of denormalize_mel (data.py:91, finetune.py:104)."""
    return (mel - mel_min) / (mel_max - mel_min) * 2 - 1
