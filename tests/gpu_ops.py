"""Test-side wrappers around the operator-level C-ABI entries (usb_op_*)."""

import ctypes

import torch

from unitspeech_b200 import abi


class OpHandle:
    def __init__(self, device=0):
        self.lib = abi.load_library()
        cfg = abi.UsbConfig()
        cfg.n_feats, cfg.dim, cfg.n_mults = 80, 64, 2
        cfg.dim_mults[0], cfg.dim_mults[1] = 1, 2
        cfg.groups, cfg.spk_emb_dim, cfg.pe_scale, cfg.beta_min, cfg.beta_max, cfg.device = 8, 256, 1000.0, 0.05, 20.0, device
        self.h = ctypes.c_void_p()
        abi.check(self.lib.usb_create(ctypes.byref(cfg), ctypes.byref(self.h)))
        self.dev = torch.device("cuda", device)

    def close(self):
        if self.h:
            self.lib.usb_destroy(self.h)
            self.h = None

    def _stream(self):
        return int(torch.cuda.current_stream(self.dev).cuda_stream)

    def conv(self, kind, x0, weight, bias=None, x1=None, mask=None, residual=None, res_scale=1.0, stats_groups=0):
        """x0/x1: NCHW fp32 torch tensors (rounded to fp16 inside); weight/bias: reference layout fp32 (CPU ok).
        Returns (out NCHW fp32, stats or None)."""
        N, C0, H, W = x0.shape
        C1 = x1.shape[1] if x1 is not None else 0
        Cout = weight.shape[1] if kind == 3 else weight.shape[0]
        Hout, Wout = (H // 2, W // 2) if kind == 1 else ((2 * H, 2 * W) if kind == 3 else (H, W))
        to_nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(self.dev, torch.float16)  # noqa: E731
        a0 = to_nhwc(x0)
        a1 = to_nhwc(x1) if x1 is not None else None
        out = torch.full((N, Hout, Wout, Cout), float("nan"), dtype=torch.float16, device=self.dev)
        w = weight.detach().float().contiguous().cpu()
        b = bias.detach().float().contiguous().cpu() if bias is not None else None
        m = mask.detach().float().contiguous().to(self.dev) if mask is not None else None
        r = to_nhwc(residual) if residual is not None else None
        st = torch.zeros(N, stats_groups, 2, dtype=torch.int64, device=self.dev) if stats_groups else None
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        abi.check(self.lib.usb_op_conv(self.h, kind, ptr(a0), ptr(a1), N, H, W, C0, C1, Cout, ptr(w), ptr(b), ptr(m),
                                       ptr(r), float(res_scale), ptr(st), stats_groups, ptr(out), self._stream()))
        torch.cuda.synchronize(self.dev)
        if st is not None:  # fixed point -> (sum, sumsq)
            st = st.double() / torch.tensor([2.0 ** 24, 2.0 ** 18], dtype=torch.float64, device=self.dev)
        return out.float().permute(0, 3, 1, 2).contiguous(), st

    def gn_apply(self, raw, stats, gamma, beta, addvec, res, mask, groups=8):
        N, C, H, W = raw.shape
        to_nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(self.dev, torch.float16)  # noqa: E731
        f = lambda t: t.detach().float().contiguous().to(self.dev) if t is not None else None  # noqa: E731
        a = to_nhwc(raw)
        r = to_nhwc(res) if res is not None else None
        out = torch.empty_like(a)
        st = (stats.double() * torch.tensor([2.0 ** 24, 2.0 ** 18], dtype=torch.float64)).round().to(torch.int64)
        st = st.to(self.dev).contiguous()
        g, b, av, m = f(gamma), f(beta), f(addvec), f(mask)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        abi.check(self.lib.usb_op_gn_apply(self.h, ptr(a), ptr(st), ptr(g), ptr(b), ptr(av), ptr(r), ptr(m), ptr(out),
                                           N, H, W, C, groups, self._stream()))
        torch.cuda.synchronize(self.dev)
        return out.float().permute(0, 3, 1, 2).contiguous()

    def attn_context(self, qkv, wo, heads=4):
        """qkv: (N, 384, H, W) fp32; wo: (C, 128) -> weff (N, C, 128) fp32"""
        N, C3, H, W = qkv.shape
        a = qkv.permute(0, 2, 3, 1).contiguous().to(self.dev, torch.float16).reshape(N, H * W, C3)
        w = wo.detach().float().contiguous().to(self.dev)
        C = w.shape[0]
        out = torch.empty(N, C, C3 // 3, dtype=torch.float16, device=self.dev)
        abi.check(self.lib.usb_op_attn_context(self.h, ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(w.data_ptr()),
                                               ctypes.c_void_p(out.data_ptr()), N, H * W, C, heads, self._stream()))
        torch.cuda.synchronize(self.dev)
        return out.float()
