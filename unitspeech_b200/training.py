"""Speaker-adaptation (fine-tune) step of the UnitSpeech decoder on the CUDA path.

Reference semantics: ``UnitSpeech.fine_tune`` -> ``compute_loss`` -> ``loss_t`` (unitspeech/unitspeech.py:393-411,452-492)
followed by ``loss.backward()``, ``clip_grad_norm_(decoder.parameters(), max_norm=1)`` and ``torch.optim.Adam(lr=2e-5)``
(finetune.py:81,131-165).  There is no autograd here: this module walks the estimator's forward graph
(GradLogPEstimator2d.forward, unitspeech/unitspeech.py:164-201) through the operator-level C ABI
(include/unitspeech_b200_train.h), keeps the activations the backward pass needs, and then walks the graph in reverse.
PyTorch only allocates the buffers and hands out pointers; every FLOP runs in the kernels of libunitspeech_b200.so
(no CPU or PyTorch fallback: the constructor raises without the library or without a B200).

Numerics: activations and activation gradients are NHWC fp16 with fp32 accumulation, parameters / gradients / Adam state
fp32.  The objective is multiplied by ``loss_scale`` before differentiation (fp16 gradient range) and divided out inside
the optimizer kernel; non-finite gradients skip the step (like torch.cuda.amp.GradScaler, finetune.py:156-162).
"""

from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import abi

K3S1, K3S2, K1, KT4, K3S2D, KT4D = 0, 1, 2, 3, 4, 5
_HID, _HEADS = 128, 4


_KH = torch.tensor([[1, 3], [0, 2]])   # ConvTranspose 4x4/s2: kernel row used by output phase ph, tap a (engine.cu pack_conv_host)


def to_train_layout(kind: int, w: torch.Tensor) -> torch.Tensor:
    """Reference conv weight -> training layout (= forward GEMM operand layout, input channels contiguous):
    3x3 (Cout, Cin, 3, 3) -> (Cout, 3, 3, Cin); ConvTranspose (Cin, Cout, 4, 4) -> (ph, pw, Cout, a, b, Cin); 1x1 unchanged."""
    if kind in (K3S1, K3S2):
        return w.permute(0, 2, 3, 1).contiguous()
    if kind == KT4:
        kh = _KH.to(w.device)
        x = w[:, :, kh[:, None, :, None], kh[None, :, None, :]]          # (Cin, Cout, ph, pw, a, b)
        return x.permute(2, 3, 1, 4, 5, 0).contiguous()
    return w


def to_reference_layout(kind: int, t: torch.Tensor, shape) -> torch.Tensor:
    """Inverse of to_train_layout; ``shape`` is the reference shape."""
    if kind in (K3S1, K3S2):
        co, ci = shape[0], shape[1]
        return t.reshape(co, 3, 3, ci).permute(0, 3, 1, 2).contiguous()
    if kind == KT4:
        ci, co = shape[0], shape[1]
        kh = _KH.to(t.device)
        x = t.reshape(2, 2, co, 2, 2, ci).permute(5, 2, 0, 1, 3, 4)      # (Cin, Cout, ph, pw, a, b)
        out = torch.empty(ci, co, 4, 4, dtype=t.dtype, device=t.device)
        out[:, :, kh[:, None, :, None], kh[None, :, None, :]] = x
        return out
    return t.reshape(shape).clone()


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def param_shapes(n_feats: int, dim: int, dim_mults: Sequence[int], spk_emb_dim: int) -> Dict[str, Tuple[int, ...]]:
    """state_dict entries of UnitSpeech (unitspeech/unitspeech.py:125-162,221-233), in a flat-buffer order that keeps the
    16 ResnetBlock.mlp Linears contiguous (they are evaluated as one stacked Linear)."""
    e = "estimator."
    dims = [2] + [dim * m for m in dim_mults]
    L = len(dim_mults)
    temb = dim + spk_emb_dim
    s: Dict[str, Tuple[int, ...]] = {}
    resnets: List[Tuple[str, int, int]] = []
    attns: List[Tuple[str, int]] = []
    for k in range(L):
        cin, cout = dims[k], dims[k + 1]
        resnets += [(f"{e}downs.{k}.0", cin, cout), (f"{e}downs.{k}.1", cout, cout)]
        attns.append((f"{e}downs.{k}.2", cout))
        if k < L - 1:
            s[f"{e}downs.{k}.3.conv.weight"] = (cout, cout, 3, 3)
            s[f"{e}downs.{k}.3.conv.bias"] = (cout,)
    mid = dims[-1]
    resnets.append((e + "mid_block1", mid, mid))
    attns.append((e + "mid_attn", mid))
    resnets.append((e + "mid_block2", mid, mid))
    for k in range(L - 1):
        j = L - 1 - k
        cj, cn = dims[j + 1], dims[j]
        resnets += [(f"{e}ups.{k}.0", 2 * cj, cn), (f"{e}ups.{k}.1", cn, cn)]
        attns.append((f"{e}ups.{k}.2", cn))
        s[f"{e}ups.{k}.3.conv.weight"] = (cn, cn, 4, 4)
        s[f"{e}ups.{k}.3.conv.bias"] = (cn,)
    for pre, _, cout in resnets:          # contiguous: the stacked Linear (J, dim + spk_emb_dim)
        s[pre + ".mlp.1.weight"] = (cout, temb)
    for pre, _, cout in resnets:
        s[pre + ".mlp.1.bias"] = (cout,)
    for pre, cin, cout in resnets:
        s[pre + ".block1.block.0.weight"] = (cout, cin, 3, 3)
        s[pre + ".block1.block.0.bias"] = (cout,)
        s[pre + ".block1.block.1.weight"] = (cout,)
        s[pre + ".block1.block.1.bias"] = (cout,)
        s[pre + ".block2.block.0.weight"] = (cout, cout, 3, 3)
        s[pre + ".block2.block.0.bias"] = (cout,)
        s[pre + ".block2.block.1.weight"] = (cout,)
        s[pre + ".block2.block.1.bias"] = (cout,)
        if cin != cout:
            s[pre + ".res_conv.weight"] = (cout, cin, 1, 1)
            s[pre + ".res_conv.bias"] = (cout,)
    for pre, c in attns:
        s[pre + ".fn.g"] = (1,)
        s[pre + ".fn.fn.to_qkv.weight"] = (3 * _HID, c, 1, 1)
        s[pre + ".fn.fn.to_out.weight"] = (c, _HID, 1, 1)
        s[pre + ".fn.fn.to_out.bias"] = (c,)
    s[e + "mlp.0.weight"] = (4 * dim, dim)
    s[e + "mlp.0.bias"] = (4 * dim,)
    s[e + "mlp.2.weight"] = (dim, 4 * dim)
    s[e + "mlp.2.bias"] = (dim,)
    s[e + "final_block.block.0.weight"] = (dim, dim, 3, 3)
    s[e + "final_block.block.0.bias"] = (dim,)
    s[e + "final_block.block.1.weight"] = (dim,)
    s[e + "final_block.block.1.bias"] = (dim,)
    s[e + "final_conv.weight"] = (1, dim, 1, 1)
    s[e + "final_conv.bias"] = (1,)
    s["text_uncon"] = (1, n_feats, 1)
    s["spk_uncon"] = (1, 1, spk_emb_dim)
    return s


class _Conv:
    """fp32 master weight view + its fp16 GEMM operands (refreshed after every optimizer step)."""

    def __init__(self, ft: "FineTuner", key: str, kind: int, cout: int, cin: int, has_bias: bool = True,
                 splits: Optional[Sequence[Tuple[int, int]]] = None, need_dgrad: bool = True):
        self.key, self.kind, self.cout, self.cin = key, kind, cout, cin
        ft.layout_kind[key + ".weight"] = kind        # stored in the training layout (to_train_layout)
        self.w, self.dw = ft.params[key + ".weight"], ft.grads[key + ".weight"]
        self.b = ft.params[key + ".bias"] if has_bias else None
        self.db = ft.grads[key + ".bias"] if has_bias else None
        taps_f = {K3S1: 9, K3S2: 9, K1: 1, KT4: 16}[kind]
        taps_d = {K3S1: 9, K3S2: 16, K1: 1, KT4: 16}[kind]
        # forward operand = the fp16 mirror of the master weight (same offset in ft.P16: one cast refreshes all of them)
        off = (self.w.data_ptr() - ft.P.data_ptr()) // 4
        self.fwd = ft.P16[off:off + cout * cin * taps_f]
        self.splits = list(splits) if splits else [(0, cin)]
        self.dgrad = [torch.empty(cout * (c1 - c0) * taps_d, dtype=torch.float16, device=ft.dev) for c0, c1 in self.splits] \
            if need_dgrad else []

    def pack(self, ft: "FineTuner"):
        for (c0, c1), d in zip(self.splits, self.dgrad):
            ft.call("usb_t_pack_conv", self.kind, _p(self.w), self.cout, self.cin, c0, c1, None, _p(d))


class FineTuner:
    """Owns the fp32 master parameters, their gradients and the Adam state of one decoder on one B200, and runs
    forward + backward + clip + Adam steps of the diffusion objective."""

    def __init__(self, n_feats=80, dim=128, dim_mults=(1, 2, 4, 8), beta_min=0.05, beta_max=20.0, pe_scale=1000,
                 spk_emb_dim=256, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, loss_scale=8192.0, device=0,
                 use_cuda_graph: bool = True, _trace_calls: Optional[list] = None):
        self.lib = abi.load_library()
        # _trace_calls (tests only): record the ABI call sequence on CPU buffers instead of launching anything --
        # it checks the host-side graph walk and computes nothing
        self._trace = _trace_calls
        if self._trace is None and not torch.cuda.is_available():
            raise abi.UsbError("unitspeech_b200 fine-tuning needs a B200 (there is no CPU or PyTorch fallback)")
        if self._trace is not None:
            self.dev = torch.device("cpu")
        else:
            self.dev = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.n_feats, self.dim, self.dim_mults, self.S = n_feats, dim, tuple(dim_mults), spk_emb_dim
        self.L = len(self.dim_mults)
        self.C = [dim * m for m in self.dim_mults]
        self.lr, self.betas, self.eps, self.max_norm, self.loss_scale = lr, betas, eps, max_norm, float(loss_scale)
        self.use_cuda_graph = bool(use_cuda_graph) and _trace_calls is None
        self.overlap_wgrad = True          # weight-gradient kernels on a side stream, overlapping the data-gradient chain
        self._readers: Dict[int, "torch.cuda.Event"] = {}
        self._graphs: Dict[Tuple[int, int], "torch.cuda.CUDAGraph"] = {}
        self._eager_steps: Dict[Tuple[int, int], int] = {}
        cfg = abi.UsbConfig()
        cfg.n_feats, cfg.dim, cfg.n_mults = n_feats, dim, self.L
        for i, m in enumerate(self.dim_mults):
            cfg.dim_mults[i] = m
        cfg.groups, cfg.spk_emb_dim, cfg.pe_scale = 8, spk_emb_dim, float(pe_scale)
        cfg.beta_min, cfg.beta_max, cfg.device = float(beta_min), float(beta_max), self.dev.index or 0
        self.h = ctypes.c_void_p()
        if self._trace is None:
            abi.check(self.lib.usb_create(ctypes.byref(cfg), ctypes.byref(self.h)))
        # ---- flat fp32 parameter / gradient / Adam buffers with named views
        self.shapes = param_shapes(n_feats, dim, self.dim_mults, spk_emb_dim)
        # every parameter starts on a 256-byte boundary (the conv epilogue reads biases as float4); the padding stays 0
        al = lambda n: (n + 63) // 64 * 64  # noqa: E731
        total = sum(al(int(torch.Size(s).numel())) for s in self.shapes.values())
        self.nparams = total
        mk = lambda: torch.zeros(total, dtype=torch.float32, device=self.dev)  # noqa: E731
        self.P, self.G, self.M, self.V = mk(), mk(), mk(), mk()
        self.P16 = torch.zeros(total, dtype=torch.float16, device=self.dev)    # fp16 mirror of P (conv forward operands)
        self.params: Dict[str, torch.Tensor] = {}     # flat views; conv weights of layout_kind are in the training layout
        self.grads: Dict[str, torch.Tensor] = {}
        self.layout_kind: Dict[str, int] = {}
        off = 0
        for k, shp in self.shapes.items():
            n = int(torch.Size(shp).numel())
            self.params[k] = self.P[off:off + n].view(shp)
            self.grads[k] = self.G[off:off + n].view(shp)
            off += al(n)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.skipped = torch.zeros(1, dtype=torch.int32, device=self.dev)   # Adam steps skipped for a non-finite gradient norm
        self.generation = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)   # optimizer steps taken (device counter)
        half = dim // 2
        import math
        e = math.log(10000) / (half - 1)          # SinusoidalPosEmb table, the reference's own torch expression (:116-118)
        self.freqs = torch.exp(torch.arange(half).float() * -e).to(self.dev)
        self._build_modules()
        self.side = torch.cuda.Stream(self.dev) if self._trace is None else None
        self._ws: Dict[str, torch.Tensor] = {}
        self._ws_key = None
        self._packed = False

    # ------------------------------------------------------------------------------------------------ plumbing
    _BAKED = ("lr", "betas", "eps", "max_norm", "loss_scale", "overlap_wgrad")

    def __setattr__(self, name, value):
        # hyper-parameters are kernel arguments baked into a captured iteration: changing one drops the captured graphs
        if name in FineTuner._BAKED and getattr(self, "_graphs", None):
            self._graphs.clear()
            self._eager_steps.clear()
        object.__setattr__(self, name, value)

    def close(self):
        if getattr(self, "h", None) and self._trace is None:
            torch.cuda.synchronize(self.dev)
            n_skipped = int(self.skipped.item())
            if n_skipped:
                import warnings
                warnings.warn(f"FineTuner: {n_skipped} optimizer step(s) were skipped for a non-finite gradient norm "
                              f"(loss_scale={self.loss_scale}); lower the loss scale", RuntimeWarning)
            self.lib.usb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name: str, *args):
        if self._trace is not None:
            assert len(args) + 2 == len(abi.SIGNATURES[name][1]), f"{name}: {len(args) + 2} arguments"
            self._trace.append((name, args))
            return
        stream = int(torch.cuda.current_stream(self.dev).cuda_stream)
        abi.check(getattr(self.lib, name)(self.h, *args, stream))

    def load_state_dict(self, state: Dict[str, torch.Tensor], strict: bool = True):
        """Reference key names and shapes (inference.py:66-73 loads ``ckpt["model"]``)."""
        missing = [k for k in self.shapes if k not in state]
        extra = [k for k in state if k not in self.shapes]
        if strict and (missing or extra):
            raise KeyError(f"state_dict mismatch: missing {missing[:4]}, unexpected {extra[:4]}")
        for k, shp in self.shapes.items():
            if k in state:
                v = state[k]
                if tuple(v.shape) != tuple(shp):
                    raise ValueError(f"{k}: shape {tuple(v.shape)} != {tuple(shp)}")
                v = v.detach().to(self.dev, torch.float32)
                if k in self.layout_kind:
                    v = to_train_layout(self.layout_kind[k], v)
                self.params[k].view(-1).copy_(v.reshape(-1))
        self._packed = False          # (captured graphs stay valid: they re-pack at the end of every step)

    def _reference(self, k: str, flat: torch.Tensor) -> torch.Tensor:
        return to_reference_layout(self.layout_kind[k], flat, self.shapes[k]) if k in self.layout_kind else flat.detach().clone()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Reference key names, shapes and layouts."""
        return {k: self._reference(k, v) for k, v in self.params.items()}

    def grad_reference(self, k: str) -> torch.Tensor:
        """Gradient of parameter ``k`` in the reference layout, still multiplied by the loss scale."""
        return self._reference(k, self.grads[k])

    def zero_grad(self):
        self.G.zero_()

    def _build_modules(self):
        e = "estimator."
        C, L = self.C, self.L
        self.resnets: List[dict] = []
        self.attns: List[dict] = []
        self.convs: List[_Conv] = []

        def conv(key, kind, cout, cin, has_bias=True, splits=None, need_dgrad=True):
            c = _Conv(self, key, kind, cout, cin, has_bias, splits, need_dgrad)
            self.convs.append(c)
            return c

        emb_off = 0

        def resnet(pre, cin, cout, first=False, split=False):
            nonlocal emb_off
            r = {"pre": pre, "cin": cin, "cout": cout, "first": first, "emb_off": emb_off, "has_res": cin != cout}
            emb_off += cout
            sp = [(0, cin // 2), (cin // 2, cin)] if split else None
            if not first:
                r["c1"] = conv(pre + ".block1.block.0", K3S1, cout, cin, splits=sp)
                if r["has_res"]:
                    r["res"] = conv(pre + ".res_conv", K1, cout, cin, splits=sp)
            r["c2"] = conv(pre + ".block2.block.0", K3S1, cout, cout)
            self.resnets.append(r)
            return r

        def attn(pre, c):
            a = {"pre": pre, "C": c, "qkv": conv(pre + ".fn.fn.to_qkv", K1, 3 * _HID, c, has_bias=False)}
            self.attns.append(a)
            return a

        self.down, self.up = [], []
        for k in range(L):
            cin, cout = (2 if k == 0 else C[k - 1]), C[k]
            resnet(f"{e}downs.{k}.0", cin, cout, first=(k == 0))
            resnet(f"{e}downs.{k}.1", cout, cout)
            attn(f"{e}downs.{k}.2", cout)
            if k < L - 1:
                self.down.append(conv(f"{e}downs.{k}.3.conv", K3S2, cout, cout))
        resnet(e + "mid_block1", C[-1], C[-1])
        attn(e + "mid_attn", C[-1])
        resnet(e + "mid_block2", C[-1], C[-1])
        for k in range(L - 1):
            j = L - 1 - k
            resnet(f"{e}ups.{k}.0", 2 * C[j], C[j - 1], split=True)
            resnet(f"{e}ups.{k}.1", C[j - 1], C[j - 1])
            attn(f"{e}ups.{k}.2", C[j - 1])
            self.up.append(conv(f"{e}ups.{k}.3.conv", KT4, C[j - 1], C[j - 1]))
        self.final_block = conv(e + "final_block.block.0", K3S1, self.dim, self.dim)
        self.J = emb_off
        # the stacked ResnetBlock.mlp Linear: contiguous views of the flat buffers
        first_w = self.resnets[0]["pre"] + ".mlp.1.weight"
        first_b = self.resnets[0]["pre"] + ".mlp.1.bias"
        K = self.dim + self.S
        ow = self.params[first_w].data_ptr() - self.P.data_ptr()
        ob = self.params[first_b].data_ptr() - self.P.data_ptr()
        self.wcat = self.P[ow // 4: ow // 4 + self.J * K].view(self.J, K)
        self.dwcat = self.G[ow // 4: ow // 4 + self.J * K].view(self.J, K)
        self.bcat = self.P[ob // 4: ob // 4 + self.J]
        self.dbcat = self.G[ob // 4: ob // 4 + self.J]

    def _pack_weights(self):
        self.call("usb_t_cast", _p(self.P), _p(self.P16), self.nparams)
        self.call("usb_t_pack_begin")
        for c in self.convs:
            c.pack(self)
        self.call("usb_t_pack_flush")
        w3 = self.params["estimator.downs.0.0.block1.block.0.weight"]     # (C, 2, 3, 3) -> tap-major (9, 2, C)
        w1 = self.params["estimator.downs.0.0.res_conv.weight"]           # (C, 2, 1, 1) -> (2, C)
        if getattr(self, "first_w3", None) is None:
            self.first_w3 = torch.empty(18, w3.shape[0], dtype=torch.float32, device=self.dev)
            self.first_w1 = torch.empty(2, w3.shape[0], dtype=torch.float32, device=self.dev)
        self.first_w3.view(3, 3, 2, -1).copy_(w3.permute(2, 3, 1, 0))
        self.first_w1.copy_(w1.reshape(-1, 2).t())
        self._packed = True

    # ------------------------------------------------------------------------------------------------ workspace
    def _buf(self, name: str, shape, dtype=torch.float16) -> torch.Tensor:
        t = self._ws.get(name)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._ws[name] = t
        if self._readers and name.startswith("bw."):
            # backward buffers are recycled: wait for side-stream kernels that still read the previous contents
            ev = self._readers.pop(t.data_ptr(), None)
            if ev is not None:
                torch.cuda.current_stream(self.dev).wait_event(ev)
        return t

    def _side(self, fn, *reads):
        """Runs ``fn`` (parameter-gradient kernels: weight gradients, bias column sums) on the side stream, ordered after
        everything enqueued so far; the buffers in ``reads`` are protected until those kernels have finished.  The data
        gradient chain is the critical path of the backward pass and rarely fills the GPU at fine-tuning sizes."""
        if self._trace is not None or not self.overlap_wgrad:
            fn()
            return
        main = torch.cuda.current_stream(self.dev)
        fork = torch.cuda.Event()
        fork.record(main)
        self.side.wait_event(fork)
        with torch.cuda.stream(self.side):
            fn()
            done = torch.cuda.Event()
            done.record(self.side)
        for t in reads:
            self._readers[t.data_ptr()] = done

    def _act(self, name: str, l: int, c: int) -> torch.Tensor:
        return self._buf(f"{name}@{l}x{c}", (self.B, self.Hs[l], self.Ws[l], c))

    def _plan(self, B: int, T: int):
        if self._ws_key == (B, T):
            return
        if T % (1 << (self.L - 1)) or T <= 0:
            raise ValueError("T must be a positive multiple of 2^(len(dim_mults)-1)")
        self._ws = {}
        self._graphs.clear()          # graphs hold pointers into the workspace of their shape
        self._eager_steps.clear()
        self._ws_key = (B, T)
        self.B, self.T = B, T
        self.Hs = [self.n_feats >> l for l in range(self.L)]
        self.Ws = [T >> l for l in range(self.L)]
        n_stats = 2 * len(self.resnets) + 1
        self.stats = torch.zeros(n_stats, B, 8, 2, dtype=torch.int64, device=self.dev)
        self.rows = torch.arange(B, dtype=torch.int32, device=self.dev)
        self.masks = [torch.empty(B, self.Ws[l], dtype=torch.float32, device=self.dev) for l in range(self.L)]
        cmax = max(self.C)
        self.gn_scratch = torch.empty(3 * B * cmax + B * 16, dtype=torch.float32, device=self.dev)
        pmax = self.Hs[0] * self.Ws[0]
        nbytes = max(int(self.lib.usb_t_attn_scratch_bytes(B, _HEADS, self.Hs[l] * self.Ws[l])) for l in range(self.L))
        self.attn_scratch = torch.empty(nbytes // 4 + 16, dtype=torch.float32, device=self.dev)
        self.loss_partial = torch.zeros(512, dtype=torch.float64, device=self.dev)
        self.msum = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
        f = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=self.dev)  # noqa: E731
        self.x0, self.mu, self.z = f(B, self.n_feats, T), f(B, self.n_feats, T), f(B, self.n_feats, T)
        self.t, self.spk = f(B), f(B, self.S)

    # ------------------------------------------------------------------------------------------------ op wrappers
    def _conv(self, kind, in0, c0, in1, c1, l_in, w, wz, bmode, cout, out, bias=None, mask=None, res=None, res_scale=None,
              stats=None):
        H, W = in0.shape[1], in0.shape[2]
        self.call("usb_t_conv", kind, _p(in0), in0.shape[3], c0, _p(in1), in1.shape[3] if in1 is not None else 0, c1,
                  self.B, H, W, _p(w), wz, bmode, cout, _p(bias), _p(mask), _p(res), _p(res_scale), _p(stats), 8, _p(out))
        return out

    def _gn_apply(self, raw, stats, gkey, addvec, res, l, out):
        C = raw.shape[3]
        self.call("usb_t_gn_apply", _p(raw), _p(stats), _p(self.params[gkey + ".weight"]), _p(self.params[gkey + ".bias"]),
                  _p(addvec), self.J, _p(res), _p(self.masks[l]), _p(out), self.B, raw.shape[1], raw.shape[2], C)
        return out

    def _gn_bwd(self, raw, stats, gkey, l, d_raw, dbias, dy=None, dys=None, wvec=None, d_emb=None, d_wvec=None):
        C = raw.shape[3]
        self.call("usb_t_gn_bwd", _p(raw), _p(stats), _p(self.params[gkey + ".weight"]), _p(self.params[gkey + ".bias"]),
                  _p(dy), None, _p(dys), _p(wvec), _p(self.masks[l]), _p(self.gn_scratch), _p(d_raw), _p(dbias),
                  _p(self.grads[gkey + ".weight"]), _p(self.grads[gkey + ".bias"]), _p(d_emb), self.J, _p(d_wvec),
                  self.B, raw.shape[1], raw.shape[2], C)
        return d_raw

    def _wgrad(self, kind, dy, x, cout, cs, ci0, cin_total, dw, per_sample=0, side=True):
        def fn():
            # flag bit 1: every weight-gradient slice is written once per backward pass, right after zero_grad
            self.call("usb_t_wgrad", kind, _p(dy), dy.shape[3], _p(x), x.shape[3], self.B, x.shape[1], x.shape[2], cout, cs,
                      ci0, cin_total, _p(dw), per_sample | 2)
        if side:
            self._side(fn, dy)
        else:
            fn()

    def _colsum(self, t, c, out, stride_n=0, side=True):
        def fn():
            self.call("usb_t_colsum", _p(t), t.shape[3], self.B, t.shape[1] * t.shape[2], c, _p(out), stride_n)
        if side:
            self._side(fn, t)
        else:
            fn()

    def _add(self, a, b, out, c=None):
        self.call("usb_t_add", _p(a), _p(b), _p(c), _p(out), a.numel())
        return out

    # ------------------------------------------------------------------------------------------------ forward
    def _resnet_fwd(self, i: int, l: int, in0, in1=None):
        r = self.resnets[i]
        pre, co = r["pre"], r["cout"]
        s1, s2 = self.stats[2 * i], self.stats[2 * i + 1]
        raw1 = self._act(f"r{i}.raw1", l, co)
        h1 = self._act(f"r{i}.h1", l, co)
        raw2 = self._act(f"r{i}.raw2", l, co)
        out = self._act(f"r{i}.out", l, co)
        E = self.E[:, r["emb_off"]:]
        if r["first"]:
            resid = self._act(f"r{i}.res", l, co)
            self.call("usb_t_first_conv", _p(self.xt), _p(self.mu), _p(self.rows), _p(self.masks[0]), _p(self.first_w3),
                      _p(self.params[pre + ".block1.block.0.bias"]), _p(self.first_w1), _p(self.params[pre + ".res_conv.bias"]),
                      _p(raw1), _p(resid), _p(s1), self.B, self.Hs[0], self.Ws[0], co)
        else:
            c0 = in0.shape[3]
            c1 = in1.shape[3] if in1 is not None else 0
            self._conv(K3S1, in0, c0, in1, c1, l, r["c1"].fwd, 1, 0, co, raw1, bias=r["c1"].b, stats=s1)
            if r["has_res"]:
                resid = self._act(f"r{i}.res", l, co)
                self._conv(K1, in0, c0, in1, c1, l, r["res"].fwd, 1, 0, co, resid, bias=r["res"].b)
            else:
                resid = in0
        self._gn_apply(raw1, s1, pre + ".block1.block.1", E, None, l, h1)
        self._conv(K3S1, h1, co, None, 0, l, r["c2"].fwd, 1, 0, co, raw2, bias=r["c2"].b, stats=s2)
        self._gn_apply(raw2, s2, pre + ".block2.block.1", None, resid, l, out)
        r["saved"] = (l, in0, in1, raw1, h1, raw2)
        return out

    def _attn_fwd(self, i: int, l: int, x):
        a = self.attns[i]
        C, pre = a["C"], a["pre"]
        P = self.Hs[l] * self.Ws[l]
        qkv = self._act(f"a{i}.qkv", l, 3 * _HID)
        weff = self._buf(f"a{i}.weff", (self.B, C, _HID))
        ctx = self._buf(f"a{i}.ctx", (self.B, _HEADS, 32, 32), torch.float32)
        ms = self._buf(f"a{i}.ms", (self.B, _HEADS, 2, 32), torch.float32)
        out = self._act(f"a{i}.out", l, C)
        self._conv(K1, x, C, None, 0, l, a["qkv"].fwd, 1, 0, 3 * _HID, qkv)
        self.call("usb_t_attn_context", _p(qkv), 3 * _HID, _HID, 2 * _HID, _p(self.params[pre + ".fn.fn.to_out.weight"]),
                  _p(self.attn_scratch), _p(weff), _p(ctx), _p(ms), self.B, P, C, _HEADS)
        self._conv(K1, qkv, _HID, None, 0, l, weff, self.B, 2, C, out, bias=self.params[pre + ".fn.fn.to_out.bias"],
                   mask=self.masks[l], res=x, res_scale=self.params[pre + ".fn.g"])
        a["saved"] = (l, x, qkv, ctx, ms)
        return out

    def forward(self, x0, mask, cond, t, spk_emb, z):
        """loss_t (unitspeech.py:393-405) with the N(0,1) draw ``z`` of forward_diffusion (:381) supplied by the caller.
        x0, cond, z: (B, n_feats, T); mask: (B, 1, T); t: (B,); spk_emb: (B, 1, S).  Returns the device loss scalar."""
        self.set_inputs(x0, mask, cond, t, spk_emb, z)
        return self._forward_body()

    def set_inputs(self, x0, mask, cond, t, spk_emb, z):
        """Copies one batch into the step's static device buffers (the captured CUDA graph reads them)."""
        B, F, T = x0.shape
        self._plan(B, T)
        self.generation += 1          # stamps the activations the next backward() will read
        self.x0.copy_(x0.detach())
        self.mu.copy_(cond.detach())
        self.z.copy_(z.detach())
        self.t.copy_(t.detach().reshape(B))
        self.spk.copy_(spk_emb.detach().reshape(B, self.S))
        self.masks[0].copy_(mask.detach().reshape(B, T))

    def _forward_body(self):
        B, F, T = self.B, self.n_feats, self.T
        if not self._packed:
            self._pack_weights()
        x0, z, t = self.x0, self.z, self.t
        for l in range(1, self.L):
            self.masks[l].copy_(self.masks[l - 1][:, ::2])
        self.xt = self._buf("xt", (B, F, T), torch.float32)
        self.zm = self._buf("zm", (B, F, T), torch.float32)
        self.call("usb_forward_diffusion", _p(x0), _p(self.masks[0]), _p(t), _p(z), _p(self.xt), _p(self.zm), B, T)
        self.stats.zero_()
        e = "estimator."
        P = self.params
        self.u = self._buf("u", (B, self.dim + self.S), torch.float32)
        self.E = self._buf("E", (B, self.J), torch.float32)
        self.call("usb_t_embed", _p(t), _p(self.spk), _p(self.freqs), _p(P[e + "mlp.0.weight"]), _p(P[e + "mlp.0.bias"]),
                  _p(P[e + "mlp.2.weight"]), _p(P[e + "mlp.2.bias"]), _p(self.wcat), _p(self.bcat), _p(self.u), _p(self.E),
                  B, self.J)
        L, C = self.L, self.C
        ri = ai = 0
        x = None
        self.skips = []
        self.up_in = {}
        for k in range(L):
            y0 = self._resnet_fwd(ri, k, x); ri += 1
            y1 = self._resnet_fwd(ri, k, y0); ri += 1
            skip = self._attn_fwd(ai, k, y1); ai += 1
            self.skips.append(skip)
            if k < L - 1:
                d = self.down[k]
                x = self._conv(K3S2, skip, C[k], None, 0, k, d.fwd, 1, 0, C[k], self._act(f"down{k}", k + 1, C[k]), bias=d.b,
                               mask=self.masks[k + 1])
        D = L - 1
        m0 = self._resnet_fwd(ri, D, self.skips[D]); ri += 1
        m1 = self._attn_fwd(ai, D, m0); ai += 1
        cur = self._resnet_fwd(ri, D, m1); ri += 1
        for k in range(L - 1):
            j = D - k
            ya = self._resnet_fwd(ri, j, cur, self.skips[j]); ri += 1
            yb = self._resnet_fwd(ri, j, ya); ri += 1
            at = self._attn_fwd(ai, j, yb); ai += 1
            u = self.up[k]
            cur = self._conv(KT4, at, C[j - 1], None, 0, j, u.fwd, 4, 1, C[j - 1], self._act(f"up{k}", j - 1, C[j - 1]),
                             bias=u.b, mask=self.masks[j - 1])
            self.up_in[k] = at
        self.final_in = cur
        fs = self.stats[2 * len(self.resnets)]
        self.final_raw = self._act("final.raw", 0, self.dim)
        self._conv(K3S1, cur, self.dim, None, 0, 0, self.final_block.fwd, 1, 0, self.dim, self.final_raw, bias=self.final_block.b,
                   stats=fs)
        self.score = self._buf("score", (B, F, T), torch.float32)
        self.call("usb_t_final", _p(self.final_raw), _p(fs), _p(P[e + "final_block.block.1.weight"]),
                  _p(P[e + "final_block.block.1.bias"]), _p(P[e + "final_conv.weight"]), _p(P[e + "final_conv.bias"]),
                  _p(self.masks[0]), _p(self.score), B, self.Hs[0], self.Ws[0], self.dim)
        self.call("usb_t_loss", _p(self.score), _p(self.zm), _p(self.masks[0]), _p(t), _p(self.loss_partial), _p(self.loss), B, T)
        return self.loss

    # ------------------------------------------------------------------------------------------------ backward
    def _conv_dgrad(self, c: _Conv, si: int, d_out, l_out: int, name: str, res=None):
        """data gradient of conv ``c`` w.r.t. the ``si``-th input-channel slice; masked with the mask of the input level."""
        c0, c1 = c.splits[si]
        cs = c1 - c0
        if c.kind == K3S1 or c.kind == K1:
            out = self._act(name, l_out, cs)
            return self._conv(c.kind, d_out, c.cout, None, 0, l_out, c.dgrad[si], 1, 0, cs, out, mask=self.masks[l_out], res=res)
        if c.kind == K3S2:      # d_out lives one level below the input
            out = self._act(name, l_out, cs)
            return self._conv(K3S2D, d_out, c.cout, None, 0, l_out, c.dgrad[si], 4, 1, cs, out, mask=self.masks[l_out])
        out = self._act(name, l_out, cs)   # KT4: d_out lives one level above the input
        return self._conv(KT4D, d_out, c.cout, None, 0, l_out, c.dgrad[si], 1, 0, cs, out, mask=self.masks[l_out])

    def _resnet_bwd(self, i: int, d_out):
        r = self.resnets[i]
        pre, co = r["pre"], r["cout"]
        l, in0, in1, raw1, h1, raw2 = r["saved"]
        s1, s2 = self.stats[2 * i], self.stats[2 * i + 1]
        d_raw2 = self._act("bw.d_raw2", l, co)     # (two buffers: the side-stream weight gradients still read d_raw2)
        self._gn_bwd(raw2, s2, pre + ".block2.block.1", l, d_raw2, r["c2"].db, dy=d_out)
        self._wgrad(K3S1, d_raw2, h1, co, co, 0, co, r["c2"].dw)
        d_h1 = self._conv_dgrad(r["c2"], 0, d_raw2, l, "bw.d_h1")
        d_raw1 = self._act("bw.d_raw1", l, co)
        b1 = self.grads[pre + ".block1.block.0.bias"]
        self._gn_bwd(raw1, s1, pre + ".block1.block.1", l, d_raw1, b1, dy=d_h1, d_emb=self.dE[:, r["emb_off"]:])
        if r["first"]:
            self._side(lambda: self.call(
                "usb_t_first_conv_wgrad", _p(d_raw1), _p(d_out), None, _p(self.xt), _p(self.mu), _p(self.masks[0]),
                _p(self.grads[pre + ".block1.block.0.weight"]), _p(self.grads[pre + ".res_conv.weight"]), self.B,
                self.Hs[0], self.Ws[0], co), d_raw1, d_out)
            self._colsum(d_out, co, self.grads[pre + ".res_conv.bias"])
            return None, None
        c0 = in0.shape[3]
        cin = r["cin"]
        self._wgrad(K3S1, d_raw1, in0, co, c0, 0, cin, r["c1"].dw)
        if in1 is not None:
            self._wgrad(K3S1, d_raw1, in1, co, in1.shape[3], c0, cin, r["c1"].dw)
        outs = []
        if r["has_res"]:
            self._colsum(d_out, co, r["res"].db)
            self._wgrad(K1, d_out, in0, co, c0, 0, cin, r["res"].dw)
            if in1 is not None:
                self._wgrad(K1, d_out, in1, co, in1.shape[3], c0, cin, r["res"].dw)
        for si in range(1 if in1 is None else 2):
            t = self._conv_dgrad(r["c1"], si, d_raw1, l, f"bw.r{i}.t{si}")
            if r["has_res"]:
                outs.append(self._conv_dgrad(r["res"], si, d_out, l, f"bw.r{i}.din{si}", res=t))
            else:
                outs.append(self._add(t, d_out, self._act(f"bw.r{i}.din{si}", l, c0)))
        return outs[0], (outs[1] if len(outs) > 1 else None)

    def _attn_bwd(self, i: int, d_out):
        a = self.attns[i]
        C, pre = a["C"], a["pre"]
        l, x, qkv, ctx, ms = a["saved"]
        P = self.Hs[l] * self.Ws[l]
        G, g = self.grads, self.params
        cs = self._buf(f"bw.a.cs{C}", (self.B, C), torch.float32)
        Gm = self._buf(f"bw.a.G{C}", (self.B, C, _HID), torch.float32)
        cs.zero_()
        Gm.zero_()
        self._colsum(d_out, C, cs, C, side=False)                                   # inputs of attn_bwd_small: critical path
        self._wgrad(K1, d_out, qkv, C, _HID, 0, _HID, Gm, per_sample=1, side=False)
        dctx = self._buf("bw.a.dctx", (self.B, _HEADS, 32, 32), torch.float32)
        weffT = self._buf(f"bw.a.weffT{C}", (self.B, _HID, C))
        self.call("usb_t_attn_bwd_small", _p(Gm), _p(cs), _p(g[pre + ".fn.fn.to_out.weight"]), _p(g[pre + ".fn.fn.to_out.bias"]),
                  _p(g[pre + ".fn.g"]), _p(ctx), _p(G[pre + ".fn.fn.to_out.weight"]), _p(G[pre + ".fn.fn.to_out.bias"]),
                  _p(G[pre + ".fn.g"]), _p(dctx), _p(weffT), self.B, C, _HEADS)
        dq = self._act("bw.a.dq", l, _HID)
        dkv = self._act("bw.a.dkv", l, 2 * _HID)
        self._conv(K1, d_out, C, None, 0, l, weffT, self.B, 2, _HID, dq)
        self.call("usb_t_attn_bwd_dkv", _p(qkv), 3 * _HID, _HID, 2 * _HID, _p(ms), _p(ctx), _p(dctx), _p(dkv), self.B, P, _HEADS)
        dwq = a["qkv"].dw
        self._wgrad(K1, dq, x, _HID, C, 0, C, dwq)
        self._wgrad(K1, dkv, x, 2 * _HID, C, 0, C, dwq.view(-1)[_HID * C:])
        d_x = self._act(f"bw.a{i}.dx", l, C)
        self._conv(K1, dq, _HID, dkv, 2 * _HID, l, a["qkv"].dgrad[0], 1, 0, C, d_x, mask=self.masks[l], res=d_out)
        return d_x

    def backward(self):
        """Gradients of ``loss_scale * loss`` of the last ``forward`` for every parameter, written into ``self.grads``.
        Call ``zero_grad()`` before every pass: weight-gradient tiles that a single CTA produces are stored, not added."""
        B, T, L, C = self.B, self.T, self.L, self.C
        e = "estimator."
        P, G = self.params, self.grads
        dscore = self._buf("bw.dscore", (B, self.n_feats, T), torch.float32)
        self.call("usb_t_loss_grad", _p(self.score), _p(self.zm), _p(self.masks[0]), _p(self.t), self.loss_scale, _p(self.msum),
                  _p(dscore), B, T)
        self.dE = self._buf("bw.dE", (B, self.J), torch.float32)
        self.call("usb_t_dot", _p(dscore), None, dscore.numel(), _p(G[e + "final_conv.bias"]))
        fs = self.stats[2 * len(self.resnets)]
        d_raw = self._act("bw.d_raw", 0, self.dim)
        self._gn_bwd(self.final_raw, fs, e + "final_block.block.1", 0, d_raw, self.final_block.db, dys=dscore,
                     wvec=P[e + "final_conv.weight"], d_wvec=G[e + "final_conv.weight"])
        self._wgrad(K3S1, d_raw, self.final_in, self.dim, self.dim, 0, self.dim, self.final_block.dw)
        d_cur = self._conv_dgrad(self.final_block, 0, d_raw, 0, "bw.d_final_in")
        ri, ai = len(self.resnets) - 1, len(self.attns) - 1
        D = L - 1
        d_skip = [None] * L
        for k in reversed(range(L - 1)):
            j = D - k
            u = self.up[k]
            self._colsum(d_cur, C[j - 1], u.db)
            self._wgrad(KT4, d_cur, self.up_in[k], C[j - 1], C[j - 1], 0, C[j - 1], u.dw)
            d_at = self._conv_dgrad(u, 0, d_cur, j, f"bw.up{k}.din")
            d_yb = self._attn_bwd(ai, d_at); ai -= 1
            d_ya, _ = self._resnet_bwd(ri, d_yb); ri -= 1
            d_cur, d_skip[j] = self._resnet_bwd(ri, d_ya); ri -= 1
        d_m1, _ = self._resnet_bwd(ri, d_cur); ri -= 1
        d_m0 = self._attn_bwd(ai, d_m1); ai -= 1
        d_s, _ = self._resnet_bwd(ri, d_m0); ri -= 1
        d_x = None
        for k in reversed(range(L)):
            if k == D:
                d_sk = self._add(d_s, d_skip[k], self._act(f"bw.dskip{k}", k, C[k])) if d_skip[k] is not None else d_s
            else:
                d = self.down[k]
                self._colsum(d_x, C[k], d.db)
                self._wgrad(K3S2, d_x, self.skips[k], C[k], C[k], 0, C[k], d.dw)
                t = self._conv_dgrad(d, 0, d_x, k, f"bw.down{k}.din")
                d_sk = self._add(t, d_skip[k], self._act(f"bw.dskip{k}", k, C[k])) if d_skip[k] is not None else t
            d_y1 = self._attn_bwd(ai, d_sk); ai -= 1
            d_y0, _ = self._resnet_bwd(ri, d_y1); ri -= 1
            d_x, _ = self._resnet_bwd(ri, d_y0); ri -= 1
        self.call("usb_t_embed_bwd", _p(self.t), _p(self.spk), _p(self.freqs), _p(P[e + "mlp.0.weight"]), _p(P[e + "mlp.0.bias"]),
                  _p(P[e + "mlp.2.weight"]), _p(P[e + "mlp.2.bias"]), _p(self.wcat), _p(self.u), _p(self.dE),
                  _p(self._buf("bw.du", (B, self.dim + self.S), torch.float32)), _p(G[e + "mlp.0.weight"]), _p(G[e + "mlp.0.bias"]),
                  _p(G[e + "mlp.2.weight"]), _p(G[e + "mlp.2.bias"]), _p(self.dwcat), _p(self.dbcat), B, self.J)
        if self._trace is None and self.overlap_wgrad:
            torch.cuda.current_stream(self.dev).wait_stream(self.side)      # join: every parameter gradient is complete
            self._readers.clear()

    # ------------------------------------------------------------------------------------------------ optimizer
    def optimizer_step(self):
        """clip_grad_norm_(max_norm) + Adam over all parameters in two launches (finetune.py:163-165)."""
        self.sumsq.zero_()
        self.call("usb_t_sumsq", _p(self.G), self.nparams, _p(self.sumsq))
        self.call("usb_t_adam", _p(self.P), _p(self.G), _p(self.M), _p(self.V), self.nparams, self.lr, self.betas[0], self.betas[1],
                  self.eps, 0, _p(self.step_dev), _p(self.sumsq), 1.0 / self.loss_scale, self.max_norm if self.max_norm else 0.0,
                  _p(self.skipped))
        self._packed = False

    def grad_norm(self) -> float:
        """L2 norm of the (unscaled) gradients of the last optimizer_step (host sync)."""
        return float(self.sumsq.sqrt().item()) / self.loss_scale

    @property
    def step_count(self) -> int:
        return int(self.step_dev.item())

    def _step_body(self):
        self.zero_grad()
        self._forward_body()
        self.backward()
        self.optimizer_step()
        self._pack_weights()          # fp16 operands of the updated weights, ready for the next step / the sampler

    @property
    def skipped_steps(self) -> int:
        """Adam steps skipped because the (loss-scaled) gradient norm was not finite (host sync).  The loss scale is fixed
        (default 8192, no GradScaler-style back-off): a count that keeps growing means the fp16 activation gradients
        overflow at this scale -- lower ``loss_scale``."""
        return int(self.skipped.item())

    def train_step(self, x0, mask, cond, t, spk_emb, z) -> torch.Tensor:
        """zero_grad -> loss_t -> backward -> clip -> Adam.  Returns the device loss scalar (before the update) -- the
        tuner's own loss buffer, overwritten by the next step: ``float()`` or ``.clone()`` it to keep a value.
        After two eager steps per (B, T) the whole step is captured in a CUDA graph and replayed (the step is ~900 short
        launches; replaying removes the host from the loop).  All state the graph touches lives in buffers owned by this
        object; the Adam step number is a device counter."""
        self.set_inputs(x0, mask, cond, t, spk_emb, z)
        key = (self.B, self.T)
        g = self._graphs.get(key)
        if g is not None:
            if not self._packed:      # weights were replaced from outside (load_state_dict) since the last step
                self._pack_weights()
            g.replay()
            return self.loss
        n = self._eager_steps.get(key, 0)
        if not self.use_cuda_graph or n < 2:
            self._eager_steps[key] = n + 1
            self._step_body()
            return self.loss
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_body()
        self._graphs[key] = g
        g.replay()
        return self.loss

    def fine_tune(self, cond_x, y, y_mask, y_lengths, y_max_length, attn, spk_emb, segment_size, n_feats, offset=1e-5):
        """One iteration of the reference loop (finetune.py:131-165) with the signature of UnitSpeech.fine_tune
        (unitspeech/unitspeech.py:452-492): crop, align, t ~ U(offset, 1-offset) and z ~ N(0,1) in the reference's RNG
        order (compute_loss :408 then forward_diffusion :381), then zero_grad + loss + backward + clip + Adam.
        Returns the loss (device scalar) evaluated before the update."""
        from .decoder import crop_segments
        y_cut, y_cut_mask, cond_y = crop_segments(cond_x, y, y_mask, y_lengths, y_max_length, attn, segment_size, n_feats)
        t = torch.rand(y_cut.shape[0], dtype=y_cut.dtype, device=y_cut.device)
        t = torch.clamp(t, offset, 1.0 - offset)
        z = torch.randn(y_cut.shape, dtype=y_cut.dtype, device=y_cut.device)
        return self.train_step(y_cut, y_cut_mask, cond_y, t, spk_emb, z)

    def unscaled_grads(self) -> Dict[str, torch.Tensor]:
        return {k: self.grad_reference(k) / self.loss_scale for k in self.grads}
