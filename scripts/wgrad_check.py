"""Weight-gradient kernel check (GPU box): usb_t_wgrad against torch autograd (fp32 on the same fp16-rounded tensors) for
the four conv kinds, skip-concat slices and the per-sample 1x1 form; then timings on the fine-tune shapes.
USB_WGRAD_MMA=1 selects the mma.sync kernel instead of the tcgen05 one."""

import ctypes
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unitspeech_b200.training import FineTuner, _p, to_reference_layout  # noqa: E402

ft = FineTuner(dim=64, dim_mults=(1, 2))
dev = ft.dev
g = torch.Generator(device="cpu").manual_seed(0)


def nhwc16(t):
    return t.permute(0, 2, 3, 1).contiguous().to(dev, torch.float16)


def run(kind, N, H, W, Cin, Cout, ci0=0, Cs=None, per_sample=False):
    Cs = Cs or Cin
    x = torch.randn(N, Cs, H, W, generator=g).half().float()
    Ho, Wo = (H // 2, W // 2) if kind == 1 else ((2 * H, 2 * W) if kind == 3 else (H, W))
    dy = torch.randn(N, Cout, Ho, Wo, generator=g).half().float()
    xs = x.to(dev).requires_grad_(False)
    if kind == 3:
        w = torch.zeros(Cs, Cout, 4, 4, device=dev, requires_grad=True)
        y = F.conv_transpose2d(xs, w, stride=2, padding=1)
    else:
        k = 1 if kind == 2 else 3
        w = torch.zeros(Cout, Cs, k, k, device=dev, requires_grad=True)
        y = F.conv2d(xs, w, stride=2 if kind == 1 else 1, padding=k // 2)
    if per_sample:
        ref = torch.stack([torch.autograd.grad(F.conv2d(xs[n:n + 1], w), w, dy[n:n + 1].to(dev))[0].reshape(Cout, Cs) for n in range(N)])
        out = torch.zeros(N, Cout, Cs, device=dev)
    else:
        ref = torch.autograd.grad(y, w, dy.to(dev))[0]
        shape = (Cin, Cout, 4, 4) if kind == 3 else (Cout, Cin, ref.shape[2], ref.shape[3])
        out = torch.zeros(shape, device=dev).reshape(-1)      # training layout (forward operand order)
    a, b = nhwc16(dy), nhwc16(x)
    ft.B = N
    ft.call("usb_t_wgrad", kind, _p(a), Cout, _p(b), Cs, N, H, W, Cout, Cs, ci0, Cin, _p(out), 1 if per_sample else 0)
    torch.cuda.synchronize()
    if not per_sample:
        out = to_reference_layout(kind, out, shape)
    got = out if per_sample else (out[ci0:ci0 + Cs] if kind == 3 else out[:, ci0:ci0 + Cs])
    err = float((got - ref).norm() / ref.norm())
    other = float(out.norm() ** 2 - got.norm() ** 2)
    print(f"kind {kind} N{N} {H}x{W} Cin {Cin}[{ci0}:{ci0 + Cs}] Cout {Cout} per_sample {int(per_sample)}: rel err {err:.2e}"
          f" (outside slice {other:.1e}) {'OK' if err < 2e-3 and abs(other) < 1e-3 else 'FAIL'}")
    return err < 2e-3


def timeit(kind, N, H, W, Cin, Cout, iters=10):
    Ho, Wo = (H // 2, W // 2) if kind == 1 else ((2 * H, 2 * W) if kind == 3 else (H, W))
    a = torch.randn(N, Ho, Wo, Cout, device=dev).half()
    b = torch.randn(N, H, W, Cin, device=dev).half()
    taps = {0: 9, 1: 9, 2: 1, 3: 16}[kind]
    out = torch.zeros(Cout * Cin * taps, device=dev)
    for _ in range(2):
        ft.call("usb_t_wgrad", kind, _p(a), Cout, _p(b), Cin, N, H, W, Cout, Cin, 0, Cin, _p(out), 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ft.call("usb_t_wgrad", kind, _p(a), Cout, _p(b), Cin, N, H, W, Cout, Cin, 0, Cin, _p(out), 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    px = N * (Ho * Wo if kind != 3 else H * W)
    fl = 2.0 * px * Cout * Cin * (taps if kind != 3 else 16)
    print(f"time kind {kind} N{N} {H}x{W} {Cin}->{Cout}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s")


ok = True
ok &= run(2, 2, 16, 24, 64, 128)
ok &= run(0, 2, 16, 24, 64, 128)
ok &= run(0, 2, 10, 22, 128, 64)
ok &= run(0, 2, 16, 24, 256, 128, ci0=128, Cs=128)
ok &= run(1, 2, 16, 24, 128, 128)
ok &= run(3, 2, 8, 12, 128, 128)
ok &= run(2, 3, 16, 24, 128, 256, per_sample=True)
ok &= run(0, 2, 20, 44, 512, 256)
print("ALL OK" if ok else "SOME FAILED")
if len(sys.argv) > 1:
    for (kind, H, W, Cin, Cout) in ((0, 80, 176, 128, 128), (0, 40, 88, 256, 256), (0, 20, 44, 512, 512), (0, 10, 22, 1024, 1024),
                                    (0, 10, 22, 2048, 512), (1, 80, 176, 128, 128), (3, 40, 88, 128, 128), (2, 80, 176, 128, 384)):
        timeit(kind, 8, H, W, Cin, Cout)
