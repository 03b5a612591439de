"""Runs the fused Snake activation kernel alone on the activation shapes of the public BigVGAN config at 16 x 512 mel
frames (stage 1: 384 ch, stage 3: 96 ch, stage 5: 24 ch), twice each (first pass = warm-up).  Used under ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from unitspeech_b200 import abi

lib = abi.load_library()
SHAPES = [(16, 8192, 384, 384), (16, 32768, 128, 96), (16, 131072, 64, 24)]
bufs = []
for N, L, C, Cr in SHAPES:
    x = (torch.randn(N, L, C, device="cuda") * 1.5).half()
    bufs.append((x, torch.empty_like(x), torch.rand(C, device="cuda") + 0.5, torch.rand(C, device="cuda") + 0.5))
for rep in range(2):
    for (N, L, C, Cr), (x, out, al, ib) in zip(SHAPES, bufs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        abi.check(lib.usb_op_snake_act(x.data_ptr(), al.data_ptr(), ib.data_ptr(), N, L, C, Cr, out.data_ptr(),
                                       int(torch.cuda.current_stream().cuda_stream)))
        e1.record()
        torch.cuda.synchronize()
        if rep:
            ms = e0.elapsed_time(e1)
            print(f"N={N} L={L} C={C} real={Cr}: {ms*1e3:.1f} us, {4.0*N*L*Cr/ms/1e6:.0f} GB/s (real fp16 read+write)")
