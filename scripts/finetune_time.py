"""BASELINE.json configs[4]: speaker-adaptive fine-tuning step -- U-Net forward + backward of the diffusion score loss on
176-frame crops (fix_len_compatibility(2*22050//256)) of a synthetic 10 s reference, batch 8, Adam lr 2e-5.
Times `iters` full steps (zero_grad + loss + backward + clip + Adam) with CUDA events after warm-up.

usage: python scripts/finetune_time.py [iters=50] [B=8] [T=176]"""

import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unitspeech_b200 import UnitSpeech  # noqa: E402
from unitspeech_b200.synthetic import random_init_state_dict  # noqa: E402
from unitspeech_b200.training import FineTuner  # noqa: E402

CONV_MFLOP_PER_FRAME = 647.27   # conv FLOPs of one estimator evaluation per mel frame


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 176
    dec = UnitSpeech(80, 128, (1, 2, 4, 8), spk_emb_dim=256)
    sd = random_init_state_dict(dec, seed=1234, out_scale=4.0)
    ft = FineTuner(lr=2e-5, use_cuda_graph=not os.environ.get('FT_NO_GRAPH'))
    ft.load_state_dict(sd)
    ft.overlap_wgrad = not os.environ.get('FT_NO_OVERLAP')
    g = torch.Generator().manual_seed(0)
    x0 = (torch.randn(B, 80, T, generator=g) * 0.5).clamp(-1, 1).cuda()
    cond = torch.randn(B, 80, T, generator=g).clamp(-1, 1).cuda()
    mask = torch.ones(B, 1, T).cuda()
    spk = torch.randn(B, 1, 256, generator=g)
    spk = (spk / spk.norm(dim=-1, keepdim=True)).cuda()
    zs = torch.randn(4, B, 80, T, generator=g).cuda()
    ts = torch.rand(4, B, generator=g).clamp(1e-5, 1 - 1e-5).cuda()
    losses = []
    for i in range(int(os.environ.get('FT_WARMUP', '5'))):
        losses.append(float(ft.train_step(x0, mask, cond, ts[i % 4], spk, zs[i % 4])))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for i in range(iters):
        ft.train_step(x0, mask, cond, ts[i % 4], spk, zs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - w0) * 1e3 / iters
    ms = e0.elapsed_time(e1) / iters
    flops = 3.0 * CONV_MFLOP_PER_FRAME * 1e6 * B * T
    print(json.dumps({"workload": f"fine-tune step B{B} x {T} frames, dim 128 (1,2,4,8), Adam", "ms_per_iter": round(ms, 3),
                      "wall_ms_per_iter": round(wall, 3), "iters_per_s": round(1e3 / ms, 2),
                      "conv_tflops_fwd_bwd": round(flops / ms / 1e9, 1), "first_losses": [round(v, 4) for v in losses],
                      "loss_after": round(float(ft.loss), 4), "skipped": int(ft.skipped),
                      "est_500_iters_s": round(ms * 0.5, 2)}))


if __name__ == "__main__":
    main()
