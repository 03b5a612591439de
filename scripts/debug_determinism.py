"""Debug: run-to-run determinism of single ops and of the estimator."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import unitspeech_oracle as O
from gpu_ops import OpHandle
from unitspeech_b200 import UnitSpeech

ops = OpHandle(0)
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 128, 40, 56, generator=g); w = torch.randn(256, 128, 3, 3, generator=g) / 34; b = torch.randn(256, generator=g)
o1, s1 = ops.conv(0, x, w, b, stats_groups=8)
for i in range(3):
    o2, s2 = ops.conv(0, x, w, b, stats_groups=8)
    print("conv repeat diff", float((o1 - o2).abs().max()), "stats rel diff", float(((s1 - s2).abs() / s1.abs().clamp_min(1)).max()))
qkv = torch.randn(2, 384, 40, 64, generator=g); wo = torch.randn(256, 128, generator=g)
a1 = ops.attn_context(qkv, wo)
for i in range(2):
    print("attn repeat diff", float((a1 - ops.attn_context(qkv, wo)).abs().max()))

for dim, mults, T in ((64, (1, 2), 16), (128, (1, 2, 4, 8), 32)):
    p = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=1 / 32)
    dec = UnitSpeech(80, dim, mults, spk_emb_dim=256); dec.load_state_dict(p); dec = dec.cuda()
    z, mask, cond, spk, noise = O.harness_inputs(2, T, 3, seed=2, scale=1 / 32, lengths=(T, T - 5))
    zc, mc, cc, sc, nc = z.cuda(), mask.cuda(), cond.cuda(), spk.cuda(), noise.cuda()
    t = torch.full((2,), 0.37).cuda()
    e1 = dec.estimator(zc, mc, cc, t, sc)
    ref = O.estimator_forward(p, z, mask, cond, torch.full((2,), 0.37), spk, dim, mults)
    print(dim, "est vs oracle", float((e1.cpu() - ref).abs().max()), "scale", float(ref.abs().max()))
    for i in range(3):
        e2 = dec.estimator(zc, mc, cc, t, sc)
        print(dim, "estimator repeat diff", float((e1 - e2).abs().max()))
    r1 = dec(zc, mc, cc, sc, 3, 1.0, 1.0, noise=nc)
    for i in range(3):
        r2 = dec(zc, mc, cc, sc, 3, 1.0, 1.0, noise=nc)
        print(dim, "sampler repeat diff", float((r1 - r2).abs().max()), "scale", float(r1.abs().max()))
    # batch 1 vs batch 2 estimator
    e_b1 = dec.estimator(zc[:1], mc[:1], cc[:1], t[:1], sc[:1])
    print(dim, "estimator B=1 vs B=2 row0 diff", float((e_b1 - e1[:1]).abs().max()))
