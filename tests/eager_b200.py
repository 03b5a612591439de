"""Second baseline of BASELINE.md section 3: the reference algorithm (the oracle's functional torch restatement, pinned to
the reference) run EAGERLY ON THE B200 by PyTorch/cuDNN -- what a user of the reference gets on this GPU today.
Checker-side measurement only (lives under tests/, imports oracle/); nothing in the product path uses it.

  python tests/eager_b200.py [utterances=2] [frames=512]

Prints one JSON line: frames/s of the 50-step CFG sampler (per-utterance B=1 loop, the only batch the reference supports),
with TF32 convs off (fp32, the parity oracle) and on (PyTorch default), and it/s of the fine-tune iteration
(8 x 176 frames, autograd + clip_grad_norm_ + Adam)."""

import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unitspeech_oracle as O  # noqa: E402


def sampler_fps(p, n_utt, T, n=50):
    z, mask, cond, spk, noise = O.harness_inputs(n_utt, T, n, seed=0, scale=1.0 / 512)
    dev = torch.device("cuda")
    mv = lambda a: a.to(dev)  # noqa: E731
    def run():
        for b in range(n_utt):
            O.reverse_diffusion(p, mv(z[b:b + 1]), mv(mask[b:b + 1]), mv(cond[b:b + 1]), mv(spk[b:b + 1]), n, 1.0, 1.0,
                                noise=mv(noise[:, b:b + 1]))
    with torch.no_grad():
        O.reverse_diffusion(p, mv(z[:1]), mv(mask[:1]), mv(cond[:1]), mv(spk[:1]), 3, 1.0, 1.0, noise=mv(noise[:3, :1]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run()
        torch.cuda.synchronize()
    return n_utt * T / (time.perf_counter() - t0)


def finetune_its(p, B=8, T=176, iters=5):
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(11)
    x0 = (torch.randn(B, 80, T, generator=g) * 0.5).clamp(-1, 1).to(dev)
    cond = torch.randn(B, 80, T, generator=g).clamp(-1, 1).to(dev)
    mask = torch.ones(B, 1, T, device=dev)
    spk = torch.randn(B, 1, 256, generator=g)
    spk = (spk / spk.norm(dim=-1, keepdim=True)).to(dev)
    z = torch.randn(B, 80, T, generator=g).to(dev)
    t = torch.rand(B, generator=g).clamp(1e-5, 1 - 1e-5).to(dev)
    w = {k: torch.nn.Parameter(v.clone()) for k, v in p.items()}
    opt = torch.optim.Adam(list(w.values()), lr=2e-5)
    def step():
        opt.zero_grad()
        loss, _ = O.loss_t.__wrapped__(w, x0, mask, cond, t, spk, z)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([v for v in w.values() if v.grad is not None], max_norm=1)
        opt.step()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    return iters / (time.perf_counter() - t0)


def main():
    n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    p = {k: v.cuda() for k, v in O.harness_params(seed=1234, out_scale=1.0 / 512).items()}
    out = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__, "utterances": n_utt, "frames": T}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        out[f"sampler_frames_per_s_{name}"] = round(sampler_fps(p, n_utt, T), 1)
        pt = {k: (v * 2048 if k.startswith("estimator.final_conv") else v) for k, v in p.items()}
        out[f"finetune_iters_per_s_{name}"] = round(finetune_its(pt), 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
