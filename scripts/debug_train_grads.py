"""Per-parameter gradient comparison of the CUDA fine-tune step against the oracle's autograd (GPU box only).

usage: python scripts/debug_train_grads.py [d64|full] [B] [T]
Prints loss, then one line per parameter (reverse graph order first) with relative L2 error and cosine.
"""

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unitspeech_oracle as O  # noqa: E402
from unitspeech_b200.training import FineTuner  # noqa: E402


def case_inputs(B, T, lengths, seed=21):
    _, mask, cond, spk, _ = O.harness_inputs(B, T, 2, seed=seed, lengths=lengths)
    g = torch.Generator().manual_seed(seed + 1)
    x0 = (torch.randn(B, 80, T, generator=g) * 0.5).clamp(-1, 1) * mask
    z = torch.randn(B, 80, T, generator=g)
    t = torch.rand(B, generator=g).clamp(0.05, 0.95)
    return x0, mask, cond, spk, z, t


def oracle_grads(params, x0, mask, cond, t, spk, z, dim, mults, device="cpu"):
    p = {k: v.detach().clone().to(device).requires_grad_(True) for k, v in params.items()}
    mv = lambda a: a.to(device)  # noqa: E731
    loss, _ = O.loss_t.__wrapped__(p, mv(x0), mv(mask), mv(cond), mv(t), mv(spk), mv(z), dim=dim, dim_mults=mults)
    loss.backward()
    return float(loss), {k: (v.grad.detach().cpu() if v.grad is not None else torch.zeros_like(v).cpu()) for k, v in p.items()}


def compare(ft, ref, verbose=True):
    got = ft.unscaled_grads()
    rows = []
    for k in reversed(list(ft.shapes)):
        a, b = got[k].detach().float().cpu().reshape(-1), ref[k].reshape(-1)
        nb = b.norm().item()
        err = (a - b).norm().item() / (nb + 1e-30)
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
        rows.append((k, err, cos, nb, a.norm().item()))
    if verbose:
        for k, err, cos, nb, na in rows:
            flag = "" if (err < 0.05 or nb < 1e-9) else "   <<<<"
            print(f"{k:55s} rel {err:9.3e} cos {cos:8.5f} |ref| {nb:9.3e} |got| {na:9.3e}{flag}")
    return rows


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "d64"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    dim, mults = (64, (1, 2)) if which == "d64" else (128, (1, 2, 4, 8))
    lengths = [T - 5 * i for i in range(B)]
    params = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=4.0)
    x0, mask, cond, spk, z, t = case_inputs(B, T, lengths)
    t0 = time.time()
    ref_loss, ref = oracle_grads(params, x0, mask, cond, t, spk, z, dim, mults, "cuda" if os.environ.get("ORACLE_CUDA") else "cpu")
    print(f"oracle loss {ref_loss:.6f} ({time.time() - t0:.1f}s)")
    ft = FineTuner(dim=dim, dim_mults=mults, loss_scale=float(os.environ.get("LOSS_SCALE", "8192")))
    ft.load_state_dict(params)
    ft.zero_grad()
    loss = ft.forward(x0, mask, cond, t, spk, z)
    torch.cuda.synchronize()
    print(f"cuda loss   {float(loss):.6f}")
    ft.backward()
    torch.cuda.synchronize()
    rows = compare(ft, ref)
    bad = [r for r in rows if r[1] >= 0.05 and r[3] >= 1e-9]
    print(f"{len(bad)} of {len(rows)} parameters above 5% relative error")
    # timing
    for _ in range(2):
        ft.train_step(x0, mask, cond, t, spk, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ft.train_step(x0, mask, cond, t, spk, z)
    e1.record()
    torch.cuda.synchronize()
    print(f"train_step {e0.elapsed_time(e1) / 5:.2f} ms, skipped {int(ft.skipped)} loss now {float(ft.loss):.6f}")


if __name__ == "__main__":
    main()
