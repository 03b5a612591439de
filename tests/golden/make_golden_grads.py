"""Generates tests/golden/grads_*.npz: gradients and Adam trajectory of the UNMODIFIED reference for the fine-tune
objective -- UnitSpeech.loss_t (unitspeech/unitspeech.py:393-405), loss.backward(), clip_grad_norm_(max_norm=1) and
torch.optim.Adam (finetune.py:81,163-165).  Build container only (imports /root/reference through oracle/ref_shim.py).

Per case: K iterations on one fixed batch (z drawn by the reference's own torch.randn under torch.manual_seed(77 + i)).
Stored: the K losses and pre-clip total gradient norms; for iteration 0 every parameter's gradient L2 norm and its dot
product with a seeded N(0,1) vector (a full gradient set would be 10-500 MB), plus a few small gradients in full;
after the K steps every parameter's L2 distance from its initial value."""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_shim  # noqa: E402
from oracle import unitspeech_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

from train_cases import CASES, FULL_KEYS, case_inputs, projection  # noqa: E402


def main():
    torch.set_num_threads(8)
    U = ref_shim.load_reference()
    for name, (dim, mults, B, T, lengths, ts, s, lr, K) in CASES.items():
        params = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
        dec = U.UnitSpeech(n_feats=80, dim=dim, dim_mults=mults, beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=256)
        dec.load_state_dict(params, strict=True)
        dec.train()
        opt = torch.optim.Adam(params=dec.parameters(), lr=lr)
        x0, mask, cond, spk = case_inputs(B, T, lengths)
        t = torch.tensor(ts)
        losses, norms = [], []
        out = {}
        named = dict(dec.named_parameters())
        for i in range(K):
            dec.zero_grad()
            torch.manual_seed(77 + i)
            loss, _ = dec.loss_t(x0, mask, cond, t, spk)
            loss.backward()
            if i == 0:
                for k, v in named.items():
                    g = v.grad if v.grad is not None else torch.zeros_like(v)
                    out["gn/" + k] = np.float64(g.double().norm().item())
                    out["gp/" + k] = np.float64((g.double() * projection(k, g.shape).double()).sum().item())
                    if k in FULL_KEYS:
                        out["gf/" + k] = g.detach().numpy().copy()
            total = torch.nn.utils.clip_grad_norm_(dec.parameters(), max_norm=1)
            opt.step()
            losses.append(loss.item())
            norms.append(float(total))
        for k, v in named.items():
            out["dw/" + k] = np.float64((v.detach().double() - params[k].double()).norm().item())
        out["losses"] = np.array(losses, dtype=np.float64)
        out["norms"] = np.array(norms, dtype=np.float64)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, losses, norms)


if __name__ == "__main__":
    main()
