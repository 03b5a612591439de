"""Generates tests/golden/loss_*.npz: the diffusion training objective (forward value) of the UNMODIFIED reference --
UnitSpeech.loss_t (unitspeech/unitspeech.py:393-405) and UnitSpeech.fine_tune (:452-492).  Build container only."""

import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle import unitspeech_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# name: (dim, dim_mults, B, T, lengths, t values, out_scale); out_scale 4 makes the estimator output O(1), so the
# objective is sensitive to it (with the sampler harness scale it would be ~1.0 = E[z^2] regardless of the network)
LOSS_CASES = {
    "loss_d64": (64, (1, 2), 2, 16, (16, 11), (0.37, 0.81), 4.0),
    "loss_full": (128, (1, 2, 4, 8), 2, 24, (24, 19), (0.05, 0.6), 4.0),
}
# name: (dim, dim_mults, B, Ltext, T, y_lengths, segment_size)
FT_CASES = {
    "finetune_d64": (64, (1, 2), 2, 7, 40, (40, 33), 24),
    "finetune_d64_short": (64, (1, 2), 1, 5, 12, (12,), 16),      # utterance shorter than the segment: padded
}


FT_OUT_SCALE = 4.0


def loss_inputs(B, T, lengths, seed=21):
    _, mask, cond, spk, _ = O.harness_inputs(B, T, 2, seed=seed, lengths=lengths)
    g = torch.Generator().manual_seed(seed + 1)
    x0 = (torch.randn(B, 80, T, generator=g) * 0.5).clamp(-1, 1) * mask
    return x0, mask, cond, spk


def finetune_inputs(B, Lt, T, y_lengths, seed=31):
    g = torch.Generator().manual_seed(seed)
    cond_x = torch.randn(B, 80, Lt, generator=g).clamp(-1, 1)
    y = (torch.randn(B, 80, T, generator=g) * 0.5).clamp(-1, 1)
    y_lengths = torch.LongTensor(list(y_lengths))
    y_mask = (torch.arange(T).unsqueeze(0) < y_lengths.unsqueeze(1)).float().unsqueeze(1)
    # a monotonic hard alignment: frame j belongs to token floor(j * Lt / len)
    attn = torch.zeros(B, Lt, T)                      # fine_tune indexes attn[i, :, lower:upper] (:472): 3-D
    for b in range(B):
        for j in range(int(y_lengths[b])):
            attn[b, min(Lt - 1, j * Lt // int(y_lengths[b])), j] = 1.0
    spk = torch.randn(B, 1, 256, generator=g)
    spk = spk / spk.norm(dim=-1, keepdim=True)
    return cond_x, y * y_mask, y_mask, y_lengths, attn, spk


def main():
    torch.set_num_threads(8)
    U = ref_shim.load_reference()
    for name, (dim, mults, B, T, lengths, ts, s) in LOSS_CASES.items():
        params = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=s)
        dec = U.UnitSpeech(n_feats=80, dim=dim, dim_mults=mults, beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=256)
        dec.load_state_dict(params, strict=True)
        x0, mask, cond, spk = loss_inputs(B, T, lengths)
        t = torch.tensor(ts)
        torch.manual_seed(77)                      # forward_diffusion draws z = randn(x0.shape) from the global RNG
        with torch.no_grad():
            loss, xt = dec.loss_t(x0, mask, cond, t, spk)
        np.savez(os.path.join(OUT, name + ".npz"), loss=np.float64(loss.item()), xt=xt.numpy())
        print(name, float(loss))
    for name, (dim, mults, B, Lt, T, y_lengths, seg) in FT_CASES.items():
        params = O.harness_params(dim=dim, dim_mults=mults, seed=1234, out_scale=FT_OUT_SCALE)
        dec = U.UnitSpeech(n_feats=80, dim=dim, dim_mults=mults, beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=256)
        dec.load_state_dict(params, strict=True)
        cond_x, y, y_mask, yl, attn, spk = finetune_inputs(B, Lt, T, y_lengths)
        random.seed(5)
        torch.manual_seed(78)
        with torch.no_grad():
            loss = dec.fine_tune(cond_x, y, y_mask, yl, T, attn, spk, seg, 80)
        np.savez(os.path.join(OUT, name + ".npz"), loss=np.float64(loss.item()))
        print(name, float(loss))


if __name__ == "__main__":
    main()
