"""Decoder checkpoint carriage, in the reference's on-disk format.

The reference stores one `torch.save`d dict per decoder: {"model": state_dict, "spk_emb": (1, 256) or (B, 1, 256),
"mel_min": (80, 1) or (1, 80, 1), "mel_max": ..., "iteration": int} (train_STEP1.py:297-304; fine-tuned copies
overwrite model/mel_min/mel_max/spk_emb, finetune.py:169-173).  `load_decoder_checkpoint` builds the CUDA decoder from
such a file exactly the way inference.py:55-74,107-108,124 does, `save_decoder_checkpoint` writes one the reference's
own scripts can read back.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from .decoder import UnitSpeech


@dataclass
class DecoderBundle:
    decoder: UnitSpeech
    spk_emb: Optional[torch.Tensor]
    mel_min: Optional[torch.Tensor]
    mel_max: Optional[torch.Tensor]
    iteration: Optional[int]

    def denormalize(self, y: torch.Tensor) -> torch.Tensor:
        """inference.py:140 -- [-1, 1] -> log-mel with the checkpoint's per-bin range."""
        return (y + 1) / 2 * (self.mel_max.to(y.device) - self.mel_min.to(y.device)) + self.mel_min.to(y.device)


def _infer_config(sd):
    n_feats = sd["text_uncon"].shape[1]
    spk_emb_dim = sd["spk_uncon"].shape[2]
    dim = sd["estimator.mlp.2.weight"].shape[0]
    mults = []
    k = 0
    while f"estimator.downs.{k}.0.block1.block.0.weight" in sd:
        mults.append(sd[f"estimator.downs.{k}.0.block1.block.0.weight"].shape[0] // dim)
        k += 1
    return n_feats, dim, tuple(mults), spk_emb_dim


def load_decoder_checkpoint(path_or_dict, device="cuda", dim_mults: Optional[Sequence[int]] = None, beta_min=0.05,
                            beta_max=20.0, pe_scale=1000) -> DecoderBundle:
    """Reads a reference decoder checkpoint (path or already-loaded dict); architecture is inferred from the tensors."""
    ckpt = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu")
    sd = ckpt["model"] if "model" in ckpt else ckpt
    n_feats, dim, mults, spk_emb_dim = _infer_config(sd)
    dec = UnitSpeech(n_feats=n_feats, dim=dim, dim_mults=tuple(dim_mults) if dim_mults else mults, beta_min=beta_min,
                     beta_max=beta_max, pe_scale=pe_scale, spk_emb_dim=spk_emb_dim)
    dec.load_state_dict(sd, strict=True)
    dec = dec.to(device).eval()
    get = lambda k: ckpt[k] if isinstance(ckpt, dict) and k in ckpt else None  # noqa: E731
    return DecoderBundle(dec, get("spk_emb"), get("mel_min"), get("mel_max"), get("iteration"))


def save_decoder_checkpoint(path, decoder: UnitSpeech, spk_emb=None, mel_min=None, mel_max=None, iteration=None) -> None:
    """Writes {"model", "spk_emb", "mel_min", "mel_max", "iteration"} like train_STEP1.py:297-304."""
    d = {"model": {k: v.detach().cpu() for k, v in decoder.state_dict().items()}}
    for k, v in (("spk_emb", spk_emb), ("mel_min", mel_min), ("mel_max", mel_max)):
        if v is not None:
            d[k] = v.detach().cpu()
    if iteration is not None:
        d["iteration"] = int(iteration)
    torch.save(d, path)
