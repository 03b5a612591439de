"""Shape/mask helpers on the decoder path (restated from unitspeech/util.py:20-59)."""

from __future__ import annotations

import torch


def sequence_mask(length: torch.Tensor, max_length=None) -> torch.Tensor:
    """unitspeech/util.py:20-24."""
    if max_length is None:
        max_length = int(length.max())
    x = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def fix_len_compatibility(length, num_downsamplings_in_unet: int = 3) -> int:
    """Smallest multiple of 2**num_downsamplings >= length (unitspeech/util.py:55-59)."""
    step = 2 ** num_downsamplings_in_unet
    return int(-(-int(length) // step) * step)


def generate_path(duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Monotonic alignment from integer durations (unitspeech/util.py:27-41).

    duration: (B, Tx); mask: (B, Tx, Ty) -> path (B, Tx, Ty) with path[b, i, j] = 1 iff frame j belongs to token i.
    """
    b, t_x, t_y = mask.shape
    cum = torch.cumsum(duration, 1)
    ar = torch.arange(t_y, dtype=cum.dtype, device=cum.device)
    upto = (ar.view(1, 1, t_y) < cum.view(b, t_x, 1)).to(mask.dtype)
    path = upto - torch.nn.functional.pad(upto, (0, 0, 1, 0))[:, :-1]
    return path * mask
