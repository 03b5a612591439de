"""Utterance sharding across the GPUs of one node (one process per GPU).

The sampler has no cross-utterance dependency (GroupNorm and attention are per sample, the update is elementwise), so
the only multi-GPU structure is: split the utterance list, run independent sampling loops, gather the mels.  There is
no collective inside the step loop.
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_items` owned by `rank`; sizes differ by at most one, earlier ranks get the extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_utterances(local: torch.Tensor, n_items: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All ranks contribute their (n_local, ...) shard (shard_range order); every rank gets the (n_items, ...) whole.

    Uses all_gather on equal-size padded shards (NCCL on GPUs, gloo in the CPU tests)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    max_n = max(e - b for b, e in sizes)
    b, e = sizes[rank]
    if local.shape[0] != e - b:
        raise ValueError(f"rank {rank} holds {local.shape[0]} utterances, expected {e - b}")
    # NCCL moves device memory only: a shard that came back through the host-buffer entry is staged through the
    # rank's GPU and the gathered result returned where the shard lived
    home = local.device
    if not local.is_cuda and dist.get_backend(group) == "nccl":
        local = local.to(torch.device("cuda", torch.cuda.current_device()))
    padded = local.new_zeros((max_n,) + tuple(local.shape[1:]))
    padded[: e - b] = local
    parts: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([parts[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0).to(home)


def sample_sharded(decoder, z, mask, cond, spk_emb, n_timesteps, text_gradient_scale=0.0, spk_gradient_scale=0.0,
                   noise: Optional[torch.Tensor] = None, group: Optional[dist.ProcessGroup] = None,
                   noise_is_local: bool = False) -> torch.Tensor:
    """Every rank passes the FULL batch; each samples its own shard with `decoder` and the result is gathered.

    ``noise`` is (n_timesteps, n_utterances, ...) for the full batch, or -- with ``noise_is_local=True`` -- only this rank's
    utterances in shard_range order (the per-step draws are n_timesteps times larger than everything else together, so a
    large job need not materialise them on every rank)."""
    n = z.shape[0]
    if not dist.is_available() or not dist.is_initialized():
        return decoder(z, mask, cond, spk_emb, n_timesteps, text_gradient_scale, spk_gradient_scale, noise=noise)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b, e = shard_range(n, rank, world)
    if noise is not None and noise.shape[1] != ((e - b) if noise_is_local else n):
        raise ValueError(f"noise holds {noise.shape[1]} utterances, expected {(e - b) if noise_is_local else n}")
    if e > b:
        local_noise = None if noise is None else (noise if noise_is_local else noise[:, b:e])
        # numerics decisions that depend on the call size (split-K mode of the CUDA decoder) are taken for the whole job,
        # so a sharded job equals the same job on one GPU bit for bit
        pin = getattr(decoder, "job_splitk_mode", None)
        if pin is not None:
            nb = 1 + (1 if float(text_gradient_scale) > 0 else 0) + (1 if float(spk_gradient_scale) > 0 else 0)
            saved = decoder.splitk_mode
            decoder.splitk_mode = pin(nb * n * z.shape[-1])
        try:
            out = decoder(z[b:e], mask[b:e], cond[b:e], spk_emb[b:e], n_timesteps, text_gradient_scale,
                          spk_gradient_scale, noise=local_noise)
        finally:
            if pin is not None:
                decoder.splitk_mode = saved
    else:
        out = z.new_zeros((0,) + tuple(z.shape[1:]))
    return gather_utterances(out, n, group)
