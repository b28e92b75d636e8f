"""Clip-level sharding of the frame loop over the GPUs of one node (BASELINE configs[3]).

Frames of a clip are independent: rank r of `world` renders the contiguous block `shard_range(F, world, r)` in batches
of `batch` frames; after every step the ranks exchange that step's frames with ONE `all_gather_into_tensor` (NCCL on
GPUs, gloo in the CPU tests) and drop them into the output clip at their global frame indices.  Every rank always
contributes exactly `batch` frames per step (ragged tails are padded by repeating the shard's last frame and discarded
on arrival), so the collective is uniform."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch


def shard_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`; the first `n_frames % world` ranks get one extra frame."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def steps_for(n_frames: int, world: int, batch: int) -> int:
    longest = shard_range(n_frames, world, 0)[1]
    return (longest + batch - 1) // batch


def render_clip(step_fn: Callable[[torch.Tensor], torch.Tensor], n_frames: int, batch: int, world: int, rank: int, dist=None,
                out: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """step_fn(frame_indices[batch] int64) -> frames [batch, ...] for THIS rank's frames.  Returns the whole clip
    [n_frames, ...] on every rank."""
    lo, hi = shard_range(n_frames, world, rank)
    clip = out
    gathered = None
    for s in range(steps_for(n_frames, world, batch)):
        idx = torch.arange(lo + s * batch, lo + (s + 1) * batch)
        idx = idx.clamp(max=max(hi - 1, lo)).clamp(max=n_frames - 1)          # pad the ragged tail with the last frame
        frames = step_fn(idx)
        if clip is None:
            clip = torch.empty((n_frames,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device if device is None else device)
        if world > 1:
            if gathered is None:
                gathered = torch.empty((world * batch,) + tuple(frames.shape[1:]), dtype=frames.dtype, device=frames.device)
            dist.all_gather_into_tensor(gathered, frames.contiguous())
        else:
            gathered = frames
        for r in range(world):
            rlo, rhi = shard_range(n_frames, world, r)
            a = rlo + s * batch
            b = min(a + batch, rhi)
            if b > a:
                clip[a:b] = gathered[r * batch: r * batch + (b - a)]
    return clip
