"""world_size-2 gloo test (CPU) of the clip sharding / frame all-gather logic used for N > 1 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from real3dportrait_b200.clip import render_clip, shard_range, steps_for


def _fake_step(idx):
    # a deterministic "frame" per global frame index
    return (idx.float().view(-1, 1, 1) * 10 + torch.arange(6).view(1, 2, 3)).contiguous()


def _worker(rank, world, port, n_frames, batch, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        clip = render_clip(_fake_step, n_frames, batch, world, rank, dist)
        ret[rank] = clip.clone()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize('n_frames,batch', [(16, 4), (13, 4), (3, 2)])
def test_two_rank_clip_equals_single_rank(n_frames, batch):
    single = render_clip(_fake_step, n_frames, batch, 1, 0)
    assert torch.equal(single, _fake_step(torch.arange(n_frames)))
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_frames, batch, ret), nprocs=world, join=True)
    for r in range(world):
        assert torch.equal(ret[r], single), r


def test_shard_ranges_partition_the_clip():
    for F in (1, 7, 128, 1024, 1025):
        for W in (1, 2, 3, 8):
            spans = [shard_range(F, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == F
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
            assert steps_for(F, W, 4) == (spans[0][1] - spans[0][0] + 3) // 4
