"""CPU, world_size 2, gloo: the N>1 host logic (ray sharding, LR gather, gradient bucket all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_sr_b200.parallel import allreduce_mean_, epoch_indices, rank_batches, render_sharded, shard_bounds
from oracle import nerf_oracle as O


def test_shard_bounds_respect_subpixel_groups():
    for n_lr, s, world in ((10, 2, 4), (47628, 2, 8), (40000, 4, 8), (3, 2, 8)):
        n = n_lr * s * s
        b = shard_bounds(n, world, s * s)
        assert b[0][0] == 0 and b[-1][1] == n
        for (lo, hi), (lo2, _) in zip(b, b[1:] + [(n, n)]):
            assert hi == lo2 and lo % (s * s) == 0 and hi % (s * s) == 0
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= s * s
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(rays, s):
    # deterministic stand-in for the CUDA renderer: per-ray value, then the reference's box average
    rgb = torch.stack([rays[:, 0], rays[:, 1] * 2, rays[:, 2] + 1], 1)
    depth = rays[:, 6] + rays[:, 7]
    return O.box_average(rgb, s), O.box_average(depth, s)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = 2
        rays = O.synthetic_rays(4 * 37, 3, "blender")          # 37 LR pixels: uneven split over 2 ranks
        (rgb, depth), bounds = render_sharded(lambda r: _fake_render(r, s), rays, s)
        ref_rgb, ref_depth = _fake_render(rays, s)
        ok = torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth)
        # gradient bucket: mean over ranks
        g = [torch.full((5, 3), float(rank + 1)), torch.full((7,), float(10 * (rank + 1)))]
        allreduce_mean_(g)
        ok = ok and torch.allclose(g[0], torch.full((5, 3), 1.5)) and torch.allclose(g[1], torch.full((7,), 15.0))
        # training-batch sharding: every rank derives its epoch indices alone; together they are a permutation (+ wrap pad)
        mine = epoch_indices(101, world, rank, epoch=3, seed=7)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        allidx = torch.stack(parts, 1).reshape(-1)                  # interleave back: rank r holds entries r, r+world, ...
        ok = ok and allidx.numel() == 102 and sorted(allidx[:101].tolist()) == list(range(101)) and int(allidx[101]) == int(allidx[0])
        q.put((rank, bool(ok), bounds))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_render_and_grad_bucket():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == [(0, 76), (76, 148)]


@pytest.mark.parametrize("n,world,shuffle,drop_last", [(101, 2, True, False), (47628, 8, True, False), (10, 4, False, False),
                                                       (3, 8, True, False), (101, 4, True, True), (64, 8, True, True)])
def test_epoch_indices_equal_torch_distributed_sampler(n, world, shuffle, drop_last):
    """The reference's DDP loader is torch's DistributedSampler(seed=opt.seed) (data/__init__.py:94-101)."""
    from torch.utils.data import DistributedSampler
    data = list(range(n))
    for epoch in (0, 5):
        for rank in range(world):
            ref = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=shuffle, seed=11, drop_last=drop_last)
            ref.set_epoch(epoch)
            got = epoch_indices(n, world, rank, epoch, seed=11, shuffle=shuffle, drop_last=drop_last)
            assert got.tolist() == list(iter(ref)), (epoch, rank)


def test_rank_batches_follow_the_ddp_loader():
    from torch.utils.data import DataLoader, DistributedSampler
    n, batch, world = 1000, 64, 4
    data = torch.arange(n)
    for rank in range(world):
        samp = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=True, seed=0)
        samp.set_epoch(2)
        ref = [b.tolist() for b in DataLoader(data, batch_size=batch // world, sampler=samp, drop_last=True)]
        got = [b.tolist() for b in rank_batches(n, batch, world, rank, epoch=2, seed=0)]
        assert got == ref and all(len(b) == 16 for b in got)
        ref_keep = [b.tolist() for b in DataLoader(data, batch_size=batch // world, sampler=samp, drop_last=False)]
        assert [b.tolist() for b in rank_batches(n, batch, world, rank, epoch=2, seed=0, keep_last=True)] == ref_keep
    with pytest.raises(ValueError):
        list(rank_batches(n, 66, 4, 0))
    with pytest.raises(ValueError):
        epoch_indices(10, 2, 2)
